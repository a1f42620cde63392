timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --section SchedulerStats --section LaunchStats --section Occupancy --import-source on --clock-control none -k regex:bucket_dedupe -c 1 -o /tmp/bd -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
ncu -i /tmp/bd.ncu-rep --page raw --csv > gpurun_out/e21_bd_raw.csv 2>/dev/null
ncu -i /tmp/bd.ncu-rep --page source --csv > gpurun_out/e21_bd_source.csv 2>/dev/null
ls -la gpurun_out/e21*

timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --section SchedulerStats --section LaunchStats --section Occupancy --import-source on --clock-control none -k regex:log_dedupe -c 1 -o gpurun_out/r01_dedupe_v1 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/e5_ncu.log 2>&1
tail -5 gpurun_out/e5_ncu.log

timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section SchedulerStats --section LaunchStats --section Occupancy --import-source on --clock-control none -k regex:ladder_kernel -s 1 -c 1 -o gpurun_out/e15_ladder_rot -f python profiles/scripts/prof_ladder.py rotated25 30 > gpurun_out/e15.log 2>&1
tail -2 gpurun_out/e15.log

# the measurements committed under profiles/ for round 1 (kernel generation v8)
set -x
python bench.py > gpurun_out/r01_bench_v8_full.json 2> gpurun_out/v8_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r01_bench_v8_reference_arm.json 2> gpurun_out/v8_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_stdc_v8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,sm__warps_active.avg.per_cycle_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 900 ncu --replay-mode application --metrics $M -k regex:"stdc_fast|dedupe" --clock-control none -c 2 --csv --log-file gpurun_out/r01_ncu_fullsize_stdc_v8.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:stdc_fast -c 1 -o /tmp/r01_stdc_v8 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --samples 2000 > /dev/null 2>&1
ncu -i /tmp/r01_stdc_v8.ncu-rep --page raw --csv > gpurun_out/r01_ncu_full_stdc_v8_raw.csv 2>/dev/null
timeout 900 python bench_configs.py --config all --out gpurun_out/r01_configs_v8.jsonl > gpurun_out/v8_configs.log 2>&1
cat gpurun_out/r01_bench_v8_full.json
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --section SchedulerStats --section LaunchStats --section Occupancy --import-source on --clock-control none -k regex:bucket_dedupe -c 1 -o /tmp/r01_bucket_dedupe_v8 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
ncu -i /tmp/r01_bucket_dedupe_v8.ncu-rep --page raw --csv > gpurun_out/r01_ncu_bucket_dedupe_v8_raw.csv 2>/dev/null
ncu -i /tmp/r01_bucket_dedupe_v8.ncu-rep --page source --csv > gpurun_out/r01_ncu_bucket_dedupe_v8_source.csv 2>/dev/null
for c in rotated25 xzzx21_biased; do timeout 600 ncu --set full --import-source on --clock-control none -k regex:ladder_kernel -s 1 -c 1 -o /tmp/r01_ladder_${c}_v8 -f python profiles/scripts/prof_ladder.py $c 100 > /dev/null 2>&1; ncu -i /tmp/r01_ladder_${c}_v8.ncu-rep --page raw --csv > gpurun_out/r01_ncu_full_ladder_${c}_v8_raw.csv 2>/dev/null; done

# XZZX top rung: cache refresh only for stabilizers on a diagonal; HEAD captures of the headline bench for the roofline
timeout 900 python -m pytest tests/test_gpu_native.py -q -k "ladder or pteq or lane_split" > gpurun_out/r2k_native.log 2>&1; tail -3 gpurun_out/r2k_native.log
python profiles/scripts/prof_ladder.py xzzx21_biased 400 4736 0.5 8 > gpurun_out/r2k_lt.txt 2>&1
python profiles/scripts/prof_ladder.py xzzx21_alpha 400 4736 0.5 8 >> gpurun_out/r2k_lt.txt 2>&1
cat gpurun_out/r2k_lt.txt
python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; cut -c1-300 gpurun_out/r2k_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_stdc.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,sm__warps_active.avg.per_cycle_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 900 ncu --replay-mode application --metrics $M -k regex:"stdc_fast|dedupe" --clock-control none -c 2 --csv --log-file gpurun_out/r02_ncu_fullsize_stdc.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:stdc_fast -c 1 -o /tmp/r02_stdc -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --samples 2000 > /dev/null 2>&1
ncu -i /tmp/r02_stdc.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_stdc_raw.csv 2>/dev/null
ncu -i /tmp/r02_stdc.ncu-rep --page source --csv > gpurun_out/r02_ncu_full_stdc_source.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:pt_kernel -c 1 -s 1 -o /tmp/r2k_pt_xzzx -f python profiles/scripts/prof_ladder.py xzzx21_biased 100 4736 0.5 8 > gpurun_out/r2k_ncu.log 2>&1
ncu -i /tmp/r2k_pt_xzzx.ncu-rep --page raw --csv > gpurun_out/r2k_pt_xzzx_raw.csv 2>/dev/null
ncu -i /tmp/r2k_pt_xzzx.ncu-rep --page source --csv > gpurun_out/r2k_pt_xzzx_source.csv 2>/dev/null

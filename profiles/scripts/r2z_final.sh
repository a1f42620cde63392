# Round-2 final measurements at HEAD: GPU suite, bench lines (own arm, reference arm), ncu launch list, full-size
# counters of the chain + dedupe kernels, --set full captures (chain kernel with source, 64-bit chain kernel, tempering kernel)
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; tail -3 gpurun_out/r02_gpu_tests.log
python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; cut -c1-200 gpurun_out/r02_bench_full.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; cut -c1-200 gpurun_out/r02_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_stdc.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,sm__warps_active.avg.per_cycle_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 900 ncu --replay-mode application --metrics $M -k regex:"stdc_fast|dedupe" --clock-control none -c 2 --csv --log-file gpurun_out/r02_ncu_fullsize_stdc.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:stdc_fast -c 1 -o /tmp/r02_stdc -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --samples 2000 > /dev/null 2>&1
ncu -i /tmp/r02_stdc.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_stdc_raw.csv 2>/dev/null
ncu -i /tmp/r02_stdc.ncu-rep --page source --csv > gpurun_out/r02_ncu_full_stdc_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:stdc_fast -c 1 -s 1 -o /tmp/r02_p21 -f python profiles/scripts/prof_planar21.py 21 3000 > gpurun_out/r02_p21.log 2>&1
ncu -i /tmp/r02_p21.ncu-rep --page raw --csv > gpurun_out/r02_ncu_stdc_planar21_u64_raw.csv 2>/dev/null
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pt_kernel -c 1 -s 1 -o /tmp/r02_pt_x -f python profiles/scripts/prof_ladder.py xzzx21_biased 100 4736 0.5 > gpurun_out/r02_pt_x.log 2>&1
ncu -i /tmp/r02_pt_x.ncu-rep --page raw --csv > gpurun_out/r02_ncu_pt_xzzx21_biased_raw.csv 2>/dev/null
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pt_kernel -c 1 -s 1 -o /tmp/r02_pt_r -f python profiles/scripts/prof_ladder.py rotated25 100 4736 0.5 > gpurun_out/r02_pt_r.log 2>&1
ncu -i /tmp/r02_pt_r.ncu-rep --page raw --csv > gpurun_out/r02_ncu_pt_rotated25_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20

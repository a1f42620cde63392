timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/e13_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e13_bench.json 2>gpurun_out/e13_bench.err
python -c "
import json
d=json.load(open('gpurun_out/e13_bench.json')); print('%.3e'%d['value'], '%.1f'%d['ms_per_step'], '%.1f'%d['roofline']['kernel_ms_per_launch'], '%.3e'%d['e2e']['value'], d['config']['syndromes_per_step_per_gpu'], d['roofline']['other_kernels_ms_per_step'])"
M=sm__warps_active.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --replay-mode application --metrics $M -k regex:"stdc_fast|log_dedupe" --clock-control none -c 2 --csv --log-file gpurun_out/e13_ncu.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
grep -h "stdc_fast\|dedupe" gpurun_out/e13_ncu.csv | awk -F'","' '{print substr($5,1,30), $13, $NF}'
cat gpurun_out/e13_tests.log

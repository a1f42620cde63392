timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stdc or dedupe or strc" 2>&1 | tail -2 > gpurun_out/e12_tests.log
B="timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/e12_$name.json 2>gpurun_out/e12_$name.err; python -c "
import json
d=json.load(open('gpurun_out/e12_$name.json')); print('$name', '%.3e'%d['value'], '%.1f'%d['ms_per_step'], '%.1f'%d['roofline']['kernel_ms_per_launch'], d['config']['syndromes_per_step_per_gpu'])"; }
run fair256 X=1
run fair64 QECMC_DEBUG_SYNC_CALLS=64
run fair1024 QECMC_DEBUG_SYNC_CALLS=1024
run nofair QECMC_DEBUG_FAIR=0
run t256_fair256 QECMC_DEBUG_T=256
M=sm__warps_active.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
timeout 300 ncu --replay-mode application --metrics $M -k regex:stdc_fast --clock-control none -c 1 --csv --log-file gpurun_out/e12_ncu.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
grep -h "stdc_fast" gpurun_out/e12_ncu.csv | awk -F'","' '{print $13, $NF}'
cat gpurun_out/e12_tests.log

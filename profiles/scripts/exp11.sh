B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
V=$PWD/mcmc-qec-toric-rl_b200/csrc/_variants/libqecmc_t1024.so
run() { name=$1; shift; env "$@" $B > gpurun_out/e11_$name.json 2>gpurun_out/e11_$name.err; python -c "
import json
d=json.load(open('gpurun_out/e11_$name.json')); print('$name', '%.3e'%d['value'], '%.1f'%d['ms_per_step'], '%.1f'%d['roofline']['kernel_ms_per_launch'], d['config']['syndromes_per_step_per_gpu'])"; }
run base X=1
B="$B --syndromes 148"; run t1024 QECMC_LIB=$V QECMC_DEBUG_T=1024; run t1024_sync QECMC_LIB=$V QECMC_DEBUG_T=1024 QECMC_DEBUG_SYNC_CALLS=256
M=sm__warps_active.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
ncu --replay-mode application --metrics $M -k regex:stdc_fast --clock-control none -c 1 --csv --log-file gpurun_out/e11_ncu_base.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
QECMC_LIB=$V QECMC_DEBUG_T=1024 QECMC_DEBUG_SYNC_CALLS=256 ncu --replay-mode application --metrics $M -k regex:stdc_fast --clock-control none -c 1 --csv --log-file gpurun_out/e11_ncu_t1024.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --syndromes 148 > /dev/null 2>&1
grep -h "stdc_fast" gpurun_out/e11_ncu_base.csv gpurun_out/e11_ncu_t1024.csv | awk -F'","' '{print $13, $NF}'

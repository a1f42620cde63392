python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/e4_tests.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/e4_mode4.json 2>gpurun_out/e4_mode4.err
QECMC_DEBUG_INSERT_MODE=2 $B > gpurun_out/e4_mode2.json 2>gpurun_out/e4_mode2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/e4_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/e4_ncu_launches.log 2>&1
for f in gpurun_out/e4_mode*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', '%.3e'%d['value'], '%.1f'%d['ms_per_step'], '%.1f'%d['roofline']['kernel_ms_per_launch'], '%.3e'%d['e2e']['value'], d['chain_stats'])"; done
cat gpurun_out/e4_tests.log
grep -v "^==" gpurun_out/e4_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12

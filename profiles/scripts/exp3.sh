set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/e3_tests.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
for m in 2 1 0 3; do QECMC_DEBUG_INSERT_MODE=$m $B > gpurun_out/e3_mode$m.json 2>gpurun_out/e3_mode$m.err; done
for f in gpurun_out/e3_mode*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', '%.3e'%d['value'], '%.1f'%d['ms_per_step'], '%.1f'%d['roofline']['kernel_ms_per_launch'], '%.3e'%d['e2e']['value'], d['chain_stats'])"; done
cat gpurun_out/e3_tests.log

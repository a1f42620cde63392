timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/e23_tests.log
timeout 600 python bench_configs.py --config planar_sweep > gpurun_out/e23_planar.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e23_bench.json 2>/dev/null
cat gpurun_out/e23_tests.log
python -c "
import json
d=json.load(open('gpurun_out/e23_bench.json')); print('%.4e'%d['value'], '%.1f'%d['ms_per_step'], '%.1f'%d['roofline']['kernel_ms_per_launch'])
c=json.loads(open('gpurun_out/e23_planar.log').read().strip().splitlines()[-1])
for r in c['points']: print(r['d'], r['p'], r['syndromes'], '%.3e'%r['steps_per_s'], round(r['syndromes_per_s']), round(r['logical_failure_rate'],4))
"

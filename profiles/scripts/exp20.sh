timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/e20_tests.log
timeout 600 python bench_configs.py --config toric5 > gpurun_out/e20_toric5.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e20_bench.json 2>/dev/null
cat gpurun_out/e20_tests.log; cut -c1-700 gpurun_out/e20_toric5.log; python -c "
import json
d=json.load(open('gpurun_out/e20_bench.json')); print('%.4e'%d['value'], '%.1f'%d['ms_per_step'], '%.4e'%d['e2e']['value'], d['roofline']['frac'])"

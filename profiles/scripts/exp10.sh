python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/e10_tests.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/e10_bench.json 2>gpurun_out/e10_bench.err
python -c "
import json
d=json.load(open('gpurun_out/e10_bench.json')); print('%.3e'%d['value'], '%.1f'%d['ms_per_step'], '%.1f'%d['roofline']['kernel_ms_per_launch'], '%.3e'%d['e2e']['value'], d['roofline']['call_ms'], d['roofline']['other_kernels_ms_per_step'])"
cat gpurun_out/e10_tests.log

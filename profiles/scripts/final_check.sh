timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_1gpu.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/final_bench_1gpu.json')); print(d['n_gpus'], '%.4e'%d['value'], '%.1f'%d['ms_per_step'], '%.4e'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['cpu_baseline']['value'])"

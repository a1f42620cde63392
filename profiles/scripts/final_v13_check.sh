timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r01_bench_v13_full.json 2> gpurun_out/v13_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_v13_reference_arm.json 2> gpurun_out/v13_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_stdc_v13.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python -c "
import json
d=json.load(open('gpurun_out/r01_bench_v13_full.json')); print(d['n_gpus'], '%.4e'%d['value'], '%.1f'%d['ms_per_step'], '%.4e'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline'].get('value_1core'), d['clocks'], d['gpu_launches'])
r=json.load(open('gpurun_out/r01_bench_v13_reference_arm.json')); print(r['impl'], '%.3e'%r['value'], r['cpu_baseline']['cores'])"
grep -c stdc_fast gpurun_out/r01_launches_stdc_v13.csv

# packed kernel: CTAs balanced over whole rounds; toric d = 17-21 again, packed A/B test, planar sweep on one GPU
timeout 900 python -m pytest tests/test_gpu_native.py -q -x -k "packed_lattice or stdc_equals_oracle" > gpurun_out/r2zo_tests.log 2>&1; tail -2 gpurun_out/r2zo_tests.log
sed -n '/^python - > gpurun_out\/r2zn_toric.txt/,/^P$/p' profiles/scripts/r2zn.sh | sed 's/r2zn_toric/r2zo_toric/' > /tmp/t.sh; bash /tmp/t.sh; cat gpurun_out/r2zo_toric.txt
python bench_configs.py --config planar_sweep --syndromes 125000 --out gpurun_out/r2zo_sweep.jsonl > gpurun_out/r2zo_sweep.log 2>&1
python - <<'P'
import json
j = json.loads(open("gpurun_out/r2zo_sweep.jsonl").readline())
print(j["syndromes"], round(j["seconds"], 2), "s", round(j["steps_per_s"] / 1e11, 3), "e11 steps/s", j.get("syndromes_per_item"))
P

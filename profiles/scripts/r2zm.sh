# bucket logs sized for 55 % of the samples when the budget holds less than one round: GPU suite, sweep on one GPU
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2zm_tests.log 2>&1; tail -3 gpurun_out/r2zm_tests.log
python bench_configs.py --config planar_sweep --syndromes 125000 --out gpurun_out/r2zm_sweep_new.jsonl > gpurun_out/r2zm_new.log 2>&1; tail -c 200 gpurun_out/r2zm_new.log
python - <<'P'
import json
j = json.loads(open("gpurun_out/r2zm_sweep_new.jsonl").readline())
print(j["syndromes"], round(j["seconds"], 2), "s", round(j["steps_per_s"] / 1e11, 3), "e11 steps/s", j.get("syndromes_per_item"))
P

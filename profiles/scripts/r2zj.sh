# packed-lattice kernel + plan-sized sweep items: full GPU suite, then the planar sweep on one GPU (1/8 of the job) old vs new
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2zj_tests.log 2>&1; tail -3 gpurun_out/r2zj_tests.log
python bench_configs.py --config planar_sweep --syndromes 125000 --ab-old-sizing --out gpurun_out/r2zj_sweep_old.jsonl > gpurun_out/r2zj_old.log 2>&1; tail -c 300 gpurun_out/r2zj_old.log
python bench_configs.py --config planar_sweep --syndromes 125000 --out gpurun_out/r2zj_sweep_new.jsonl > gpurun_out/r2zj_new.log 2>&1; tail -c 300 gpurun_out/r2zj_new.log
python - <<'P'
import json
for f in ("old", "new"):
    j = json.loads(open("gpurun_out/r2zj_sweep_%s.jsonl" % f).readline())
    print(f, j["syndromes"], round(j["seconds"], 2), "s", round(j["steps_per_s"] / 1e11, 3), "e11 steps/s", j.get("syndromes_per_item"))
P

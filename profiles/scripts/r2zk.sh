# config 5 at scale with the packed-lattice kernel and plan-sized items: 1 000 032 planar syndromes on 8 GPUs
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
rm -f gpurun_out/r02_config5_planar_sweep_8gpu.jsonl
timeout 900 $TR --nproc-per-node 8 --master-port 29601 bench_configs.py --config planar_sweep --syndromes 1000000 --out gpurun_out/r02_config5_planar_sweep_8gpu.jsonl > gpurun_out/r2zk_sweep.log 2>&1; tail -c 300 gpurun_out/r2zk_sweep.log
python - <<'P'
import json
j = json.loads(open("gpurun_out/r02_config5_planar_sweep_8gpu.jsonl").readline())
print(j["n_gpus"], j["syndromes"], round(j["seconds"], 2), "s", round(j["syndromes_per_s"]), "syndromes/s", round(j["steps_per_s"] / 1e12, 3), "e12 steps/s", j.get("syndromes_per_item"), j["per_rank_seconds"])
P

# 8-GPU refresh at HEAD: weak-scaling bench, config 5 at scale, config 2 as a fixed job at N = 8, 4, 2, 1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29600 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r2ze_bench8.err; cut -c1-250 gpurun_out/r02_bench_8gpu.json
rm -f gpurun_out/r02_config5_planar_sweep_8gpu.jsonl gpurun_out/r02_config2_strong_scaling.jsonl
timeout 900 $TR --nproc-per-node 8 --master-port 29601 bench_configs.py --config planar_sweep --syndromes 1000000 --out gpurun_out/r02_config5_planar_sweep_8gpu.jsonl > gpurun_out/r2ze_sweep.log 2>&1; tail -c 300 gpurun_out/r2ze_sweep.log
for n in 8 4 2 1; do
timeout 600 $TR --nproc-per-node $n --master-port 2961$n bench_configs.py --config toric15_strong --syndromes 10000 --out gpurun_out/r02_config2_strong_scaling.jsonl > gpurun_out/r2ze_strong_$n.log 2>&1; tail -c 200 gpurun_out/r2ze_strong_$n.log
done

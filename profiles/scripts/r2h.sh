# BASELINE config 5 at scale (1M planar syndromes, 8 GPUs, gather) and config 2 as a fixed job at N = 8, 4, 2, 1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29601 bench_configs.py --config planar_sweep --syndromes 1000000 --out gpurun_out/r02_config5_planar_sweep_8gpu.jsonl > gpurun_out/r2h_sweep.log 2>&1; tail -c 400 gpurun_out/r2h_sweep.log
for n in 8 4 2 1; do
timeout 600 $TR --nproc-per-node $n --master-port 2961$n bench_configs.py --config toric15_strong --syndromes 10000 --out gpurun_out/r02_config2_strong_scaling.jsonl > gpurun_out/r2h_strong_$n.log 2>&1; tail -c 300 gpurun_out/r2h_strong_$n.log
done

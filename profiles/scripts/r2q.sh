# weak scaling of the headline bench on 8 GPUs at HEAD
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r2q_bench8.err; cut -c1-300 gpurun_out/r02_bench_8gpu.json; tail -2 gpurun_out/r2q_bench8.err

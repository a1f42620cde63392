# sanity of the sharded drivers on one GPU (small), and the fixed CI test
timeout 600 python bench_configs.py --config planar_sweep --syndromes 44000 --out gpurun_out/r2g_sweep_1gpu.jsonl > gpurun_out/r2g_sweep.log 2>&1; tail -c 1500 gpurun_out/r2g_sweep.log
timeout 600 python bench_configs.py --config toric15_strong --syndromes 600 --out gpurun_out/r2g_strong_1gpu.jsonl > gpurun_out/r2g_strong.log 2>&1; tail -c 800 gpurun_out/r2g_strong.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench_configs.py --config toric15_strong --syndromes 300 > gpurun_out/r2g_strong_tr.log 2>&1; tail -c 600 gpurun_out/r2g_strong_tr.log
timeout 900 python -m pytest tests/test_gpu_native.py -q -k "binomial" > gpurun_out/r2g_ci.log 2>&1; tail -3 gpurun_out/r2g_ci.log

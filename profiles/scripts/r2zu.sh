# the single-GPU table of the BASELINE configurations at HEAD
timeout 900 python bench_configs.py --config all --out gpurun_out/r02_configs.jsonl > gpurun_out/r2zu.log 2>&1; tail -c 1500 gpurun_out/r2zu.log

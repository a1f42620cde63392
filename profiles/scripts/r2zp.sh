# final HEAD verification: GPU suite log, launch list, configs 3 / 4 decoded to convergence again at HEAD
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; tail -3 gpurun_out/r02_gpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_stdc.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 400 python profiles/scripts/run_config34.py gpu rotated25 4000000 > gpurun_out/r2zp_rot.json 2> gpurun_out/r2zp_rot.err; cat gpurun_out/r2zp_rot.json | cut -c1-900; tail -2 gpurun_out/r2zp_rot.err
timeout 400 python profiles/scripts/run_config34.py gpu xzzx21_biased 2000000 > gpurun_out/r2zp_xb.json 2> gpurun_out/r2zp_xb.err; cat gpurun_out/r2zp_xb.json | cut -c1-900; tail -2 gpurun_out/r2zp_xb.err

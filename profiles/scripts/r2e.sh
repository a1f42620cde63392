# full GPU suite at HEAD, then configs 3 / 4 decoded to convergence (cap 2e6 ladder steps)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; tail -3 gpurun_out/r2e_tests.log
timeout 400 python profiles/scripts/run_config34.py gpu rotated25 2000000 > gpurun_out/r2e_rot.json 2> gpurun_out/r2e_rot.err; cat gpurun_out/r2e_rot.json; tail -3 gpurun_out/r2e_rot.err
timeout 400 python profiles/scripts/run_config34.py gpu xzzx21_biased 2000000 > gpurun_out/r2e_xb.json 2> gpurun_out/r2e_xb.err; cat gpurun_out/r2e_xb.json; tail -3 gpurun_out/r2e_xb.err
timeout 400 python profiles/scripts/run_config34.py gpu xzzx21_alpha 2000000 > gpurun_out/r2e_xa.json 2> gpurun_out/r2e_xa.err; cat gpurun_out/r2e_xa.json; tail -3 gpurun_out/r2e_xa.err

timeout 600 python -m pytest tests/test_gpu_native.py -q -x -k "last_plan" > gpurun_out/r2zq_tests.log 2>&1; tail -12 gpurun_out/r2zq_tests.log

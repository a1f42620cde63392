timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/e22_tests.log
cat gpurun_out/e22_tests.log

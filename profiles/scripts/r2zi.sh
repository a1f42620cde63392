# packed-lattice kernel at realistic chain lengths (40 000 samples = 200 000 steps per chain)
timeout 900 python -m pytest tests/test_gpu_native.py -q -x -k "packed_lattice" > gpurun_out/r2zi_tests.log 2>&1; tail -2 gpurun_out/r2zi_tests.log
python profiles/scripts/prof_packed.py 21 17 > gpurun_out/r2zi_packed.txt 2>&1; cat gpurun_out/r2zi_packed.txt

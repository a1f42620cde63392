# packed-lattice chain kernel for 17 <= L <= 24: oracle parity, A/B against the 64-bit row-word kernel, timing
timeout 1200 python -m pytest tests/test_gpu_native.py -q -x -k "stdc_equals_oracle or packed_lattice or strc_equals or single_temp" > gpurun_out/r2zf_tests.log 2>&1; tail -15 gpurun_out/r2zf_tests.log
python profiles/scripts/prof_packed.py > gpurun_out/r2zf_packed.txt 2>&1; cat gpurun_out/r2zf_packed.txt

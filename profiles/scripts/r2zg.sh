# packed-lattice kernel: timing with warmed allocations, ncu of the d = 21 launch
timeout 900 python -m pytest tests/test_gpu_native.py -q -x -k "packed_lattice" > gpurun_out/r2zg_tests.log 2>&1; tail -3 gpurun_out/r2zg_tests.log
python profiles/scripts/prof_packed.py > gpurun_out/r2zg_packed.txt 2>&1; cat gpurun_out/r2zg_packed.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:stdc_pk -c 1 -s 2 -o /tmp/r2zg_pk -f python profiles/scripts/prof_planar21.py 21 3000 > gpurun_out/r2zg_ncu.log 2>&1; tail -2 gpurun_out/r2zg_ncu.log
ncu -i /tmp/r2zg_pk.ncu-rep --page raw --csv > gpurun_out/r2zg_pk_raw.csv 2>/dev/null
ncu -i /tmp/r2zg_pk.ncu-rep --page source --csv > gpurun_out/r2zg_pk_source.csv 2>/dev/null

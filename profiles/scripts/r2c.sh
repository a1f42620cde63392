# rung-major tempering kernel: lanes per top replica, old kernel for reference, ncu of the default
for lt in 2 4 8 16 32; do python profiles/scripts/prof_ladder.py rotated25 400 4736 0.5 $lt; done > gpurun_out/r2c_lt.txt 2>&1
python profiles/scripts/prof_ladder.py rotated25 400 4736 0.5 8 1 >> gpurun_out/r2c_lt.txt 2>&1
python profiles/scripts/prof_ladder.py rotated25 400 4736 0.0 8 >> gpurun_out/r2c_lt.txt 2>&1
for lt in 4 8 16; do python profiles/scripts/prof_ladder.py xzzx21_biased 400 4736 0.5 $lt; done >> gpurun_out/r2c_lt.txt 2>&1
python profiles/scripts/prof_ladder.py toric15 400 4736 0.5 8 >> gpurun_out/r2c_lt.txt 2>&1
cat gpurun_out/r2c_lt.txt
ncu --set full --import-source on --clock-control none -k regex:pt_kernel -c 1 -s 1 -o gpurun_out/r2c_pt_rot25 -f python profiles/scripts/prof_ladder.py rotated25 100 4736 0.5 8 > gpurun_out/r2c_ncu.log 2>&1
ncu -i gpurun_out/r2c_pt_rot25.ncu-rep --page raw --csv > gpurun_out/r2c_pt_rot25_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c_pt_rot25.ncu-rep --page source --csv > gpurun_out/r2c_pt_rot25_source.csv 2>/dev/null
tail -3 gpurun_out/r2c_ncu.log

# ncu of the 64-bit row-word STDC chain kernel (planar d=21) and, for comparison, the 32-bit one at planar d=15
python profiles/scripts/planar_split.py > gpurun_out/r2s_planar_split.txt 2>&1; cat gpurun_out/r2s_planar_split.txt
ncu --set full --import-source on --clock-control none -k regex:stdc_fast -c 1 -s 1 -o /tmp/r2s_p21 -f python profiles/scripts/prof_planar21.py 21 3000 > gpurun_out/r2s_ncu.log 2>&1; tail -2 gpurun_out/r2s_ncu.log
ncu -i /tmp/r2s_p21.ncu-rep --page raw --csv > gpurun_out/r2s_p21_raw.csv 2>/dev/null
ncu -i /tmp/r2s_p21.ncu-rep --page source --csv > gpurun_out/r2s_p21_source.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:stdc_fast -c 1 -s 1 -o /tmp/r2s_p15 -f python profiles/scripts/prof_planar21.py 15 3000 > gpurun_out/r2s_ncu2.log 2>&1; tail -2 gpurun_out/r2s_ncu2.log
ncu -i /tmp/r2s_p15.ncu-rep --page raw --csv > gpurun_out/r2s_p15_raw.csv 2>/dev/null

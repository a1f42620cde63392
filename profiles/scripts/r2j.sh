# swap sweep with a fast chunk path, top-rung warps first in the CTA, alpha mask fix
timeout 900 python -m pytest tests/test_gpu_native.py -q -k "ladder or pteq or lane_split" > gpurun_out/r2j_native.log 2>&1; tail -4 gpurun_out/r2j_native.log
for lt in 8; do python profiles/scripts/prof_ladder.py rotated25 400 4736 0.5 $lt; done > gpurun_out/r2j_lt.txt 2>&1
for lt in 8; do python profiles/scripts/prof_ladder.py xzzx21_biased 400 4736 0.5 $lt; done >> gpurun_out/r2j_lt.txt 2>&1
python profiles/scripts/prof_ladder.py xzzx21_alpha 400 4736 0.5 8 >> gpurun_out/r2j_lt.txt 2>&1
python profiles/scripts/prof_ladder.py toric15 400 4736 0.5 8 >> gpurun_out/r2j_lt.txt 2>&1
cat gpurun_out/r2j_lt.txt
ncu --set full --import-source on --clock-control none -k regex:pt_kernel -c 1 -s 1 -o gpurun_out/r2j_pt_xzzx -f python profiles/scripts/prof_ladder.py xzzx21_biased 100 4736 0.5 8 > gpurun_out/r2j_ncu.log 2>&1
ncu -i gpurun_out/r2j_pt_xzzx.ncu-rep --page raw --csv > gpurun_out/r2j_pt_xzzx_raw.csv 2>/dev/null
ncu -i gpurun_out/r2j_pt_xzzx.ncu-rep --page source --csv > gpurun_out/r2j_pt_xzzx_source.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:pt_kernel -c 1 -s 1 -o gpurun_out/r2j_pt_rot -f python profiles/scripts/prof_ladder.py rotated25 100 4736 0.5 8 > gpurun_out/r2j_ncu2.log 2>&1
ncu -i gpurun_out/r2j_pt_rot.ncu-rep --page raw --csv > gpurun_out/r2j_pt_rot_raw.csv 2>/dev/null
ncu -i gpurun_out/r2j_pt_rot.ncu-rep --page source --csv > gpurun_out/r2j_pt_rot_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2j_tests.log 2>&1; tail -8 gpurun_out/r2j_tests.log

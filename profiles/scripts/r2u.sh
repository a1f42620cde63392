# HEAD ncu captures of the tempering kernel (XZZX d=21 biased at 72 registers, rotated d=25 with 4 lanes per top replica)
ncu --set full --import-source on --clock-control none -k regex:pt_kernel -c 1 -s 1 -o /tmp/r2u_x -f python profiles/scripts/prof_ladder.py xzzx21_biased 100 4736 0.5 > gpurun_out/r2u_ncu.log 2>&1
ncu -i /tmp/r2u_x.ncu-rep --page raw --csv > gpurun_out/r2u_pt_xzzx_raw.csv 2>/dev/null
ncu -i /tmp/r2u_x.ncu-rep --page source --csv > gpurun_out/r2u_pt_xzzx_source.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:pt_kernel -c 1 -s 1 -o /tmp/r2u_r -f python profiles/scripts/prof_ladder.py rotated25 100 4736 0.5 > gpurun_out/r2u_ncu2.log 2>&1
ncu -i /tmp/r2u_r.ncu-rep --page raw --csv > gpurun_out/r2u_pt_rot_raw.csv 2>/dev/null
ncu -i /tmp/r2u_r.ncu-rep --page source --csv > gpurun_out/r2u_pt_rot_source.csv 2>/dev/null
tail -2 gpurun_out/r2u_ncu.log gpurun_out/r2u_ncu2.log

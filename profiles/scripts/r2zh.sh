timeout 600 ncu --set full --import-source on --clock-control none -k regex:stdc_pk -c 1 -s 1 -o /tmp/r2zg_pk -f python profiles/scripts/prof_planar21.py 21 3000 > gpurun_out/r2zg_ncu.log 2>&1; tail -2 gpurun_out/r2zg_ncu.log
ncu -i /tmp/r2zg_pk.ncu-rep --page raw --csv > gpurun_out/r2zg_pk_raw.csv 2>/dev/null
ncu -i /tmp/r2zg_pk.ncu-rep --page source --csv > gpurun_out/r2zg_pk_source.csv 2>/dev/null
ls -la gpurun_out/r2zg_pk_*

# ncu of the packed-lattice chain kernel (planar d = 21, 1184 syndromes x 20 000 samples)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:stdc_pk -c 1 -s 1 -o /tmp/r02_pk -f python profiles/scripts/prof_planar21.py 21 20000 > gpurun_out/r02_pk.log 2>&1; tail -2 gpurun_out/r02_pk.log
ncu -i /tmp/r02_pk.ncu-rep --page raw --csv > gpurun_out/r02_ncu_stdc_planar21_packed_raw.csv 2>/dev/null
ncu -i /tmp/r02_pk.ncu-rep --page source --csv > gpurun_out/r02_ncu_stdc_planar21_packed_source.csv 2>/dev/null
python profiles/scripts/prof_planar21.py 21 20000

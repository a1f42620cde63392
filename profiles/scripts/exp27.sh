bash profiles/scripts/exp26.sh
c=xzzx21_biased
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ladder_kernel -s 1 -c 1 -o /tmp/r01_ladder_${c}_v12 -f python profiles/scripts/prof_ladder.py $c 100 > /dev/null 2>&1
ncu -i /tmp/r01_ladder_${c}_v12.ncu-rep --page source --csv > gpurun_out/r01_ncu_ladder_${c}_v12_source.csv 2>/dev/null
ls -la gpurun_out/r01_ncu_ladder_${c}_v12_source.csv

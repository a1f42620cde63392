for c in rotated25 xzzx21_biased; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ladder_kernel -s 1 -c 1 -o /tmp/r01_ladder_${c}_v13 -f python profiles/scripts/prof_ladder.py $c 100 > /dev/null 2>&1
ncu -i /tmp/r01_ladder_${c}_v13.ncu-rep --page source --csv > gpurun_out/r01_ncu_ladder_${c}_v13_source.csv 2>/dev/null
ncu -i /tmp/r01_ladder_${c}_v13.ncu-rep --page raw --csv > gpurun_out/r01_ncu_full_ladder_${c}_v13_raw.csv 2>/dev/null
done
ls -la gpurun_out/*v13*

timeout 900 python bench_configs.py --config all --out gpurun_out/r01_configs_v11.jsonl > gpurun_out/v11_configs.log 2>&1
for c in rotated25 xzzx21_biased; do timeout 600 ncu --set full --import-source on --clock-control none -k regex:ladder_kernel -s 1 -c 1 -o /tmp/r01_ladder_${c}_v11 -f python profiles/scripts/prof_ladder.py $c 100 > /dev/null 2>&1; ncu -i /tmp/r01_ladder_${c}_v11.ncu-rep --page raw --csv > gpurun_out/r01_ncu_full_ladder_${c}_v11_raw.csv 2>/dev/null; done
tail -c 300 gpurun_out/v11_configs.log

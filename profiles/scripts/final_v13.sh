timeout 900 python bench_configs.py --config all --out gpurun_out/r01_configs_v13.jsonl > gpurun_out/v13_configs.log 2>&1
tail -c 600 gpurun_out/v13_configs.log
for c in rotated25 xzzx21_biased xzzx21_alpha; do python profiles/scripts/prof_ladder.py $c 2000; done

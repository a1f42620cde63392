set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/e2_tests.log
timeout 900 python bench_configs.py --config all --out gpurun_out/r01_configs.jsonl > gpurun_out/e2_configs.log 2>&1
python profiles/scripts/prof_ladder.py rotated25 100 > gpurun_out/e2_ladder_plain.log 2>&1
python profiles/scripts/prof_ladder.py xzzx21_biased 100 >> gpurun_out/e2_ladder_plain.log 2>&1
python profiles/scripts/prof_ladder.py xzzx21_alpha 100 >> gpurun_out/e2_ladder_plain.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ladder_kernel -c 1 -o gpurun_out/r01_ladder_rotated25 -f python profiles/scripts/prof_ladder.py rotated25 100 > gpurun_out/e2_ncu_ladder.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ladder_kernel -c 1 -o gpurun_out/r01_ladder_xzzx21_biased -f python profiles/scripts/prof_ladder.py xzzx21_biased 100 >> gpurun_out/e2_ncu_ladder.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:stdc_fast -c 1 -o gpurun_out/r01_stdc_v5 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --samples 2000 > gpurun_out/e2_ncu_stdc.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/e2_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_stdc_v5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/e2_ncu_launches.log 2>&1
cat gpurun_out/e2_tests.log; cat gpurun_out/e2_ladder_plain.log; tail -3 gpurun_out/e2_configs.log | cut -c1-600

timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/e18_tests.log
for c in rotated25 xzzx21_biased xzzx21_alpha; do timeout 120 python profiles/scripts/prof_ladder.py $c 200; done > gpurun_out/e18_ladder.log 2>&1
cat gpurun_out/e18_tests.log gpurun_out/e18_ladder.log

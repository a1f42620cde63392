for c in rotated25 xzzx21_biased xzzx21_alpha; do timeout 120 python profiles/scripts/prof_ladder.py $c 200; done > gpurun_out/e16_ladder.log 2>&1
cat gpurun_out/e16_ladder.log

for rep in 1 2 3; do for c in rotated25 xzzx21_biased xzzx21_alpha; do timeout 120 python profiles/scripts/prof_ladder.py $c 200; done; done > gpurun_out/e19_ladder.log 2>&1
sort gpurun_out/e19_ladder.log

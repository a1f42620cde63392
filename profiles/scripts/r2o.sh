# lanes per top-rung replica after the top-rung rewrites (lt = 4 makes CTAs of <= 448 threads: the 72-register instantiation)
for lt in 4 8 16; do python profiles/scripts/prof_ladder.py rotated25 400 4736 0.5 $lt; done > gpurun_out/r2o_lt.txt 2>&1
for lt in 2 4 8; do python profiles/scripts/prof_ladder.py xzzx21_biased 400 4736 0.5 $lt; done >> gpurun_out/r2o_lt.txt 2>&1
for lt in 4 8; do python profiles/scripts/prof_ladder.py xzzx21_alpha 400 4736 0.5 $lt; done >> gpurun_out/r2o_lt.txt 2>&1
for lt in 4 8; do python profiles/scripts/prof_ladder.py toric15 400 4736 0.5 $lt; done >> gpurun_out/r2o_lt.txt 2>&1
cat gpurun_out/r2o_lt.txt

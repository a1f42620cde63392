# XZZX weighted top rung: per-group iteration counters, one move type per warp round (instead of both branches per iteration)
timeout 900 python -m pytest tests/test_gpu_native.py -q -k "ladder or pteq or lane_split" > gpurun_out/r2za_native.log 2>&1; tail -3 gpurun_out/r2za_native.log
for c in xzzx21_biased xzzx21_alpha rotated25 toric15; do python profiles/scripts/prof_ladder.py $c 400 4736 0.5; done > gpurun_out/r2za_lt.txt 2>&1; cat gpurun_out/r2za_lt.txt

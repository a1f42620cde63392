# tempering kernel: rung thresholds through an opaque shared address (rung step 60 -> 50 instructions)
timeout 900 python -m pytest tests/test_gpu_native.py -q -k "ladder or pteq or lane_split" > gpurun_out/r2w_native.log 2>&1; tail -3 gpurun_out/r2w_native.log
for c in rotated25 xzzx21_biased xzzx21_alpha toric15; do python profiles/scripts/prof_ladder.py $c 400 4736 0.5; done > gpurun_out/r2w_lt.txt 2>&1; cat gpurun_out/r2w_lt.txt

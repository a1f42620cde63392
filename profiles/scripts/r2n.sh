# XZZX top rung software-pipelined (next iteration's draws / decision / descriptor fetched ahead); planar time split by distance
timeout 900 python -m pytest tests/test_gpu_native.py -q -k "ladder or pteq or lane_split" > gpurun_out/r2n_native.log 2>&1; tail -3 gpurun_out/r2n_native.log
for c in xzzx21_biased xzzx21_alpha; do python profiles/scripts/prof_ladder.py $c 400 4736 0.5 8; done > gpurun_out/r2n_lt.txt 2>&1
cat gpurun_out/r2n_lt.txt
python profiles/scripts/planar_split.py > gpurun_out/r2n_planar_split.txt 2>&1; cat gpurun_out/r2n_planar_split.txt

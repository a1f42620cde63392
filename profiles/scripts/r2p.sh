# HEAD validation: full GPU suite, smoke, tempering throughput at the defaults
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2p_tests.log 2>&1; tail -4 gpurun_out/r2p_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in rotated25 xzzx21_biased xzzx21_alpha toric15; do python profiles/scripts/prof_ladder.py $c 400 4736 0.5; done > gpurun_out/r2p_lt.txt 2>&1; cat gpurun_out/r2p_lt.txt
timeout 600 python profiles/scripts/run_config34.py gpu rotated25 4000000 > gpurun_out/r2p_rot.json 2> gpurun_out/r2p_rot.err; cut -c1-400 gpurun_out/r2p_rot.json

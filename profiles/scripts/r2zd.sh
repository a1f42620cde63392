# config 5 points: native decode vs the oracle's MT19937 decode of the same planar syndromes; bench line with the HEAD ncu summary
timeout 1200 python profiles/scripts/run_planar_ci.py 400 96 > gpurun_out/r02_planar_ci.jsonl 2> gpurun_out/r02_planar_ci.err; cat gpurun_out/r02_planar_ci.jsonl; tail -3 gpurun_out/r02_planar_ci.err
python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; cut -c1-200 gpurun_out/r02_bench_full.json

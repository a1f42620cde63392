# last check of the round: GPU suite, smoke, bench at the final HEAD
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; tail -3 gpurun_out/r02_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; cut -c1-160 gpurun_out/r02_bench_full.json

# HEAD check after the container was re-created: full GPU suite, headline bench, smoke
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; tail -3 gpurun_out/r2x_tests.log
python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; cat gpurun_out/r2x_bench.json | head -c 600
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; tail -2 gpurun_out/r2x_smoke.log

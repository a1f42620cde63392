# configs 3 / 4 decoded to convergence with the HEAD tempering kernel (larger caps), final bench lines
cp gpurun_out/r02_rotated25_gpu.npz gpurun_out/r02_rotated25_cap2e6_gpu.npz 2>/dev/null
timeout 600 python profiles/scripts/run_config34.py gpu rotated25 4000000 > gpurun_out/r2l_rot.json 2> gpurun_out/r2l_rot.err; cut -c1-900 gpurun_out/r2l_rot.json; tail -2 gpurun_out/r2l_rot.err
timeout 600 python profiles/scripts/run_config34.py gpu xzzx21_biased 2000000 > gpurun_out/r2l_xb.json 2> gpurun_out/r2l_xb.err; cut -c1-900 gpurun_out/r2l_xb.json; tail -2 gpurun_out/r2l_xb.err
timeout 900 python profiles/scripts/run_config34.py gpu xzzx21_alpha 8000000 > gpurun_out/r2l_xa.json 2> gpurun_out/r2l_xa.err; cut -c1-900 gpurun_out/r2l_xa.json; tail -2 gpurun_out/r2l_xa.err
python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r2l_bench.err; cut -c1-200 gpurun_out/r02_bench_full.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r2l_ref.err; cut -c1-200 gpurun_out/r02_bench_reference_arm.json

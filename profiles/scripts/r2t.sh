# 64-bit row-word STDC kernel: expanded descriptors (ready byte offsets)
timeout 1200 python -m pytest tests/test_gpu_native.py tests/test_gpu_parity.py -q -k "stdc or strc or single_temp or wide or row_word or dedupe or waves or general_noise" > gpurun_out/r2t_tests.log 2>&1; tail -3 gpurun_out/r2t_tests.log
python profiles/scripts/planar_split.py > gpurun_out/r2t_planar_split.txt 2>&1; cat gpurun_out/r2t_planar_split.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

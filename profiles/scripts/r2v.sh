# chain kernel accept path: 20 -> 15 instructions (opaque table base, "state changed" from the accept counter)
timeout 1500 python -m pytest tests/test_gpu_native.py tests/test_gpu_parity.py tests/test_gpu_decoders.py -q -k "stdc or strc or single_temp or wide or row_word or dedupe or waves or general_noise or binomial or chain" > gpurun_out/r2v_tests.log 2>&1; tail -3 gpurun_out/r2v_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2v_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch'], d['chain_stats'])"
python profiles/scripts/planar_split.py > gpurun_out/r2v_planar_split.txt 2>&1; cat gpurun_out/r2v_planar_split.txt

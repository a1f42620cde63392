python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stdc or dedupe or strc" 2>&1 | tail -2 > gpurun_out/e9_tests.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/e9_$name.json 2>gpurun_out/e9_$name.err; python -c "
import json
d=json.load(open('gpurun_out/e9_$name.json')); print('$name', '%.3e'%d['value'], '%.1f'%d['ms_per_step'], '%.1f'%d['roofline']['kernel_ms_per_launch'], '%.3e'%d['e2e']['value'])"; }
run nosync QECMC_DEBUG_SYNC_CALLS=2000000000
run sync1024 QECMC_DEBUG_SYNC_CALLS=1024
run sync128 QECMC_DEBUG_SYNC_CALLS=128
run sync16 QECMC_DEBUG_SYNC_CALLS=16
run t640_sync1024 QECMC_DEBUG_T=640 QECMC_DEBUG_SYNC_CALLS=1024
run t640_sync64 QECMC_DEBUG_T=640 QECMC_DEBUG_SYNC_CALLS=64
run t640_nosync QECMC_DEBUG_T=640 QECMC_DEBUG_SYNC_CALLS=2000000000
run t128_sync128 QECMC_DEBUG_T=128 QECMC_DEBUG_SYNC_CALLS=128
cat gpurun_out/e9_tests.log

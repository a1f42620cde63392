python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stdc or dedupe" 2>&1 | tail -4 > gpurun_out/e6_tests.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/e6_base.json 2>gpurun_out/e6_base.err
QECMC_LIB=$PWD/mcmc-qec-toric-rl_b200/csrc/_variants/libqecmc_t512.so $B --syndromes 222 > gpurun_out/e6_t512.json 2>gpurun_out/e6_t512.err
QECMC_LIB=$PWD/mcmc-qec-toric-rl_b200/csrc/_variants/libqecmc_t512.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stdc or dedupe" 2>&1 | tail -4 >> gpurun_out/e6_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/e6_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for f in gpurun_out/e6_base.json gpurun_out/e6_t512.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', '%.3e'%d['value'], '%.1f'%d['ms_per_step'], '%.1f'%d['roofline']['kernel_ms_per_launch'], '%.3e'%d['e2e']['value'], d['config']['syndromes_per_step_per_gpu'])"; done
cat gpurun_out/e6_tests.log
grep -v "^==" gpurun_out/e6_launches.csv | awk -F'","' '{print $5, $NF}' | tail -6

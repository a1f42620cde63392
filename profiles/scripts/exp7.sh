python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stdc or dedupe or strc" 2>&1 | tail -4 > gpurun_out/e7_tests.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/e7_base.json 2>gpurun_out/e7_base.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 8 --csv --log-file gpurun_out/e7_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --section SchedulerStats --section LaunchStats --section Occupancy --import-source on --clock-control none -k regex:log_dedupe -c 1 -o gpurun_out/r01_dedupe_v3 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/e7_ncu.log 2>&1
for f in gpurun_out/e7_base.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', '%.3e'%d['value'], '%.1f'%d['ms_per_step'], '%.1f'%d['roofline']['kernel_ms_per_launch'], '%.3e'%d['e2e']['value'], d['config']['syndromes_per_step_per_gpu'])"; done
cat gpurun_out/e7_tests.log
grep -v "^==" gpurun_out/e7_launches.csv | awk -F'","' '{print $5, $NF}' | tail -4

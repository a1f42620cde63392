set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/e1_tests.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/e1_default.json 2> gpurun_out/e1_default.err
QECMC_L2_FETCH=0 $B > gpurun_out/e1_l2f0.json 2>/dev/null
QECMC_L2_FETCH=64 $B > gpurun_out/e1_l2f64.json 2>/dev/null
QECMC_L2_FETCH=128 $B > gpurun_out/e1_l2f128.json 2>/dev/null
QECMC_DEBUG_INSERT_MODE=3 $B > gpurun_out/e1_noset.json 2>/dev/null
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_atom.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --replay-mode application --metrics $M -k regex:stdc_fast --clock-control none -c 1 --csv --log-file gpurun_out/e1_ncu_full_size.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/e1_ncu.log 2>&1
QECMC_L2_FETCH=128 timeout 900 ncu --replay-mode application --metrics $M -k regex:stdc_fast --clock-control none -c 1 --csv --log-file gpurun_out/e1_ncu_full_size_l2f128.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/e1_ncu2.log 2>&1
for f in gpurun_out/e1_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'])"; done
cat gpurun_out/e1_tests.log

# A/B of a bucket-dedupe change: the bucket-log parity tests, then the headline bench
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dedupe or bucket or golden or conv" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('%.4e'%d['value'], 'ms/step %.2f'%d['ms_per_step'], 'chain %.2f'%r['kernel_ms_per_launch'], 'other %.2f'%r['other_kernels_ms_per_step'])"

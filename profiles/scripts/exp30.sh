timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dedupe or bucket or golden or conv" 2>&1 | tail -3
timeout 600 python profiles/scripts/prof_modes.py 2>&1 | tail -12

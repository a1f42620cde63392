# table-driven step of the two-layer codes in the ladder kernel: identical native results, replay parity, timing
QECMC_LIB=/root/repo/build/libqecmc_prev.so python profiles/scripts/native_dump.py /tmp/prev.npz 2>&1 | tail -1
python profiles/scripts/native_dump.py /tmp/new.npz 2>&1 | tail -1
python - <<'P'
import numpy as np
a=np.load('/tmp/prev.npz'); b=np.load('/tmp/new.npz')
print('arrays', len(a.files), 'differing', [k for k in a.files if not np.array_equal(a[k],b[k])])
P
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
QECMC_LIB=/root/repo/build/libqecmc_prev.so python profiles/scripts/prof_ladder.py toric15 200 9472
python profiles/scripts/prof_ladder.py toric15 200 9472
python profiles/scripts/prof_ptdc.py

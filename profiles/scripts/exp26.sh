# new swap sweep: native results must be identical to the previous build's (copy the previous libqecmc.so to build/libqecmc_prev.so first), tests green, timings
QECMC_LIB=/root/repo/build/libqecmc_prev.so python profiles/scripts/native_dump.py /tmp/prev.npz 2>&1 | tail -1
python profiles/scripts/native_dump.py /tmp/new.npz 2>&1 | tail -1
python - <<'P'
import numpy as np
a=np.load('/tmp/prev.npz'); b=np.load('/tmp/new.npz')
bad=[k for k in a.files if not np.array_equal(a[k],b[k])]
print('arrays', len(a.files), 'differing', bad)
for k in a.files:
    if k.endswith('_tops0'): print(k, int(a[k].sum()), int(b[k].sum()))
P
bash profiles/scripts/exp18.sh

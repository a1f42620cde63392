import time, numpy as np, sys
sys.path.insert(0, '/root/repo')
from mcmc_qec_toric_rl_b200 import _lib
ctx = _lib.Context(0)
rng = np.random.default_rng(1)
L, S = 5, 100
qm = ((rng.random((S, 2 * L * L)) < 0.1) * rng.integers(1, 4, (S, 2 * L * L))).astype(np.uint8)
for name, fn in (("STDC", ctx.stdc), ("STRC", ctx.strc), ("STDC", ctx.stdc)):
    for i in range(3):
        t = time.perf_counter(); out = fn(_lib.TORIC, _lib.TORIC, L, qm, 0.1, 0.25, 10, 3125, seed=1 + i); dt = time.perf_counter() - t
        st = out[1]
        print(name, i, "wall %.2f ms" % (dt * 1e3), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}, flush=True)
ctx.debug_set("insert_mode", 4)
for i in range(2):
    t = time.perf_counter(); out = ctx.stdc(_lib.TORIC, _lib.TORIC, L, qm, 0.1, 0.25, 10, 3125, seed=1 + i); dt = time.perf_counter() - t
    print("STDC per-chain logs", i, "wall %.2f ms" % (dt * 1e3), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in out[1].items()}, flush=True)

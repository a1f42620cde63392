"""Where a planar STDC call spends its time by code distance: chain kernel vs everything else (dedupe), GPU-filling batches."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcmc_qec_toric_rl_b200 import _lib
ctx = _lib.Context(0)
info = ctx.device_info()
ctx.set_table_budget(int(info["free_mem"] * 0.8))
rng = np.random.default_rng(5)
for d in (11, 15, 17, 21):
    S, droplets, steps = 2368, 16, d ** 4
    q = ((rng.random((S, 2, d, d)) < 0.15) * rng.integers(1, 4, (S, 2, d, d))).astype(np.uint8)
    q[:, 1, -1, :] = 0
    q[:, 1, :, -1] = 0
    qm = np.ascontiguousarray(q.reshape(S, -1))
    ctx.stdc(_lib.PLANAR, _lib.PLANAR, d, qm, 0.15, 0.25, droplets, steps, seed=1)
    out, st = ctx.stdc(_lib.PLANAR, _lib.PLANAR, d, qm, 0.15, 0.25, droplets, steps, seed=2)
    print("planar d=%d: total %.1f ms, chain kernel %.1f ms (%.0f %%), waves %d, table_slots %d, steps/s %.3g, launches %d" % (
        d, st["total_ms"], st["chain_kernel_ms"], 100 * st["chain_kernel_ms"] / st["total_ms"], st["waves"], st["table_slots"],
        st["metropolis_steps"] / (st["total_ms"] * 1e-3), st["kernel_launches"]))

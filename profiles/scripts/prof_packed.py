"""Planar STDC at the sweep's sizes: packed-lattice chain kernel (copies of the hot tables 1 / 2 / 4) against the 64-bit
row-word kernel ("packed" = 0).  One GPU-filling call each."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcmc_qec_toric_rl_b200 import _lib
ctx = _lib.Context(0)
rng = np.random.default_rng(5)
droplets = 16
only = [int(a) for a in sys.argv[1:]]
STEPS = int(os.environ.get("PK_STEPS", "40000"))
for d, S, steps in [(21, 2960, STEPS), (17, 4736, STEPS), (19, 4144, STEPS)]:
    if only and d not in only:
        continue
    q = ((rng.random((S, 2, d, d)) < 0.15) * rng.integers(1, 4, (S, 2, d, d))).astype(np.uint8)
    q[:, 1, -1, :] = 0
    q[:, 1, :, -1] = 0
    qm = np.ascontiguousarray(q.reshape(S, -1))
    ref = None
    for pk in (0, 2, 4, 8):
        ctx.debug_set("packed", pk)
        ctx.stdc(_lib.PLANAR, _lib.PLANAR, d, qm[:64], 0.15, 0.25, droplets, 50, seed=1)
        ctx.stdc(_lib.PLANAR, _lib.PLANAR, d, qm, 0.15, 0.25, droplets, 200, seed=2)      # full-batch warm-up
        out, st = ctx.stdc(_lib.PLANAR, _lib.PLANAR, d, qm, 0.15, 0.25, droplets, steps, seed=2)[:2]
        if ref is None:
            ref = out
        print("planar d=%d S=%d packed=%d: chain kernel %.2f ms (%.3g steps/s), whole call %.2f ms (%.3g steps/s), same result %s" % (
            d, S, pk, st["chain_kernel_ms"], st["metropolis_steps"] / (st["chain_kernel_ms"] * 1e-3), st["total_ms"],
            st["metropolis_steps"] / (st["total_ms"] * 1e-3), np.array_equal(out, ref)), flush=True)
ctx.debug_set("packed", -1)

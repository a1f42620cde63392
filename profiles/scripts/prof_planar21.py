"""One GPU-filling planar d=21 (or argv[1]) STDC call for ncu captures of the 64-bit row-word chain kernel."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcmc_qec_toric_rl_b200 import _lib
d = int(sys.argv[1]) if len(sys.argv) > 1 else 21
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
ctx = _lib.Context(0)
rng = np.random.default_rng(5)
S, droplets = 1184, 16
q = ((rng.random((S, 2, d, d)) < 0.15) * rng.integers(1, 4, (S, 2, d, d))).astype(np.uint8)
q[:, 1, -1, :] = 0
q[:, 1, :, -1] = 0
qm = np.ascontiguousarray(q.reshape(S, -1))
ctx.stdc(_lib.PLANAR, _lib.PLANAR, d, qm[:64], 0.15, 0.25, droplets, 50, seed=1)
out, st = ctx.stdc(_lib.PLANAR, _lib.PLANAR, d, qm, 0.15, 0.25, droplets, steps, seed=2)
print("planar d=%d steps %d: chain kernel %.2f ms, %.3g steps/s in the kernel" % (d, steps, st["chain_kernel_ms"], st["metropolis_steps"] / (st["chain_kernel_ms"] * 1e-3)))

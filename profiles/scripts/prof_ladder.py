"""Small driver for ncu captures of the tempering-ladder kernel: one GPU-filling PTEQ launch of a given config."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcmc_qec_toric_rl_b200 import _lib

which = sys.argv[1] if len(sys.argv) > 1 else "rotated25"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
S = int(sys.argv[3]) if len(sys.argv) > 3 else 4736
p_logical = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
ctx = _lib.Context(0)
if len(sys.argv) > 5:
    ctx.debug_set("pt_lt", int(sys.argv[5]))          # lanes per top-rung replica (rung-major kernel)
if len(sys.argv) > 6:
    ctx.debug_set("ladder_kernel", int(sys.argv[6]))  # 1: the warp-per-ladder kernel
rng = np.random.default_rng(3)
if which == "rotated25":
    g, L, kind, bottom, b = _lib.ROTATED, 25, _lib.LADDER_DEPOLARIZING, 0.15, 0.0
elif which == "toric15":
    g, L, kind, bottom, b = _lib.TORIC, 15, _lib.LADDER_DEPOLARIZING, 0.15, 0.0
elif which == "xzzx21_biased":
    g, L, kind, bottom, b = _lib.XZZX, 21, _lib.LADDER_BIASED, 0.15, 100.0
else:
    g, L, kind, bottom, b = _lib.XZZX, 21, _lib.LADDER_ALPHA, 0.17474, 0.6447
ns = 2 * L * L if g == _lib.TORIC else L * L
q = ((rng.random((S, ns)) < 0.15) * rng.integers(1, 4, (S, ns))).astype(np.uint8)
ctx.pteq(g, L, kind, q[:64], bottom, param_b=b, steps=3, conv=False, seed=1, p_logical=p_logical)      # module load, allocations
pct, info = ctx.pteq(g, L, kind, q, bottom, param_b=b, steps=steps, conv=False, seed=11, p_logical=p_logical)
st = info["stats"]
print(which, "lt", sys.argv[5] if len(sys.argv) > 5 else "default", "S", S, "steps", steps, "p_logical", p_logical, "kernel_ms", st["chain_kernel_ms"], "metropolis steps/s", st["metropolis_steps"] / (st["chain_kernel_ms"] * 1e-3))

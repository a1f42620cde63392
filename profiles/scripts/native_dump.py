"""Dump native-mode ladder results (fixed seeds) to an .npz: two builds of libqecmc.so that claim the same native
decisions must produce identical files (QECMC_LIB selects the build)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcmc_qec_toric_rl_b200 import _lib

out = sys.argv[1]
ctx = _lib.Context(0)
rng = np.random.default_rng(5)
res = {}
cases = [("rot9", _lib.ROTATED, 9, _lib.LADDER_DEPOLARIZING, 0.15, 0.0, None),
         ("rot25", _lib.ROTATED, 25, _lib.LADDER_DEPOLARIZING, 0.15, 0.0, None),
         ("rot5nc3", _lib.ROTATED, 5, _lib.LADDER_DEPOLARIZING, 0.1, 0.0, 3),
         ("rot5nc2", _lib.ROTATED, 5, _lib.LADDER_DEPOLARIZING, 0.1, 0.0, 2),
         ("rot5nc1", _lib.ROTATED, 5, _lib.LADDER_DEPOLARIZING, 0.1, 0.0, 1),
         ("tor5nc2", _lib.TORIC, 5, _lib.LADDER_DEPOLARIZING, 0.1, 0.0, 2),
         ("tor7", _lib.TORIC, 7, _lib.LADDER_DEPOLARIZING, 0.12, 0.0, None),
         ("pla9", _lib.PLANAR, 9, _lib.LADDER_DEPOLARIZING, 0.12, 0.0, None),
         ("xzzx11b", _lib.XZZX, 11, _lib.LADDER_BIASED, 0.15, 30.0, None),
         ("xzzx21b", _lib.XZZX, 21, _lib.LADDER_BIASED, 0.15, 100.0, None),
         ("xzzx9a", _lib.XZZX, 9, _lib.LADDER_ALPHA, 0.17474, 0.6447, None)]
for name, g, L, kind, bottom, b, nc in cases:
    ns = 2 * L * L if g in (_lib.TORIC, _lib.PLANAR) else L * L
    S = 300
    q = ((rng.random((S, ns)) < 0.12) * rng.integers(1, 4, (S, ns))).astype(np.uint8)
    if g == _lib.PLANAR:
        q = q.reshape(S, 2, L, L); q[:, 1, L - 1, :] = 0; q[:, 1, :, L - 1] = 0; q = q.reshape(S, ns)
    kw = dict(param_b=b, steps=400, conv=False, seed=77, p_logical=0.5)
    if nc: kw["Nc"] = nc
    pct, info = ctx.pteq(g, L, kind, q, bottom, **kw)
    res[name + "_pct"] = np.asarray(pct)
    for k, v in info.items():
        if isinstance(v, np.ndarray): res[name + "_" + k] = v
    res[name + "_acc"] = np.array([info["stats"]["accepted"] if "accepted" in info["stats"] else 0])
np.savez(out, **res)
print("wrote", out, len(res), "arrays")

"""PTDC / PTRC at a GPU-filling size: toric d=9, Nc=9, 4 droplets per class, 2000 ladder steps, 300 syndromes."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcmc_qec_toric_rl_b200 import _lib
ctx = _lib.Context(0)
g, L, S, Nc, dr, steps = _lib.TORIC, 9, 300, 9, 4, 2000
qm, truth = ctx.generate_errors(g, L, S, p_error=0.1, seed=3)
for name, f in (("PTDC", lambda: ctx.ptdc(g, L, qm, 0.1, 0.25, dr, Nc, steps, seed=5)),
                ("PTRC", lambda: ctx.ptrc(g, L, qm, 0.1, 0.25, dr, Nc, steps, seed=5))):
    f()
    t = time.perf_counter(); out = f(); dt = time.perf_counter() - t
    st = out[1]
    fail = float((np.asarray(out[0]).argmax(1) != truth).mean())
    print("%s wall %.1f ms  ladder kernel %.1f ms  total %.1f ms  Metropolis steps/s %.3e  launches %d  failure rate %.3f" % (
        name, dt * 1e3, st["chain_kernel_ms"], st["total_ms"], st["metropolis_steps"] / dt, st["kernel_launches"], fail))

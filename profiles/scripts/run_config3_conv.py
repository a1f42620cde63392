"""BASELINE config 3 with the reference's convergence criterion left to run: rotated d=25, depolarizing p=0.15, PTEQ
(Nc=25, iters=10, SEQ=2, TOPS=10, tops_burn=2, eps=0.1), step cap from argv; syndromes from the device workload kernels."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcmc_qec_toric_rl_b200 import _lib
cap = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4736
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
ctx = _lib.Context(0)
g, L = _lib.ROTATED, 25
qm, truth = ctx.generate_errors(g, L, S, p_xyz=(p / 3, p / 3, p / 3), seed=3)
ctx.pteq(g, L, _lib.LADDER_DEPOLARIZING, qm, p, steps=50, conv=True, seed=1)
t = time.perf_counter()
pct, info = ctx.pteq(g, L, _lib.LADDER_DEPOLARIZING, qm, p, steps=cap, conv=True, seed=11)
dt = time.perf_counter() - t
st = info["stats"]
out = {"config": "rotated d=25 depolarizing p=%.2f PTEQ with the error-based convergence criterion, step cap %d" % (p, cap), "ladders": S,
       "seconds": dt, "syndromes_per_s": S / dt, "metropolis_steps": int(st["metropolis_steps"]), "steps_per_s": st["metropolis_steps"] / dt,
       "kernel_ms": st["chain_kernel_ms"], "converged": float(info["converged"].mean()), "mean_ladder_steps": float(info["steps"].mean()),
       "max_ladder_steps": int(info["steps"].max()), "tops0_mean": float(info["tops0"].mean()),
       "logical_failure_rate": float((pct.argmax(1) != truth).mean())}
print(json.dumps(out))

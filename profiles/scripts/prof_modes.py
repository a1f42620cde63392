"""Headline-size timing of the other entry points that share the STDC chain kernel: STRC, single_temp, conv_mult early
stop, planar per-class inits.  One GPU-filling call each (148 syndromes of toric d=15 unless stated)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcmc_qec_toric_rl_b200 import _lib

ctx = _lib.Context(0)
rng = np.random.default_rng(2)
L, S, dr, steps = 15, 148, 64, 15 ** 4
q = ((rng.random((S, 2 * L * L)) < 0.15) * rng.integers(1, 4, (S, 2 * L * L))).astype(np.uint8)
def run(name, f):
    f()   # warm-up: module load, allocations
    t = time.perf_counter(); out = f(); dt = time.perf_counter() - t
    st = out[1]
    print("%-28s wall %.1f ms  chain kernel %.1f ms  total %.1f ms  steps/s %.3e  table_slots %d  launches %d" % (
        name, dt * 1e3, st["chain_kernel_ms"], st["total_ms"], st["metropolis_steps"] / dt, st["table_slots"], st["kernel_launches"]))
run("STDC toric15", lambda: ctx.stdc(_lib.TORIC, _lib.TORIC, L, q, 0.15, 0.25, dr, steps, seed=1))
run("STRC toric15", lambda: ctx.strc(_lib.TORIC, _lib.TORIC, L, q, 0.15, 0.25, dr, steps, seed=1))
run("STRC toric15 + hist", lambda: ctx.strc(_lib.TORIC, _lib.TORIC, L, q, 0.15, 0.25, dr, steps, seed=1, want_hist=True)[:2])
run("STDC toric15 + N_hist", lambda: ctx.stdc(_lib.TORIC, _lib.TORIC, L, q, 0.15, 0.25, dr, steps, seed=1, want_hist=True)[:2])
run("STDC toric15 conv_mult=2", lambda: ctx.stdc(_lib.TORIC, _lib.TORIC, L, q, 0.15, 0.25, dr, steps, seed=1, conv_mult=2.0))
q10 = np.concatenate([q] * 7)[:943]   # 943 syndromes x 16 classes x 10 chains: 147.9 CTAs of 1020 threads (102 tables each)
run("STDC toric15 droplets=10", lambda: ctx.stdc(_lib.TORIC, _lib.TORIC, L, q10, 0.15, 0.25, 10, steps, seed=1))
ctx.debug_set("insert_mode", 4)
run("  same, per-chain logs", lambda: ctx.stdc(_lib.TORIC, _lib.TORIC, L, q10, 0.15, 0.25, 10, steps, seed=1))
ctx.debug_set("insert_mode", -1)
def st_single():
    out = ctx.single_temp(_lib.TORIC, _lib.TORIC, L, np.repeat(q, 8, 0), 0.15, steps)
    return out if isinstance(out, tuple) else (out, {})
try:
    run("single_temp toric15", st_single)
except Exception as e:
    print("single_temp:", e)
def st_single_full():
    # 9472 syndromes x 16 classes = 151 552 chains = 1024 per SM: the size at which single_temp fills the GPU
    out = ctx.single_temp(_lib.TORIC, _lib.TORIC, L, np.repeat(q, 64, 0), 0.15, 4000)
    return out if isinstance(out, tuple) else (out, {})
try:
    run("single_temp toric15 x64", st_single_full)
except Exception as e:
    print("single_temp x64:", e)

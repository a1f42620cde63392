"""BASELINE config 5 points decoded natively and by the oracle on MT19937 streams, the same syndromes: logical failure counts
with the binomial sigma of their difference, for planar d in {7, 11, 15} and p across the sweep's range.
Usage: run_planar_ci.py [syndromes_small_d] [syndromes_d15]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from mcmc_qec_toric_rl_b200 import _lib  # noqa: E402

S_small = int(sys.argv[1]) if len(sys.argv) > 1 else 400
S_15 = int(sys.argv[2]) if len(sys.argv) > 2 else 96
g, droplets = O.PLANAR, 16
threads = len(os.sched_getaffinity(0))
ctx = _lib.default_context(0)
rng = np.random.default_rng(20255)
rows = []
for L, p, S in [(7, 0.10, S_small), (7, 0.16, S_small), (7, 0.20, S_small), (11, 0.10, S_small), (11, 0.16, S_small), (11, 0.20, S_small),
                (15, 0.16, S_15)]:
    qs, truth = [], []
    for _ in range(S):
        q = ((rng.random((2, L, L)) < p) * rng.integers(1, 4, (2, L, L))).astype(np.uint8)
        q[1, -1, :] = 0
        q[1, :, -1] = 0
        truth.append(O.eq_class(g, L, q))
        q2, _ = O.apply_random_logical(g, L, q, O.Stream.mt(int(rng.integers(1 << 30))))
        qs.append(np.asarray(q2, np.uint8).reshape(-1))
    qm, truth = np.stack(qs), np.array(truth)
    steps = L ** 4
    gpu, st = ctx.stdc(g, g, L, qm, p, 0.25, droplets, steps, seed=5)
    t0 = time.time()
    ref = O.stdc_batch(g, g, L, qm, p, 0.25, droplets, steps, seed=17, threads=threads)
    dt = time.time() - t0
    f_gpu, f_ref = int((gpu.argmax(1) != truth).sum()), int((ref.argmax(1) != truth).sum())
    ph = (f_ref + f_gpu) / (2.0 * S)
    sig = float(np.sqrt(max(2 * ph * (1 - ph) * S, 1e-12)))
    row = dict(d=L, p=p, syndromes=S, droplets=droplets, samples=steps, native_failures=f_gpu, oracle_failures=f_ref,
               difference=f_gpu - f_ref, sigma_of_difference=round(sig, 2), same_choice=float((gpu.argmax(1) == ref.argmax(1)).mean()),
               mean_abs_diff_points=float(np.abs(gpu - ref).mean()), oracle_seconds=round(dt, 1), oracle_threads=threads)
    rows.append(row)
    print(json.dumps(row), flush=True)

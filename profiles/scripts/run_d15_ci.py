"""Native decode vs the oracle's MT19937 decode of the same toric d=15 p=0.15 syndromes (BASELINE config 1 sizes):
logical failure counts with the binomial sigma of their difference.  Usage: run_d15_ci.py [syndromes] [droplets] [oracle_samples]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from mcmc_qec_toric_rl_b200 import _lib  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
droplets = int(sys.argv[2]) if len(sys.argv) > 2 else 16
oracle_samples = int(sys.argv[3]) if len(sys.argv) > 3 else 15 ** 4
g, L, p = O.TORIC, 15, 0.15
rng = np.random.default_rng(20251)
qs, truth = [], []
for _ in range(S):
    q = ((rng.random((2, L, L)) < p) * rng.integers(1, 4, (2, L, L))).astype(np.uint8)
    truth.append(O.eq_class(g, L, q))
    q2, _ = O.apply_random_logical(g, L, q, O.Stream.mt(int(rng.integers(1 << 30))))
    qs.append(np.asarray(q2, np.uint8).reshape(-1))
qm, truth = np.stack(qs), np.array(truth)
ctx = _lib.default_context(0)
out = {"syndromes": S, "droplets": droplets, "p": p, "native": {}, "oracle": {}}
for samples in (5000, 15000, 15 ** 4):
    gpu, st = ctx.stdc(g, g, L, qm, p, 0.25, droplets, samples, seed=5)
    out["native"][samples] = int((gpu.argmax(1) != truth).sum())
    if samples == oracle_samples:
        t0 = time.time()
        ref = O.stdc_batch(g, g, L, qm, p, 0.25, droplets, samples, seed=17, threads=len(os.sched_getaffinity(0)))
        f_ref, f_gpu = int((ref.argmax(1) != truth).sum()), out["native"][samples]
        ph = (f_ref + f_gpu) / (2.0 * S)
        out["oracle"][samples] = dict(failures=f_ref, seconds=round(time.time() - t0, 1), same_choice=float((gpu.argmax(1) == ref.argmax(1)).mean()),
                                      mean_abs_diff_points=float(np.abs(gpu - ref).mean()),
                                      sigma_of_difference=float(np.sqrt(2 * ph * (1 - ph) * S)), difference=f_gpu - f_ref)
print(json.dumps(out))

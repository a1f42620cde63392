"""BASELINE config 1 end to end: 10 000 toric d=15 syndromes at p=0.15 generated, labelled, hidden, decoded (STDC, 16 classes
x 64 chains x 15^4 samples x 5 steps) and scored on one GPU through generate_data.generate_batch (nothing but the results
leaves the device)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mcmc_qec_toric_rl_b200 import generate_data as G, _lib
import torch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
batch = 148
params = dict(code='toric', method='STDC', size=15, noise='depolarizing', p_error=0.15, p_sampling=0.25, droplets=64, steps=15**4,
              mwpm_init=False)
ctx = _lib.default_context(0)
ctx.set_table_budget(int(ctx.device_info()["free_mem"] * 0.85))
G.generate_batch(params, batch, seed=1)          # warm-up: allocations, module load
torch.cuda.synchronize()
t0 = time.perf_counter()
done = fails = 0
while done < N:
    S = min(batch, N - done)
    res = G.generate_batch(params, S, seed=1000 + done)
    fails += res['failures']
    done += S
torch.cuda.synchronize()
dt = time.perf_counter() - t0
steps = done * 16 * 64 * 15**4 * 5
print(json.dumps({"config": "toric d=15, depolarizing p=0.15, STDC 16 classes x 64 chains x 15^4 samples x 5 steps, p_sampling=0.25",
                  "syndromes": done, "seconds": dt, "syndromes_per_s": done / dt, "metropolis_steps_per_s": steps / dt,
                  "logical_failures": fails, "logical_failure_rate": fails / done,
                  "binomial_sigma": float(np.sqrt(max(fails, 1) * (1 - fails / done)) / done), "device": ctx.device_info()["name"]}))

# MWPM start states with the folded matching: GPU test + workload timings (planar STDC, 1184 syndromes per call)
timeout 600 python -m pytest tests/test_mwpm.py -q -m gpu > gpurun_out/r2zs_mwpm.log 2>&1; tail -2 gpurun_out/r2zs_mwpm.log
python - > gpurun_out/r02_mwpm_workload.txt 2>&1 <<'P'
import time, numpy as np, os
from mcmc_qec_toric_rl_b200 import generate_data as G
for d in (7, 11, 15, 21):
    params = dict(code="planar", method="STDC", size=d, noise="depolarizing", p_error=0.15, p_sampling=0.25, droplets=16, steps=d ** 4, mwpm_init=True)
    S = 1184
    G.generate_batch(params, 148, seed=1)
    for init in (True, False):
        t = time.perf_counter(); r = G.generate_batch(dict(params, mwpm_init=init), S, seed=3); dt = time.perf_counter() - t
        print("planar d=%d p=0.15 STDC 16 droplets, %d syndromes, mwpm_init=%s: %.2f s, %d failures (host cores %d)" % (d, S, init, dt, r["failures"], os.cpu_count()), flush=True)
P
cat gpurun_out/r02_mwpm_workload.txt

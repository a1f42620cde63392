# MWPM start states on the GPU path; single_temp at a GPU-filling batch
timeout 600 python -m pytest tests/test_mwpm.py -q -m gpu > gpurun_out/r2zb_mwpm.log 2>&1; tail -3 gpurun_out/r2zb_mwpm.log
python - > gpurun_out/r2zb_mwpm_workload.txt 2>&1 <<'P'
import time, numpy as np, os
from mcmc_qec_toric_rl_b200 import generate_data as G
for d in (7, 11, 15):
    params = dict(code="planar", method="STDC", size=d, noise="depolarizing", p_error=0.15, p_sampling=0.25, droplets=16, steps=d ** 4, mwpm_init=True)
    S = 1184
    G.generate_batch(params, 148, seed=1)
    for init in (True, False):
        t = time.perf_counter(); r = G.generate_batch(dict(params, mwpm_init=init), S, seed=3); dt = time.perf_counter() - t
        print("planar d=%d p=0.15 STDC 16 droplets, %d syndromes, mwpm_init=%s: %.2f s, %d failures (host cores %d)" % (d, S, init, dt, r["failures"], os.cpu_count()), flush=True)
P
cat gpurun_out/r2zb_mwpm_workload.txt
python profiles/scripts/prof_modes.py > gpurun_out/r2zb_modes.txt 2>&1; cat gpurun_out/r2zb_modes.txt

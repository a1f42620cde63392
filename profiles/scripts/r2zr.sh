timeout 900 python -m pytest tests/test_workload.py -q -x -m gpu > gpurun_out/r2zr_tests.log 2>&1; tail -12 gpurun_out/r2zr_tests.log
python - <<'P'
import time, tempfile, os
from mcmc_qec_toric_rl_b200 import generate_data as G
params = dict(code='planar', method='STDC', size=7, noise='depolarizing', p_error=0.15, p_sampling=0.25, droplets=16, steps=7 ** 4, mwpm_init=False)
d = tempfile.mkdtemp()
for b in (256, None):
    t = time.perf_counter(); f, n = G.generate(os.path.join(d, 'x.xz'), params, nbr_datapoints=60000, batch=b, seed=1, verbose=False); dt = time.perf_counter() - t
    print("generate(): planar d=7 STDC, 60000 syndromes, batch=%s: %.1f s, %d failures" % (b, dt, f), flush=True)
P

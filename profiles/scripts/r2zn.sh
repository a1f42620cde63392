# final lines at HEAD: smoke (incl. the packed-lattice kernel), bench own arm + reference arm, toric d = 17-21 on the packed kernel
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zn_smoke.log 2>&1; tail -2 gpurun_out/r2zn_smoke.log
python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; cut -c1-200 gpurun_out/r02_bench_full.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; cut -c1-200 gpurun_out/r02_bench_reference_arm.json
python - > gpurun_out/r2zn_toric.txt 2>&1 <<'P'
import numpy as np
from mcmc_qec_toric_rl_b200 import _lib
ctx = _lib.Context(0)
rng = np.random.default_rng(3)
for d, S, droplets, steps in [(17, 592, 16, 40000), (19, 592, 16, 40000), (21, 592, 16, 40000)]:
    qm = ((rng.random((S, 2 * d * d)) < 0.12) * rng.integers(1, 4, (S, 2 * d * d))).astype(np.uint8)
    for pk in (0, -1):
        ctx.debug_set("packed", pk)
        ctx.stdc(_lib.TORIC, _lib.TORIC, d, qm, 0.12, 0.25, droplets, 200, seed=2)
        out, st = ctx.stdc(_lib.TORIC, _lib.TORIC, d, qm, 0.12, 0.25, droplets, steps, seed=2)[:2]
        print("toric d=%d, %d syndromes x 16 classes x %d chains x %d samples, %s: chain kernel %.1f ms (%.3g steps/s), whole call %.1f ms (%.3g steps/s)" % (
            d, S, droplets, steps, "64-bit row words" if pk == 0 else "packed lattice", st["chain_kernel_ms"],
            st["metropolis_steps"] / st["chain_kernel_ms"] / 1e-3, st["total_ms"], st["metropolis_steps"] / st["total_ms"] / 1e-3), flush=True)
P
cat gpurun_out/r2zn_toric.txt

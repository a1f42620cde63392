"""Planar STDC failure rate against d across the 32-bit / 64-bit row-word boundary (d = 16 | 17) at fixed p: a smooth trend
says the two word widths decode alike."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_configs as B
from mcmc_qec_toric_rl_b200 import _lib
from oracle import oracle as O
ctx = _lib.Context(0)
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.15
r = B.run_planar_sweep(ctx, O, ps=(p,), ds=(11, 13, 15, 16, 17, 19, 21))
for q in r["points"]:
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in q.items()}))

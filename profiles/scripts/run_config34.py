"""BASELINE configs 3 and 4 decoded to convergence by the reference's own criterion (decoders.py:57-105,
decoders_biasednoise.py:28-237), and the same syndromes decoded by the CPU oracle on the reference's MT19937 streams.

    run_config34.py gpu    <config> <cap> [S]        -> gpurun_out/r02_<config>_gpu.npz + one JSON line   (needs the GPU)
    run_config34.py oracle <config> <cap> <n> [thr]  -> gpurun_out/r02_<config>_oracle.npz + one JSON line (CPU only)
    run_config34.py merge  <config>                  -> one JSON line comparing the two on the syndromes both decoded

config: rotated25 (PTEQ, depolarizing p=0.15), xzzx21_biased (PTEQ_biased eta=100 p=0.15), xzzx21_alpha (PTEQ_alpha).
The syndromes are drawn on the host from a seeded numpy generator so that both sides see the same ones: an i.i.d. error,
its equivalence class remembered, then a random logical operator applied (generate_data.py:122-131)."""
import concurrent.futures as cf
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (measurement script: the oracle is the checker here, as in bench.py's cpu_baseline)

OUT = os.path.join(ROOT, "gpurun_out")


def config(name):
    p = 0.15
    if name == "rotated25":
        return dict(g=O.ROTATED, L=25, kind=0, bottom=p, b=0.0, pxyz=(p / 3, p / 3, p / 3),
                    text="rotated surface code d=25, depolarizing p=0.15, PTEQ Nc=25 iters=10 p_logical=0.5 SEQ=2 TOPS=10 tops_burn=2 eps=0.1")
    eta = 100.0
    pz, px = p * eta / (eta + 1), p / (2 * (eta + 1))
    if name == "xzzx21_biased":
        return dict(g=O.XZZX, L=21, kind=2, bottom=p, b=eta, pxyz=(px, px, pz),
                    text="XZZX d=21, Z-biased eta=100 p=0.15, PTEQ_biased Nc=21 iters=10 p_logical=0.5 SEQ=2 TOPS=10 tops_burn=2 eps=0.1")
    if name == "xzzx21_alpha":
        pz_tilde = (p / (1 + 1 / eta)) / (1 - p)
        alpha = float(np.log(pz_tilde / (2 * eta)) / np.log(pz_tilde))
        return dict(g=O.XZZX, L=21, kind=1, bottom=pz_tilde, b=alpha, pxyz=(px, px, pz),
                    text="XZZX d=21, Z-biased eta=100 p=0.15, PTEQ_alpha pz_tilde=%.5f alpha=%.4f Nc=21" % (pz_tilde, alpha))
    raise SystemExit("unknown config " + name)


def syndromes(cfg, S):
    g, L = cfg["g"], cfg["L"]
    rng = np.random.default_rng(20262 + L)
    px, py, pz = cfg["pxyz"]
    qs, truth = np.zeros((S, L * L), np.uint8), np.zeros(S, np.int64)
    for s in range(S):
        r = rng.random(L * L)
        q = np.zeros(L * L, np.uint8)
        q[r < pz] = 3
        q[(r >= pz) & (r < pz + px)] = 1
        q[(r >= pz + px) & (r < pz + px + py)] = 2
        truth[s] = O.eq_class(g, L, q.reshape(L, L))
        q2, _ = O.apply_random_logical(g, L, q.reshape(L, L), O.Stream.mt(int(rng.integers(1 << 30))))
        qs[s] = np.asarray(q2, np.uint8).reshape(-1)
    return qs, truth


def summary(steps, converged, pct, truth, tops0):
    conv = converged.astype(bool)
    fail = pct.argmax(1) != truth
    n = len(truth)
    fr = float(fail.mean())
    return {"ladders": int(n), "converged": float(conv.mean()), "mean_ladder_steps": float(steps.mean()),
            "median_ladder_steps": float(np.median(steps)), "p90_ladder_steps": float(np.quantile(steps, 0.9)),
            "max_ladder_steps": int(steps.max()), "tops0_mean": float(tops0.mean()),
            "logical_failures": int(fail.sum()), "logical_failure_rate": fr, "sigma": float(np.sqrt(max(fr * (1 - fr), 1e-12) / n)),
            "logical_failure_rate_converged_only": float(fail[conv].mean()) if conv.any() else None}


def main():
    mode, name = sys.argv[1], sys.argv[2]
    cfg = config(name)
    g, L, kind = cfg["g"], cfg["L"], cfg["kind"]
    os.makedirs(OUT, exist_ok=True)
    if mode == "gpu":
        from mcmc_qec_toric_rl_b200 import _lib
        cap, S = int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 4736
        qs, truth = syndromes(cfg, S)
        ctx = _lib.Context(0)
        ctx.pteq(g, L, kind, qs, cfg["bottom"], param_b=cfg["b"], steps=50, conv=True, seed=1)
        t = time.perf_counter()
        pct, info = ctx.pteq(g, L, kind, qs, cfg["bottom"], param_b=cfg["b"], steps=cap, conv=True, seed=11)
        dt = time.perf_counter() - t
        st = info["stats"]
        np.savez_compressed(os.path.join(OUT, "r02_%s_gpu.npz" % name), pct=pct, truth=truth, steps=info["steps"], converged=info["converged"],
                            tops0=info["tops0"], counts=info["counts"], cap=cap)
        out = {"config": cfg["text"], "side": "gpu", "step_cap": cap, "seconds": dt, "syndromes_per_s": S / dt,
               "metropolis_steps": int(st["metropolis_steps"]), "steps_per_s": st["metropolis_steps"] / dt, "kernel_ms": st["chain_kernel_ms"],
               **summary(info["steps"], info["converged"], pct, truth, info["tops0"]), "device": ctx.device_info()["name"]}
        print(json.dumps(out), flush=True)
    elif mode == "oracle":
        cap, n = int(sys.argv[3]), int(sys.argv[4])
        threads = int(sys.argv[5]) if len(sys.argv) > 5 else len(os.sched_getaffinity(0))
        qs, truth = syndromes(cfg, n)          # the generator is sequential: the first n of any S are the same syndromes
        O.lib().qo_set_fast_windows(1)   # alpha ladders: running window sums instead of the O(history) recomputation per step

        def one(i):
            pct, w = O.pteq(kind, g, L, qs[i].reshape(L, L), cfg["bottom"], O.Stream.mt(7000 + i), O.Stream.py(9000 + i), param_b=cfg["b"], steps=cap)
            return pct, w["steps"], w["tops0"], w.get("converged", int(w["steps"] < cap))
        t = time.perf_counter()
        with cf.ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(one, range(n)))
        dt = time.perf_counter() - t
        pct = np.stack([r[0] for r in res])
        steps = np.array([r[1] for r in res])
        tops0 = np.array([r[2] for r in res])
        conv = np.array([r[3] for r in res])
        np.savez_compressed(os.path.join(OUT, "r02_%s_oracle.npz" % name), pct=pct, truth=truth, steps=steps, converged=conv, tops0=tops0, cap=cap)
        out = {"config": cfg["text"], "side": "oracle (C port, MT19937 streams)", "step_cap": cap, "seconds": dt, "threads": threads,
               "steps_per_s": float(steps.sum()) * L * 10 / dt, **summary(steps, conv, pct, truth, tops0)}
        print(json.dumps(out), flush=True)
    else:
        a = np.load(os.path.join(OUT, "r02_%s_gpu.npz" % name))
        b = np.load(os.path.join(OUT, "r02_%s_oracle.npz" % name))
        n = len(b["truth"])
        assert np.array_equal(a["truth"][:n], b["truth"])
        fa, fb = a["pct"][:n].argmax(1) != a["truth"][:n], b["pct"].argmax(1) != b["truth"]
        ph = (fa.sum() + fb.sum()) / (2.0 * n)
        out = {"config": cfg["text"], "side": "gpu vs oracle on the same syndromes", "syndromes": int(n),
               "gpu_failures": int(fa.sum()), "oracle_failures": int(fb.sum()), "sigma_of_difference": float(np.sqrt(2 * ph * (1 - ph) * n)),
               "same_choice": float((a["pct"][:n].argmax(1) == b["pct"].argmax(1)).mean()),
               "mean_abs_diff_points": float(np.abs(a["pct"][:n].astype(float) - b["pct"].astype(float)).mean()),
               "gpu_converged": float(a["converged"][:n].mean()), "oracle_converged": float(b["converged"].mean()),
               "gpu_mean_steps": float(a["steps"][:n].mean()), "oracle_mean_steps": float(b["steps"].mean())}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

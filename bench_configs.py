#!/usr/bin/env python
"""Secondary measurements: the BASELINE.json configs other than the headline one (which bench.py times).

    python bench_configs.py [--config toric5|rotated25|xzzx21|planar_points|all] [--out profiles/xxx.jsonl]
    torchrun --nproc-per-node N bench_configs.py --config planar_sweep [--syndromes 1000000]    (BASELINE config 5, sharded)
    torchrun --nproc-per-node N bench_configs.py --config toric15_strong [--syndromes 10000]    (config 2, strong scaling)

One JSON line per configuration: Metropolis steps/s and syndromes/s on one GPU through the host-buffer C ABI
(H2D + kernels + D2H inside the timed region), the logical failure rate of the decoded batch (true classes and hidden
lattices come from the device workload kernels), and -- the only use of oracle/ here, as in bench.py's cpu_baseline --
the oracle port timed on the host cores on a bounded sample of the same workload.  These lines are evidence for DESIGN.md; the
driver's contract line comes from bench.py."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def synth(shape, p, rng, px=None, py=None, pz=None):
    """i.i.d. Pauli errors: depolarizing rate p, or (px, py, pz)."""
    r = rng.random(shape)
    q = np.zeros(shape, np.uint8)
    if px is None:
        px = py = pz = p / 3.0
    q[r < pz] = 3
    q[(r >= pz) & (r < pz + px)] = 1
    q[(r >= pz + px) & (r < pz + px + py)] = 2
    return q


def hide_class(ctx, g, L, qs, rng):
    """generate_data.py:122-131 on the device: remember the class, then apply a random logical operator."""
    S = qs.shape[0]
    flat = np.ascontiguousarray(qs.reshape(S, -1))
    truth = ctx.define_equivalence_class(g, L, flat)
    hidden, _ = ctx.apply_random_logical(g, L, flat, seed=int(rng.integers(1, 2**31)))
    return hidden.reshape(qs.shape), truth


def timed(fn):
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def run_toric5(ctx, O):
    g, L, S, p = O.TORIC, 5, 100, 0.10
    rng = np.random.default_rng(1)
    qs, truth = hide_class(ctx, g, L, synth((S, 2, L, L), p, rng), rng)
    qm = np.ascontiguousarray(qs.reshape(S, -1))
    drop, steps = 10, 5 * L ** 4
    res = {}
    for name, fn in (("STDC", ctx.stdc), ("STRC", ctx.strc)):
        for w in range(2):                                # warm-up at full size (allocations, module load, memory query)
            fn(g, g, L, qm, p, 0.25, drop, steps, seed=1 + w)
        # a 4 ms problem: the best of three calls (host-side jitter is a large part of a single one)
        (out, st), dt = min((timed(lambda: fn(g, g, L, qm, p, 0.25, drop, steps, seed=7)) for _ in range(3)), key=lambda t: t[1])
        res[name] = {"steps_per_s": st["metropolis_steps"] / dt, "syndromes_per_s": S / dt, "seconds": dt,
                     "logical_failure_rate": float((out.argmax(1) != truth).mean())}
    t0 = time.perf_counter()
    ref = O.stdc_batch(g, g, L, qm, p, 0.25, drop, steps, seed=3, threads=cores())
    dt = time.perf_counter() - t0
    res["cpu_baseline"] = {"value": S * 16 * drop * steps * 5 / dt, "unit": "steps/s", "cores": cores(), "kind": "port",
                           "sample": f"all {S} syndromes, STDC", "logical_failure_rate": float((ref.argmax(1) != truth).mean())}
    return {"config": "toric d=5, depolarizing p=0.10, STDC/STRC of 100 syndromes, droplets=10, steps=3125, p_sampling=0.25",
            **res}


def _pteq_config(ctx, O, name, g, L, kind, bottom, b, S, steps, qs, truth, cpu_ladders, cpu_steps):
    qm = np.ascontiguousarray(qs.reshape(S, -1))
    # warm-up at full size: the context's device buffers (ladder states, n_err history) grow to the batch on first use
    ctx.pteq(g, L, kind, qm, bottom, param_b=b, steps=50, conv=False, seed=1)
    ctx.pteq(g, L, kind, qm, bottom, param_b=b, steps=steps, conv=True, seed=1)
    (pct, info), dt = timed(lambda: ctx.pteq(g, L, kind, qm, bottom, param_b=b, steps=steps, conv=False, seed=11))
    msteps = info["stats"]["metropolis_steps"]
    res = {"config": name, "ladders": S, "Nc": L, "ladder_steps": steps, "steps_per_s": msteps / dt, "syndromes_per_s": S / dt,
           "seconds": dt, "kernel_ms": info["stats"]["chain_kernel_ms"], "kernel_steps_per_s": msteps / (info["stats"]["chain_kernel_ms"] * 1e-3),
           "accept_rate": info["stats"]["accepted"] / msteps, "logical_failure_rate": float((pct.argmax(1) != truth).mean()),
           "tops0_mean": float(info["tops0"].mean())}
    # with the reference's convergence criterion
    (pct2, info2), dt2 = timed(lambda: ctx.pteq(g, L, kind, qm, bottom, param_b=b, steps=steps, conv=True, seed=12))
    res["with_convergence"] = {"seconds": dt2, "syndromes_per_s": S / dt2, "converged": float(info2["converged"].mean()),
                               "mean_steps": float(info2["steps"].mean()),
                               "logical_failure_rate": float((pct2.argmax(1) != truth).mean())}
    import concurrent.futures as cf

    def one(i):
        return O.pteq(kind, g, L, qs[i], bottom, O.Stream.mt(100 + i), O.Stream.py(200 + i), param_b=b, steps=cpu_steps, conv=False)[0]
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(cores()) as ex:
        outs = list(ex.map(one, range(cpu_ladders)))
    dtc = time.perf_counter() - t0
    res["cpu_baseline"] = {"value": cpu_ladders * L * 10 * cpu_steps / dtc, "unit": "steps/s", "cores": cores(), "kind": "port",
                           "sample": f"{cpu_ladders} ladders x {cpu_steps} ladder steps ({dtc:.1f} s)"}
    return res


def run_rotated25(ctx, O, S=4736, steps=2000):
    g, L, p = O.ROTATED, 25, 0.15
    rng = np.random.default_rng(3)
    qs, truth = hide_class(ctx, g, L, synth((S, L, L), p, rng), rng)
    return _pteq_config(ctx, O, "rotated surface code d=25, depolarizing p=0.15, PTEQ Nc=25 iters=10 p_logical=0.5",
                        g, L, 0, p, 0.0, S, steps, qs, truth, 2 * cores(), 12000)


def run_xzzx21(ctx, O, S=4736, steps=2000):
    g, L, p, eta = O.XZZX, 21, 0.15, 100.0
    rng = np.random.default_rng(4)
    pz, px = p * eta / (eta + 1), p / (2 * (eta + 1))
    qs, truth = hide_class(ctx, g, L, synth((S, L, L), p, rng, px, px, pz), rng)
    out = {"biased": _pteq_config(ctx, O, "XZZX d=21, Z-biased eta=100 p=0.15, PTEQ_biased Nc=21", g, L, 2, p, eta, S, steps, qs, truth,
                                  2 * cores(), 1500)}
    pz_tilde = (p / (1 + 1 / eta)) / (1 - p)
    alpha = float(np.log(pz_tilde / (2 * eta)) / np.log(pz_tilde))
    out["alpha"] = _pteq_config(ctx, O, f"XZZX d=21, pz_tilde={pz_tilde:.5f} alpha={alpha:.4f}, PTEQ_alpha Nc=21", g, L, 1, pz_tilde,
                                alpha, S, steps, qs, truth, 2 * cores(), 1500)
    return out


def run_planar_sweep(ctx, O, ps=(0.10, 0.15, 0.20), ds=(7, 11, 15, 21), droplets=16):
    """Threshold sweep: STDC, classes reached on device, steps = d^4, one GPU-filling batch per (d, p)."""
    g = O.PLANAR
    info = ctx.device_info()
    # fix the table budget once: otherwise every call asks the driver for the free memory, which costs ~10 ms on a context
    # that already holds tens of GB -- more than a whole d=7 batch takes
    ctx.set_table_budget(int(info["free_mem"] * 0.8))
    rows = []
    for d in ds:
        steps = d ** 4
        cap = 1
        while cap < droplets * steps * 1.25 + 1:
            cap *= 2
        fit = int(info["free_mem"] * 0.8) // (4 * cap * 8)
        S = max(8, min(fit, (info["sm_count"] * 1024) // (4 * droplets)))
        # one batch = whole rounds of the kernel's CTAs within one wave, from the library's plan of a small probe call
        probe = synth((64, 2, d, d), 0.1, np.random.default_rng(1))
        probe[:, 1, -1, :] = 0
        probe[:, 1, :, -1] = 0
        ctx.stdc(g, g, d, np.ascontiguousarray(probe.reshape(64, -1)), 0.1, 0.25, droplets, steps, seed=1)
        wave_cap, round_chains = ctx.last_plan()
        per_round = max(1, round_chains // (4 * droplets))
        S = int(per_round if wave_cap >= per_round else wave_cap)
        first = True
        for p in ps:
            rng = np.random.default_rng(50 + d)
            raw = synth((S, 2, d, d), p, rng)
            raw[:, 1, -1, :] = 0
            raw[:, 1, :, -1] = 0
            qs, truth = hide_class(ctx, g, d, raw, rng)
            qm = np.ascontiguousarray(qs.reshape(S, -1))
            if first:   # allocations for this size (key logs of the whole batch at full length) happen outside the timed call
                ctx.stdc(g, g, d, qm, p, 0.25, droplets, steps, seed=1)
                first = False
            (out, st), dt = timed(lambda: ctx.stdc(g, g, d, qm, p, 0.25, droplets, steps, seed=5))
            rows.append({"d": d, "p": p, "syndromes": S, "steps_per_s": st["metropolis_steps"] / dt, "syndromes_per_s": S / dt,
                         "seconds": dt, "logical_failure_rate": float((out.argmax(1) != truth).mean()),
                         "distinct_per_sample": st["distinct"] / (S * 4 * droplets * steps)})
    return {"config": "planar code threshold sweep, STDC, 4 classes x 16 chains, steps = d^4, p_sampling = 0.25", "points": rows}


AB_OLD_SIZING = False


def _dist_env():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    return rank, world, local


def _run_sharded(name, text, params_of, points, per_point, chunk_of, steps_per_syndrome, out_path, warm, rounds_of=None):
    """Common driver of the two multi-GPU workloads: one process per GPU (torchrun), the (point, chunk) items balanced by
    cost over the ranks, each decoded by generate_data.generate_batch with the lattices resident in HBM from the error
    draw to the failure count, and the failure counts + class distributions gathered on the host at the end."""
    import torch
    import torch.distributed as dist
    from mcmc_qec_toric_rl_b200 import _lib, generate_data, sharding
    rank, world, local = _dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.default_context(local)
    info = ctx.device_info()
    ctx.set_table_budget(int(info["free_mem"] * 0.8))
    if AB_OLD_SIZING:
        ctx.debug_set("packed", 0)
        rounds_of = None
    planned = {}
    for pt in warm:                                            # module load and first allocations, untimed
        generate_data.generate_batch(params_of(pt), 64, seed=1, device=local)
        if rounds_of is not None:
            # size the items of this distance from the library's own plan (qecmc_last_plan): whole rounds of full-size CTAs
            # over the SMs, within what one wave may hold in the table budget -- no SM idles behind a short last round
            par = params_of(pt)
            wave_cap, round_chains = ctx.last_plan()
            per_round = max(1, round_chains // ((16 if par["code"] == "toric" else 4) * par["droplets"]))
            want = per_round * rounds_of(pt)
            n = want if wave_cap >= want else (wave_cap // per_round * per_round if wave_cap >= per_round else wave_cap)
            planned[pt["d"]] = int(max(64, n))
    if planned:
        ds_sorted = sorted(planned)
        t = torch.tensor([planned[d] for d in ds_sorted], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)           # every rank cuts the same items
        planned = {d: int(v) for d, v in zip(ds_sorted, t.tolist())}
        chunk_of = lambda pt: planned[pt["d"]]                  # noqa: E731

    def decode_chunk(pt, n, item):
        res = generate_data.generate_batch(params_of(pt), n, seed=1000003 * (item + 1), device=local)
        return res["failures"], res["distr"].astype(np.float32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    # the gather inside run_sweep_sharded is part of the job; the decode time alone is taken per rank before it
    local_t = {}

    def timed_chunk(pt, n, item):
        out = decode_chunk(pt, n, item)
        torch.cuda.synchronize()
        local_t["decode_s"] = time.perf_counter() - t0
        return out
    curve = sharding.run_sweep_sharded(points, per_point, chunk_of, timed_chunk, rank=rank, world=world)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3, local_t.get("decode_s", 0.0), wall], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [[round(float(x), 3) for x in a.tolist()] for a in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    seconds, decode_s, wall = [float(x) for x in t.tolist()]
    n_total = sum(o["syndromes"] for o in curve)
    steps_total = float(sum(o["syndromes"] * steps_per_syndrome(o["point"]) for o in curve))
    line = None
    if rank == 0:
        rows = []
        for o in curve:
            row = dict(o["point"], syndromes=o["syndromes"], failures=o["failures"], logical_failure_rate=o["rate"], sigma=o["sigma"])
            if o["extra"] is not None:
                row["mean_class_distribution"] = [round(float(x), 3) for x in o["extra"].mean(0)]
            rows.append(row)
        line = {"name": name, "config": text, "n_gpus": world, "syndromes": n_total, "seconds": seconds, "syndromes_per_s": n_total / seconds,
                "metropolis_steps": steps_total, "steps_per_s": steps_total / seconds, "decode_seconds_max_over_ranks": decode_s,
                "gathered": "failure counts and float32 class distributions of every syndrome (all_gather_object of host arrays)",
                "gathered_bytes": int(sum(o["extra"].nbytes for o in curve if o["extra"] is not None)),
                "timing": "CUDA events on rank-local default streams around decode + gather, max over ranks",
                "per_rank_seconds": per_rank, "curve": rows, "device": info["name"]}
        if planned:
            line["syndromes_per_item"] = planned
        print(json.dumps(line), flush=True)
        if out_path:
            with open(out_path, "a") as f:
                f.write(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()
    return line


def run_planar_sweep_sharded(args):
    """BASELINE config 5: planar threshold sweep d in {7,11,15,21}, p in [0.10,0.20], `--syndromes` in total (default 1e6),
    STDC with 4 classes x 16 chains, d^4 samples per chain, p_sampling 0.25 (generate_data.py:53-261 per syndrome)."""
    ds, ps, droplets = (7, 11, 15, 21), [round(0.10 + 0.01 * i, 2) for i in range(11)], 16
    points = [dict(d=d, p=p) for d in ds for p in ps]
    per_point = -(-args.syndromes // len(points))
    fill = 148 * 1024 // (4 * droplets)

    def params_of(pt):
        return dict(code="planar", method="STDC", size=pt["d"], noise="depolarizing", p_error=pt["p"], p_sampling=0.25,
                    droplets=droplets, steps=pt["d"] ** 4)

    def chunk_of(pt):                                           # about the same work per item, never less than a full GPU
        return fill * max(1, int(round((15.0 / pt["d"]) ** 4)))
    return _run_sharded("planar_sweep", "planar code threshold sweep d in {7,11,15,21} x p in {0.10..0.20}, STDC, 4 classes x 16 chains, "
                        "d^4 samples x 5 steps per chain, p_sampling 0.25; errors drawn, labelled, hidden, decoded and scored on the device",
                        params_of, points, per_point, chunk_of, lambda pt: 4 * droplets * pt["d"] ** 4 * 5, args.out,
                        [dict(d=d, p=0.15) for d in ds], rounds_of=lambda pt: max(1, int(round((15.0 / pt["d"]) ** 4))))


def run_toric15_strong(args):
    """BASELINE config 2 as a fixed job (strong scaling): `--syndromes` (default 10000) toric d=15 p=0.15 syndromes,
    16 classes x 64 chains x 15^4 samples x 5 steps, split evenly over the ranks."""
    _, world, _ = _dist_env()
    pt = dict(d=15, p=0.15)

    def params_of(pt):
        return dict(code="toric", method="STDC", size=15, p_error=0.15, p_sampling=0.25, droplets=64, steps=15 ** 4)
    share = -(-args.syndromes // world)
    return _run_sharded("toric15_strong", "toric d=15 depolarizing p=0.15, STDC 16 classes x 64 chains x 15^4 samples x 5 steps; fixed job "
                        "split over the ranks (strong scaling)", params_of, [pt], args.syndromes, share,
                        lambda pt: 16 * 64 * 15 ** 4 * 5, args.out, [pt])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="all")
    ap.add_argument("--out", default="")
    ap.add_argument("--gpus", type=int, default=1, help="informational: the rank count comes from torchrun's WORLD_SIZE")
    ap.add_argument("--syndromes", type=int, default=0, help="total syndromes of the sharded workloads")
    ap.add_argument("--ab-old-sizing", action="store_true",
                    help="A/B switch of the planar sweep: fixed items of 2368 syndromes and the 64-bit row-word kernel for d = 21")
    args = ap.parse_args()
    global AB_OLD_SIZING
    AB_OLD_SIZING = args.ab_old_sizing
    if args.config == "planar_sweep":
        args.syndromes = args.syndromes or 1000000
        return run_planar_sweep_sharded(args)
    if args.config == "toric15_strong":
        args.syndromes = args.syndromes or 10000
        return run_toric15_strong(args)
    from mcmc_qec_toric_rl_b200 import _lib
    from oracle import oracle as O
    O.lib()
    ctx = _lib.Context(0)
    runs = {"toric5": run_toric5, "rotated25": run_rotated25, "xzzx21": run_xzzx21, "planar_points": run_planar_sweep}
    todo = list(runs) if args.config == "all" else [args.config]
    lines = []
    for name in todo:
        res = runs[name](ctx, O)
        res["name"] = name
        res["device"] = ctx.device_info()["name"]
        line = json.dumps(res)
        print(line, flush=True)
        lines.append(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()

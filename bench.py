#!/usr/bin/env python
"""Benchmark of the Metropolis-chain decoder hot path (BASELINE.json north_star).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): toric code d=15, depolarizing p=0.15, STDC-style decoding with
16 equivalence classes x 64 chains per syndrome, 15^4 = 50625 samples per chain, 5 Metropolis steps
per sample, p_sampling = 0.25.  One bench "step" decodes one batch of synthetic syndromes
(--syndromes per GPU; default: enough to fill every SM once).  With N > 1 (torchrun, one rank per
GPU) every rank decodes its own batch (weak scaling, no collective on the data path).

Prints ONE JSON line (rank 0).  value = Metropolis steps/s with inputs resident in HBM;
e2e = the same through the host-buffer C-ABI call (H2D + kernels + D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L = 15
P_ERROR = 0.15
P_SAMPLING = 0.25
DROPLETS = 64
SAMPLES = L ** 4
ITERS = 5
N_EQ = 16
W_INT = 170.0      # SURVEY.md 8d's pre-implementation estimate of int32 lane-ops per toric/planar step (kept for reference only)
W_LOG = 8.0        # bytes a chain appends to a bucket log per offered sample (the chain kernel's only steady HBM traffic)
NCU_SUMMARY = "profiles/r02_ncu_summary.json"   # written by profiles/scripts/ncu_summary.py from ncu captures of this command


def ncu_summary():
    """Per-kernel ncu figures of the last committed capture (git sha inside): the chain kernel's counted lane-ops per step
    (thread instructions executed / Metropolis steps), issue-slot and active-lane utilisation, pipes, DRAM bytes."""
    try:
        with open(os.path.join(ROOT, NCU_SUMMARY)) as f:
            return json.load(f)
    except Exception:
        return {"git_sha": None, "kernels": {}}


def synth_syndromes(n, seed, L=L, p=P_ERROR):
    """i.i.d. depolarizing errors (distribution of Toric_code.generate_random_error, toric_model.py:15-24)."""
    rng = np.random.default_rng(seed)
    q = (rng.random((n, 2, L, L)) < p) * rng.integers(1, 4, (n, 2, L, L))
    return np.ascontiguousarray(q.astype(np.uint8).reshape(n, -1))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """keep the samples taken inside [t0, t1] (the timed region); NVML starts up during the warm-up steps"""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        inside = [ln for (t, ln) in self.lines if t0 is None or t0 - 0.1 <= t <= t1 + 0.1]
        for ln in inside or [ln for (_, ln) in self.lines]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(n_syndromes, droplets, samples, threads, seed=4242):
    """Oracle port (oracle/qec_oracle.c) of STDC on the host cores; returns (steps/s, seconds, steps)."""
    from oracle import oracle as O
    qm = synth_syndromes(n_syndromes, seed)
    O.lib()
    t0 = time.perf_counter()
    O.stdc_batch(O.TORIC, O.TORIC, L, qm, P_ERROR, P_SAMPLING, droplets, samples, seed=seed, iters=ITERS, threads=threads)
    dt = time.perf_counter() - t0
    steps = n_syndromes * N_EQ * droplets * samples * ITERS
    return steps / dt, dt, steps


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the reference is Python/numba and cannot travel to the GPU box, so this arm times the
    oracle port of the same path (kind 'port') on all host cores.  Rank 0 only."""
    if rank != 0:
        return
    cores = host_cores()
    # same configuration as the GPU arm (16 classes x 64 chains x 15^4 samples x 5 steps per syndrome); the bounded sample
    # is the number of syndromes per step: one per 16 host threads (a syndrome is 2.6e8 Metropolis steps, ~65 s of one
    # core; its 16 classes are the parallel jobs)
    drop = DROPLETS
    n_syn = max(1, cores // 16)
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_port_rate(max(1, n_syn // 8), 1, 2000, cores)
    t_total, steps_total = 0.0, 0
    for k in range(args.steps):
        rate, dt, steps = cpu_port_rate(n_syn, drop, args.samples, cores, seed=100 + k)
        t_total += dt
        steps_total += steps
    value = steps_total / t_total
    sample = f"{n_syn} syndromes x 16 classes x {drop} chains x {args.samples} samples x {ITERS} steps per bench step"
    line = {
        "impl": "reference", "metric": "metropolis_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "syndromes_per_s": value / (N_EQ * DROPLETS * args.samples * ITERS),
        "config": workload_config(args.gpus), "syndromes_per_step_per_gpu": n_syn,
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    """identical for both arms (the batch of a step is reported beside it, not inside)"""
    return {"workload": "toric d=15 depolarizing p=0.15, STDC: 16 classes x 64 chains x 15^4 samples x 5 steps, p_sampling=0.25",
            "parallelism": f"syndrome-sharded x{n_gpus}, no collective",
            "l2": "working set = per-chain key logs (tens of GB, rewritten every step) >> 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--syndromes", type=int, default=0, help="syndromes per step per GPU (0 = fill the GPU once)")
    ap.add_argument("--samples", type=int, default=SAMPLES, help="samples per chain (default 15^4; smaller only for debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # NCCL (and anything else native) may print to stdout; the contract is ONE JSON line there, so fd 1 points at
    # stderr until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from mcmc_qec_toric_rl_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.Context(local_rank)
    info = ctx.device_info()
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    samples = args.samples
    steps_per_syndrome = N_EQ * DROPLETS * samples * ITERS
    # table arena: leave room for torch + staging
    ctx.set_table_budget(int(info["free_mem"] * 0.88))
    keys_max = DROPLETS * samples                              # worst case: every sample of every droplet logs a key
    per_table = (128 * ((keys_max + 127) // 128 + 66) + max(1024, keys_max // 16)) * 8   # bucket logs + overflow log
    fit = int(info["free_mem"] * 0.88) // (N_EQ * per_table)
    fill = (info["sm_count"] * 1024) // (N_EQ * DROPLETS)  # chains resident per wave: one 1024-thread CTA per SM
    batch = args.syndromes or max(1, min(fit, fill))
    n_steps = args.steps + args.warmup
    host = [torch.from_numpy(synth_syndromes(batch, 1000 * rank + k)).pin_memory() for k in range(n_steps)]
    dev = [h.cuda(non_blocking=True) for h in host]
    out = torch.zeros((batch, N_EQ), dtype=torch.float64, device="cuda")
    out_host = torch.zeros((batch, N_EQ), dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(k):
        return ctx.stdc_dev(_lib.TORIC, _lib.TORIC, L, dev[k].data_ptr(), batch, out.data_ptr(), P_ERROR, P_SAMPLING,
                            DROPLETS, samples, iters=ITERS, randomize=True, seed=7 + k)

    def step_e2e(k):
        q = host[k].numpy()
        res, st = ctx.stdc(_lib.TORIC, _lib.TORIC, L, q, P_ERROR, P_SAMPLING, DROPLETS, samples, iters=ITERS, randomize=True,
                           seed=7 + k)
        return res, st

    # ---- device-resident timing ----
    # nvidia-smi starts before the warm-up steps: its start-up (NVML initialisation) holds driver locks for tens of
    # milliseconds and would otherwise land inside the timed region; only samples from the timed region are reported
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for k in range(args.warmup):
        step_dev(k)
    barrier()
    t_region0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    kern_ms, launches, accepted, offered, distinct = 0.0, 0, 0, 0, 0
    call_ms = []
    for k in range(args.warmup, n_steps):
        st = step_dev(k)
        kern_ms += st["chain_kernel_ms"]
        call_ms.append(round(st["total_ms"], 2))
        launches += st["kernel_launches"]
        accepted += st["accepted"]
        offered += st["samples"]
        distinct += st["distinct"]
        waves = st["waves"]
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.window(t_region0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None
    # ---- end-to-end timing (host buffers through the C ABI) ----
    step_e2e(0)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    n_e2e = args.steps
    for k in range(n_e2e):
        res, _ = step_e2e(args.warmup + k)
    f1.record(stream)
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    per_rank = None
    if world > 1:
        t = torch.tensor([ms, ms_e2e, kern_ms], dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [[round(float(x), 3) for x in a.tolist()] for a in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, kern_ms = [float(x) for x in t.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_steps = world * batch * steps_per_syndrome * args.steps
    value = total_steps / (ms * 1e-3)
    e2e_value = world * batch * steps_per_syndrome * n_e2e / (ms_e2e * 1e-3)
    peaks, peak_src = measured_peaks()
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    peak_ops = info["sm_count"] * 128 * sm_max_mhz * 1e6            # int32 lane-ops/s, one GPU
    kern_steps_per_s = batch * steps_per_syndrome * args.steps / (kern_ms * 1e-3)   # per GPU, chain kernel only (CUDA events)
    ncu = ncu_summary()
    ck = ncu["kernels"].get("chain", {})
    full_size = samples == SAMPLES and ck.get("steps_per_launch") == batch * steps_per_syndrome
    # useful lane-ops per Metropolis step = thread instructions the chain kernel executes per step (ncu, predicated-off and
    # idle lanes excluded); issue slots per step = warp instructions x 32 / steps (idle lanes included)
    useful = ck.get("thread_inst_per_step")
    slots = ck.get("issue_slots_per_step")
    achieved = kern_steps_per_s * useful if useful else None
    hbm_alg = (offered / args.steps) * W_LOG / ((kern_ms / args.steps) * 1e-3) / 1e9  # GB/s, rank 0
    keep = ("duration_ms", "issue_slot_utilisation", "active_lane_utilisation", "lanes_per_inst", "issue_slots_per_step",
            "thread_inst_per_step", "pipe_alu_pct", "pipe_fma_pct", "pipe_lsu_pct", "pipe_fp64_pct", "smem_wavefront_pct",
            "dram_read_bytes", "dram_write_bytes", "registers", "warps_active_pct", "stall_barrier", "stall_wait",
            "stall_short_scoreboard", "kernel", "source")
    others = {k: {f: v[f] for f in keep if f in v} for k, v in ncu["kernels"].items() if k != "chain"}
    line = {
        "metric": "metropolis_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "syndromes_per_s": value / steps_per_syndrome,
        "config": workload_config(world), "syndromes_per_step_per_gpu": batch,
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": int(batch * 2 * L * L),
                "d2h_bytes_per_step": int(batch * N_EQ * 8), "steps_timed": n_e2e},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "alu", "achieved": achieved / 1e12 if achieved else None, "peak": peak_ops / 1e12, "unit": "Tlaneop/s",
                     "frac": achieved / peak_ops if achieved else None,
                     "traffic": (ck.get("dram_read_bytes", 0) + ck.get("dram_write_bytes", 0)) if full_size and ck else None,
                     "what": "achieved = steps/s of the chain kernel (CUDA events, this run) x useful lane-ops per step (thread "
                             "instructions executed per Metropolis step, counted by ncu on this command at ncu.git_sha: idle and "
                             "predicated-off lanes do not count); peak = SMs x 128 int32 lanes x max SM clock",
                     "useful_laneops_per_step": useful, "issue_slots_per_step": slots,
                     "issue_slot_utilisation": (kern_steps_per_s * slots / peak_ops) if slots else None,
                     "active_lane_utilisation": (kern_steps_per_s * useful / peak_ops) if useful else None,
                     "pipes_pct_of_peak": {k: ck.get(k) for k in ("pipe_alu_pct", "pipe_fma_pct", "pipe_lsu_pct", "smem_wavefront_pct")},
                     "ncu": {"file": NCU_SUMMARY, "git_sha": ncu.get("git_sha"), "capture": ck.get("source"),
                             "issue_slot_utilisation": ck.get("issue_slot_utilisation"),
                             "active_lane_utilisation": ck.get("active_lane_utilisation"), "kernel_ms": ck.get("duration_ms")},
                     "survey_8d_estimate": {"laneops_per_step": W_INT, "frac": kern_steps_per_s * W_INT / peak_ops,
                                            "note": "SURVEY.md 8d's pre-implementation figure; the kernel needs a third of it, "
                                                    "so this exceeds 1 and says nothing about headroom"},
                     "kernel": "stdc_fast_kernel<TORIC,u32,native,STDC,BLOG>", "kernel_ms_per_launch": kern_ms / (args.steps * waves),
                     "units_per_launch": batch * steps_per_syndrome / waves,
                     "peak_source": f"{info['sm_count']} SMs x 128 int32 lanes x {sm_max_mhz:.0f} MHz ({peak_src} max SM clock)",
                     "note": "north_star: this path is integer-issue bound, not HBM or tensor bound (SURVEY.md 8d)",
                     "hbm": {"achieved": hbm_alg, "peak": float(peaks.get("hbm_gbs", 6650.0)), "unit": "GB/s",
                             "frac": hbm_alg / float(peaks.get("hbm_gbs", 6650.0)),
                             "what": "bucket-log appends, 8 B per offered sample (the dedupe kernel streams them afterwards)",
                             "peak_source": peak_src},
                     "other_kernels_ms_per_step": (ms - kern_ms) / args.steps, "call_ms": call_ms,
                     "other_kernels_ncu": others},
        "chain_stats": {"accept_rate": accepted / (batch * steps_per_syndrome * args.steps),
                        "offered_per_sample": offered / (batch * N_EQ * DROPLETS * samples * args.steps),
                        "distinct_per_sample": distinct / (batch * N_EQ * DROPLETS * samples * args.steps),
                        "waves_per_step": waves},
        "device": info["name"],
    }
    if per_rank:
        line["per_rank_ms"] = {"what": "[timed region, e2e region, chain kernels] per rank", "values": per_rank}
    if not args.no_cpu_baseline and world == 1:
        cores = host_cores()
        n_syn = 3 * max(1, cores // 16)   # sized for 10-30 s of host work: a syndrome's 16 classes are the parallel jobs
        cpu_port_rate(1, 1, 2000, 1)  # warm the library
        cpu_drop = DROPLETS
        rate, dt, steps = cpu_port_rate(n_syn, cpu_drop, samples, cores)
        rate1, dt1, _ = cpu_port_rate(1, 1, samples, 1)   # the one-core figure beside the all-core one (~3 s)
        line["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": cores, "kind": "port", "value_1core": rate1,
                                "sample": f"{n_syn} syndromes x 16 classes x {cpu_drop} chains x {samples} samples x {ITERS} steps "
                                          f"({steps:.3g} Metropolis steps, {dt:.1f} s) with oracle/qec_oracle.c"}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

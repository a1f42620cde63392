"""Turn ncu exports into the per-kernel figures bench.py reports (profiles/r02_ncu_summary.json).

    ncu_summary.py <git-sha> <out.json> <name>=<csv>[:steps_per_launch] ...

Accepted inputs: the long format of `ncu --csv --metrics ...` (one row per metric, as written by --log-file) and the wide
format of `ncu -i x.ncu-rep --page raw --csv` (one row per kernel launch).  For every named kernel the summary keeps

  duration_ms, warp_inst (smsp__inst_executed.sum), thread_inst (smsp__thread_inst_executed.sum),
  issue_slot_utilisation  = smsp__issue_active.avg.pct_of_peak_sustained_active / 100,
  active_lane_utilisation = thread_inst / (32 * warp_inst) * issue_slot_utilisation   (lanes that did work per issue slot of the SM),
  lanes_per_inst          = thread_inst / warp_inst,
  pipe_alu / pipe_fma / pipe_lsu / pipe_fp64 (% of peak), smem_wavefront_pct,
  dram_read_bytes / dram_write_bytes, registers, block, grid, warps_active_pct, stall ratios,
  and, when steps_per_launch is given, issue_slots_per_step = warp_inst * 32 / steps and thread_inst_per_step.
"""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "smsp__inst_executed.sum": "warp_inst",
    "smsp__thread_inst_executed.sum": "thread_inst",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst",
    "launch__registers_per_thread": "registers",
    "launch__block_size": "block",
    "launch__grid_size": "grid",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__warps_active.avg.per_cycle_active": "warps_active_per_cycle",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def read(path, match):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    out = {}
    if "Metric Name" in hdr:                                   # long format
        iK, iN, iU, iV = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
        iID = hdr.index("ID")
        first = None
        for r in rows[1:]:
            if match not in r[iK]:
                continue
            first = r[iID] if first is None else first
            if r[iID] != first or r[iN] not in WANT:
                continue
            v = num(r[iV])
            if v is not None:
                out[WANT[r[iN]]] = v * UNIT_SCALE.get(r[iU], 1.0)
                out["kernel"] = r[iK]
    else:                                                      # wide format: header, units, one row per launch
        units = rows[1]
        iK = hdr.index("Kernel Name")
        for r in rows[2:]:
            if match not in r[iK]:
                continue
            out["kernel"] = r[iK]
            for k, nm in WANT.items():
                if k in hdr:
                    v = num(r[hdr.index(k)])
                    if v is not None:
                        out[nm] = v * UNIT_SCALE.get(units[hdr.index(k)], 1.0)
            break
    return out


def main():
    sha, dest = sys.argv[1], sys.argv[2]
    res = {"git_sha": sha, "kernels": {}}
    for spec in sys.argv[3:]:
        name, rest = spec.split("=", 1)
        parts = rest.split(":")
        path, match = parts[0], parts[1] if len(parts) > 1 and parts[1] else name
        steps = float(parts[2]) if len(parts) > 2 else None
        k = read(path, match)
        if not k:
            print("no launch of", match, "in", path, file=sys.stderr)
            continue
        k["source"] = path
        if "duration_ns" in k:
            k["duration_ms"] = k.pop("duration_ns") / 1e6
        if k.get("warp_inst") and k.get("thread_inst"):
            k["lanes_per_inst"] = k["thread_inst"] / k["warp_inst"]
        elif k.get("warp_inst") and k.get("lanes_per_inst"):
            k["thread_inst"] = k["warp_inst"] * k["lanes_per_inst"]
        if "issue_active_pct" in k and "lanes_per_inst" in k:
            k["issue_slot_utilisation"] = k["issue_active_pct"] / 100.0
            k["active_lane_utilisation"] = k["issue_slot_utilisation"] * k["lanes_per_inst"] / 32.0
        if steps:
            k["steps_per_launch"] = steps
            if k.get("warp_inst"):
                k["issue_slots_per_step"] = k["warp_inst"] * 32.0 / steps
            if k.get("thread_inst"):
                k["thread_inst_per_step"] = k["thread_inst"] / steps
        res["kernels"][name] = k
    json.dump(res, open(dest, "w"), indent=1, sort_keys=True)
    print(json.dumps(res["kernels"], indent=1, sort_keys=True))


if __name__ == "__main__":
    main()

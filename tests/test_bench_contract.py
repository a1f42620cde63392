"""bench.py's reference arm (the oracle port on the host cores) runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "metropolis_steps_per_s" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["value"] > 1e5
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "toric d=15" in d["config"]["workload"]


def test_product_arm_refuses_to_run_without_a_gpu():
    """no CPU fallback: without a CUDA device the product arm must fail loudly, not fall back to the oracle"""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert "CUDA" in (out.stderr + out.stdout)


def test_roofline_figures_come_from_a_committed_ncu_capture():
    """bench.py holds no ncu constants: the chain kernel's counted lane-ops per step, issue-slot and active-lane
    utilisation are read from profiles/r02_ncu_summary.json, which profiles/scripts/ncu_summary.py rebuilds from the
    committed ncu exports (the git sha of the capture rides along)."""
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ncu_summary()
    assert s["git_sha"]
    ck = s["kernels"]["chain"]
    assert "stdc_fast_kernel" in ck["kernel"] and os.path.exists(os.path.join(ROOT, "profiles", os.path.basename(ck["source"])))
    assert ck["steps_per_launch"] == 148 * 16 * 64 * 15 ** 4 * 5
    assert 40 < ck["thread_inst_per_step"] < ck["issue_slots_per_step"] < 120       # useful lane-ops <= issue slots
    assert 0 < ck["active_lane_utilisation"] < ck["issue_slot_utilisation"] < 1
    for k in ("dedupe", "ladder_rotated25", "ladder_xzzx21_biased"):
        assert 0 < s["kernels"][k]["issue_slot_utilisation"] < 1, k
    # the summary is reproducible from the committed export
    sys.path.insert(0, os.path.join(ROOT, "profiles", "scripts"))
    import ncu_summary
    again = ncu_summary.read(os.path.join(ROOT, "profiles", "r02_ncu_fullsize_stdc.csv"), "stdc_fast_kernel")
    assert again["warp_inst"] == ck["warp_inst"] and again["thread_inst"] == ck["thread_inst"]

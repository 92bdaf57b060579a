"""The committed ncu summaries under profiles/ are reproducible from the committed raw
CSVs with the tools that made them (CPU only; no GPU, no ncu needed)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(tool, *args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), *args],
                          capture_output=True, text=True, check=True).stdout


def body(text):
    return [ln for ln in text.splitlines() if ln and not ln.startswith("#")]


def test_launch_summary_matches_committed():
    got = run("launch_summary.py", os.path.join(ROOT, "profiles", "r1_launches_c2.csv"))
    want = open(os.path.join(ROOT, "profiles", "r1_launches_c2_summary.txt")).read()
    assert body(got) == body(want)
    assert any("sensor_kernel" in ln for ln in body(got)[:2])   # K7 tops the launch list


def test_kernel_tables_match_committed():
    for w in ("c2", "c3"):
        got = run("ncu_step_summary.py", os.path.join(ROOT, "profiles", f"r1_kernels_{w}_ncu_raw.csv"))
        want = open(os.path.join(ROOT, "profiles", f"r1_kernels_{w}_ncu.txt")).read()
        assert body(got) == body(want), w


def test_k9_lane_summary_model_matches_sequential_sum(tmp_path):
    """tools/k9_lanes_model.cpp: the CPU model of the lane-wise resolution of binade-crossing chunks
    (prefix mode 2, prefix_dev.cuh lane_chunk)
    reproduces the sequential sum bit for bit on its adversarial inputs."""
    src = os.path.join(ROOT, "tools", "k9_lanes_model.cpp")
    exe = str(tmp_path / "k9model")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe, src], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "mismatches 0" in out.stdout

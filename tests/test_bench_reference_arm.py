"""bench.py --impl reference on CPU: the line the driver reads for the reference arm.

With oracle/_ref built (this container, or its prebuilt library on the GPU box) the arm times
the reference's own compiled sources in a child process ("kind": "reference"); with
QMCL_BENCH_CPU_PORT set it times the oracle port.  Both must print the contract's keys.
"""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_binding as rb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_arm(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                          "--steps", "2", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         env=env, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    lines = [ln for ln in out.stdout.decode().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def check_line(d):
    assert d["impl"] == "reference" and d["metric"] == "particle_point_likelihood_evals_per_sec"
    assert d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "C1"
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] == 1 and cb["sample"]
    assert set(cb["stage_ms"]) == {"odometry", "sensor", "resample", "estimate"}
    return cb


@pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built and no reference tree")
def test_reference_arm_times_the_compiled_reference():
    cb = check_line(run_arm())
    assert cb["kind"] == "reference"
    assert "oracle port" in cb["note"]          # the port's rate on the same sample rides along
    if cb.get("all_cores_reference"):           # one reference filter per host thread (>= 64 particles each)
        assert cb["all_cores_reference"]["kind"] == "reference" and cb["all_cores_reference"]["value"] > 0


def test_reference_arm_falls_back_to_the_oracle_port():
    cb = check_line(run_arm({"QMCL_BENCH_CPU_PORT": "1"}))
    assert cb["kind"] == "port" and "QMCL_BENCH_CPU_PORT" in cb["note"]

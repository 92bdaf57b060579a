"""The C++ host mirror of the reference's plugin seam (quickmcl_b200/host):
builds with plain g++ against the C ABI, refuses loudly without a device, and —
on the GPU box — reproduces the oracle through the reference-shaped classes
(tests/host/host_selftest.cpp drives it like laser_handler.cpp:236-276)."""
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELFTEST = os.path.join(ROOT, "tests", "host", "host_selftest")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "quickmcl_b200", "host")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "host")])


def test_host_library_builds_and_exports_the_seam():
    build()
    lib = os.path.join(ROOT, "quickmcl_b200", "host", "libquickmcl_b200_host.so")
    syms = subprocess.check_output(["nm", "-DC", lib], text=True)
    for want in ["quickmcl_b200::create_likelihood_map()",
                 "quickmcl_b200::create_particle_filter(",
                 "quickmcl_b200::CudaParticleFilter::update_importance_from_observations(",
                 "quickmcl_b200::CudaParticleFilter::resample_and_cluster()",
                 "quickmcl_b200::CudaParticleFilter::get_estimated_pose(",
                 "quickmcl_b200::CudaMapLikelihood::new_map(",
                 "quickmcl_b200::CudaMapLikelihood::update_importance("]:
        assert want in syms, want
    # the host library never links the oracle
    needed = subprocess.check_output(["readelf", "-d", lib], text=True)
    assert "oracle" not in needed


def test_host_refuses_without_a_device():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    build()
    out = subprocess.run([SELFTEST, "--expect-no-device"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_host_cycle_matches_oracle_on_gpu():
    build()
    out = subprocess.run([SELFTEST], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_selftest ok" in out.stdout


# ------------------------------------------------------------------ PoseCheckpoint vs the reference's PoseRestorer
PROBE = os.path.join(ROOT, "tests", "host", "pose_checkpoint_probe")


def _probe(*args):
    build()
    out = subprocess.check_output([PROBE] + [str(a) for a in args], text=True)
    return {ln.split()[0]: ln.split()[1:] for ln in out.splitlines()}


@pytest.mark.parametrize("stored", [
    None, [1.5, -2.25, 0.7, 0.04, 0.09, 0.01], [1.5, -2.25], [float("nan"), 3.0, float("nan"), 0.2, float("nan")],
    [], [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7], [1e-3 + 1e-12, 7.000000001, -3.14159265358979, 1e-9, 0.25, 0.01],
])
def test_cpp_pose_checkpoint_restores_like_the_reference_pose_restorer(stored):
    """quickmcl_b200::PoseCheckpoint::restore (C++ host mirror) against the reference's own
    PoseRestorer::restore_pose (pose_restore.cpp:57-97, compiled into oracle/_ref): the same
    re-initialisation arguments bit for bit.  CPU only."""
    import numpy as np
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built and no reference tree")
    calls, pose_ref, cov_ref, _ = rb.pose_restore(stored)
    if stored is None:
        got = _probe("restore", "none")
    elif not stored:
        got = _probe("restore", "empty")
    else:
        got = _probe("restore", *[repr(float(v)) for v in stored])
    assert int(got["calls"][0]) == calls == 1
    pose = np.array([float.fromhex(v) for v in got["pose"]], np.float32)
    cov = np.array([float.fromhex(v) for v in got["cov"]], np.float32).reshape(3, 3)
    assert pose.tobytes() == pose_ref.tobytes()
    assert cov.tobytes() == cov_ref.tobytes()


def test_cpp_pose_checkpoint_stores_like_the_reference_pose_restorer():
    import numpy as np
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built and no reference tree")
    pose = np.array([1.25, -0.5, 2.0])
    cov = np.array([[0.04, 0.01, 0.0], [0.01, 0.09, 0.0], [0.0, 0.0, 0.02]])
    stored_ref, _ = rb.pose_store(True, pose, cov)
    got = _probe("store", 1, *[repr(float(v)) for v in pose], *[repr(float(v)) for v in cov.reshape(9)])
    assert [float.fromhex(v) for v in got["stored"][1:]] == stored_ref and int(got["stored"][0]) == 6
    none_ref, _ = rb.pose_store(False, np.zeros(3), np.zeros((3, 3)))
    assert none_ref is None and _probe("store", 0)["stored"] == ["0"]


# ------------------------------------------------------------------ ScanCycle vs the reference's LaserHandler
CYCLE_PROBE = os.path.join(ROOT, "tests", "host", "scan_cycle_probe")


@pytest.mark.parametrize("resample_count,seed", [(2, 10), (1, 11), (4, 12)])
def test_cpp_scan_cycle_makes_the_calls_of_the_reference_laser_handler(resample_count, seed):
    """quickmcl_b200::ScanCycle (C++ host mirror, node-default motion thresholds) against the
    reference's own LaserHandler::Impl::cloud_callback (laser_handler.cpp:179-288, compiled into
    oracle/_ref): same IParticleFilter calls with the same arguments in the same order, same
    recompute_transform flag, over a random walk with stops, wraps, lost sets and pose resets."""
    import math
    import numpy as np
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built and no reference tree")
    build()
    rng = np.random.default_rng(seed)
    ref = rb.RefScanCycle(resample_count)                      # node defaults: 0.2f, float(pi / 6)
    odom = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-math.pi, math.pi)])
    last_run = None
    scenario, expected = [], []
    edge = float(np.float32(0.2))
    for step in range(150):
        kind = rng.integers(6)
        if kind == 1:
            odom = odom + np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(-0.2, 0.2)])
        elif kind == 2:
            odom = odom + np.array([rng.uniform(-0.6, 0.6), rng.uniform(-0.6, 0.6), rng.uniform(-0.3, 0.3)])
        elif kind == 3:
            odom = odom + np.array([0.0, 0.0, rng.uniform(-1.5, 1.5)])
        elif kind == 4:
            odom = odom + np.array([0.0, 0.0, 2 * math.pi * rng.choice([-1, 1])])
        elif kind == 5 and last_run is not None:               # on and next to the float threshold
            odom = np.array([last_run[0] + rng.choice([edge, np.nextafter(edge, 1.0), 0.2, -0.2]), last_run[1],
                             last_run[2]])
        reset, lose = int(rng.random() < 0.08), int(rng.random() < 0.1)
        n_points = int(rng.integers(0, 300))
        if reset:
            ref.force_pose_reset()
        if lose:
            ref.set_particles_present(False)
        lines = ref.on_scan(odom, n_points, now=float(step))
        if any(ln.startswith("handle_odometry") for ln in lines) or last_run is None:
            last_run = odom.copy()
        scenario.append(" ".join(float(v).hex() for v in odom) + f" {n_points} {lose} {reset}")
        expected.append(lines)
    out = subprocess.run([CYCLE_PROBE, str(resample_count)], input="\n".join(scenario) + "\n", text=True,
                         stdout=subprocess.PIPE, check=True).stdout.splitlines()

    def canon(ln):
        head, *rest = ln.split(" ")
        return (head, tuple(float.fromhex(v) for v in rest)) if head == "handle_odometry" else (ln,)

    pos = 0
    for step, lines in enumerate(expected):
        end = pos
        while not out[end].startswith("result "):
            end += 1
        got_calls, flags = out[pos:end], [int(v) for v in out[end].split()[1:]]
        pos = end + 1
        ref_calls = [ln for ln in lines if not ln.startswith(("publish_", "log: "))]
        assert [canon(ln) for ln in got_calls] == [canon(ln) for ln in ref_calls], step
        processed, resampled, did_global, recompute = flags
        assert bool(recompute) == ("publish_estimated_pose recompute" in lines), step
        assert bool(processed) == any(ln.startswith("handle_odometry") for ln in lines)
        assert bool(resampled) == ("resample_and_cluster" in lines)
        assert bool(did_global) == ("global_localization" in lines)
    assert pos == len(out)

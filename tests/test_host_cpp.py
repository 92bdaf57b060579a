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

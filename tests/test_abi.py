"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/qmcl.h declares, and struct layouts match the reference's."""
import ctypes as C
import os
import re

import pytest

from quickmcl_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qmcl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qmcl_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    # the reference interface: 6 Map virtuals + 10 IParticleFilter virtuals + 2 factories
    for must in ["qmcl_map_create", "qmcl_map_set_params", "qmcl_map_load",
                 "qmcl_map_update_importance", "qmcl_map_generate_random_pose",
                 "qmcl_map_likelihood_grid", "qmcl_map_state_at",
                 "qmcl_filter_create", "qmcl_filter_initialise", "qmcl_filter_global_localization",
                 "qmcl_filter_handle_odometry", "qmcl_filter_update_sensor",
                 "qmcl_filter_resample_and_cluster", "qmcl_filter_cluster",
                 "qmcl_filter_set_params", "qmcl_filter_get_estimate",
                 "qmcl_filter_copy_particles", "qmcl_filter_num_particles"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_abi.LIB_PATH):
        pytest.fail(f"{_abi.LIB_PATH} not built (run __graft_entry__.build())")
    L = C.CDLL(_abi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol():
    bound = set(_abi.SIGNATURES) | set(_abi._OTHER)
    assert set(declared_symbols()) == bound


def test_struct_layouts():
    # quickmcl::WeightedParticle: Pose2D<float> + double = 24 bytes (particle.h:32-47)
    assert C.sizeof(_abi.Particle) == 24
    assert _abi.Particle.weight.offset == 16
    assert _abi.PARTICLE_DTYPE.itemsize == 24
    assert C.sizeof(_abi.LikelihoodParams) == 24
    assert C.sizeof(_abi.PFParams) == 56


def test_no_device_fails_loudly():
    """Without a GPU the product path must refuse, not fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _abi.lib()
    h = C.c_void_p()
    st = L.qmcl_map_create(-1, C.byref(h))
    assert st == 2  # QMCL_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.qmcl_last_error()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the package may import, link or load it."""
    pkg = os.path.join(ROOT, "quickmcl_b200")
    pat = re.compile(r"(from\s+oracle|import\s+oracle|oracle/|libqmcl_oracle|qmcl_oracle\.h)")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert not pat.search(text), (dirpath, fn)

"""CPU tests (no GPU): the C-ABI libraries load, export every symbol include/dfsph_b200.h declares, and fail loudly
(no CPU fallback) when no CUDA device is present.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from tests.parity import ROOT
from splishsplash_b200 import capi


def _declared():
    src = open(os.path.join(ROOT, "include", "dfsph_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfsph_b200_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_library_exports_every_declared_symbol(prec):
    lib = capi.load(prec)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dfsph_b200.h but not exported by libdfsph_b200_{prec}.so"
    assert sorted(capi.EXPORTS) == names
    assert lib.dfsph_b200_sizeof_real() == (4 if prec == "f32" else 8)
    assert b"sm_100a" in lib.dfsph_b200_version()


def test_defaults_match_reference_defaults():
    lib = capi.load("f32")
    p = capi.Params()
    lib.dfsph_b200_default_params(C.byref(p))
    # TimeStepDFSPH.cpp:28-41, Simulation.cpp:67-71, TimeManager.cpp:12
    assert (p.min_iterations, p.max_iterations, p.max_iterations_v) == (2, 100, 100)
    assert (p.max_error, p.max_error_v) == (0.01, 0.1)
    assert p.enable_divergence_solver == 1 and p.cfl_method == 1
    assert (p.cfl_factor, p.cfl_min_time_step_size, p.cfl_max_time_step_size, p.time_step_size) == (0.5, 1e-4, 5e-3, 1e-3)
    assert list(p.gravitation) == [0.0, -9.81, 0.0]
    c = capi.Config()
    lib.dfsph_b200_default_config(C.byref(c))
    assert c.kernel == capi.KERNEL_PRECOMPUTED_CUBIC and c.grad_kernel == -1 and c.particle_radius == 0.025


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from splishsplash_b200.solver import TimeStepDFSPH_B200
    with pytest.raises(capi.DFSPHError) as e:
        TimeStepDFSPH_B200("f32")
    assert e.value.code == capi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_create_rejects_bad_arguments():
    lib = capi.load("f64")
    c = capi.Config()
    lib.dfsph_b200_default_config(C.byref(c))
    ctx = C.c_void_p()
    c.kernel = 5   # ids 5, 6 are the 2-D kernels: not on this path
    assert lib.dfsph_b200_create(C.byref(c), C.byref(ctx)) == capi.ERR_UNSUPPORTED
    c.kernel = 4
    c.particle_radius = 0.0
    assert lib.dfsph_b200_create(C.byref(c), C.byref(ctx)) == capi.ERR_INVALID
    assert lib.dfsph_b200_create(None, C.byref(ctx)) == capi.ERR_INVALID

"""GPU test of the drop-in boundary: the reference's OWN Simulation / FluidModel / BoundaryModel_Akinci2012 /
TimeManager objects (compiled unmodified into oracle/_ref) run with the product's C++ solver class
TimeStepDFSPH_B200 (splishsplash_b200/host) installed in place of TimeStepDFSPH, and the result is compared with the
same stack running the reference's TimeStepDFSPH.  Skipped where oracle/_ref is not present."""
import os

import numpy as np
import pytest

from oracle import refsim
from splishsplash_b200 import scenes
from tests.parity import ROOT, TOL, conditioned_errors, dtype_of, scaled_err

pytestmark = pytest.mark.gpu


PATCHED_F64 = os.path.join(ROOT, "oracle", "_ref", "libsplish_ref_patched_f64.so")


def _run(scene, prec, b200, steps, kernel=4, grad=4, lib_path=None):
    sim = refsim.build_ref_scene(scene, prec, kernel=kernel, grad_kernel=grad, b200=b200, lib_path=lib_path)
    out = []
    try:
        name = sim.method_name
        for _ in range(steps):
            sim.step(1)
            out.append({"iters": (sim.iterations_v, sim.iterations), "h": sim.h, "time": sim.time,
                        **{f: sim.field_by_id(f) for f in ("position", "velocity", "density", "factor", "p / rho^2",
                                                           "p_v / rho^2", "advected density")}})
    finally:
        sim.destroy()
    return name, out


@pytest.mark.parametrize("prec,kernel,grad", [("f32", 4, 4), ("f64", 4, 4), ("f64", 1, 1), ("f64", 2, 3)])
def test_reference_stack_with_b200_solver(prec, kernel, grad):
    if not refsim.ref_available(prec):
        pytest.skip("oracle/_ref not present")
    sc = scenes.dam_break("small", dtype=dtype_of(prec))
    steps = 4   # free-running from the same initial state; short enough that rounding differences stay below tolerance
    name_ref, ref = _run(sc, prec, False, steps, kernel, grad)
    name_dev, dev = _run(sc, prec, True, steps, kernel, grad)
    assert name_ref == "DFSPH" and name_dev == "DFSPH_B200"
    tol = TOL[prec] * (1 if prec == "f64" else 4)   # float: four free-running steps accumulate rounding differences
    for s in range(steps):
        assert ref[s]["iters"] == dev[s]["iters"]
        assert abs(ref[s]["h"] - dev[s]["h"]) <= tol * ref[s]["h"]
        assert abs(ref[s]["time"] - dev[s]["time"]) <= 1e-6 * max(ref[s]["time"], 1e-3)
        for f in ("position", "velocity", "density", "factor", "advected density", "p_v / rho^2"):
            assert scaled_err(dev[s][f], ref[s][f]) <= tol, (s, f)
        e = scaled_err(dev[s]["p / rho^2"], ref[s]["p / rho^2"])
        if e > tol:
            h = ref[s]["h"]
            d = np.abs(dev[s]["p / rho^2"].astype(np.float64) - ref[s]["p / rho^2"].astype(np.float64))
            alpha = ref[s]["factor"].astype(np.float64) * h * h
            e = float(np.max(d[alpha > 0] / alpha[alpha > 0]))
        assert e <= tol, (s, "p / rho^2", e)


def test_registered_method_id_runs_the_b200_solver():
    """INTEGRATION.md option B end to end: the reference's Simulation.{h,cpp} with patches/register_dfsph_b200.patch
    applied (oracle/_ref/libsplish_ref_patched_f64.so) select the drop-in like a built-in method, by
    "simulationMethod" id 7; the run must match the same stack running the reference's TimeStepDFSPH (id 4)."""
    if not os.path.exists(PATCHED_F64):
        pytest.skip("patched reference library not present")
    prec, steps = "f64", 3
    sc = scenes.dam_break("small", dtype=dtype_of(prec))
    name_ref, ref = _run(sc, prec, False, steps, lib_path=PATCHED_F64)
    name_dev, dev = _run(sc, prec, "enum", steps, lib_path=PATCHED_F64)
    assert name_ref == "DFSPH" and name_dev == "DFSPH_B200"
    for s in range(steps):
        assert ref[s]["iters"] == dev[s]["iters"]
        for f in ("position", "velocity", "density", "factor", "advected density", "p_v / rho^2"):
            assert scaled_err(dev[s][f], ref[s][f]) <= TOL[prec], (s, f)


def test_neighborhood_search_facade_matches_reference_lists():
    """NeighborhoodSearch_B200 (CompactNSearch-style accessors served by the device search) against the reference stack's
    own neighbour lists: counts per particle and an order-independent checksum over (i, j) pairs, fluid and boundary."""
    import ctypes as C
    prec = "f64"
    if not refsim.ref_available(prec):
        pytest.skip("oracle/_ref not present")
    sc = scenes.dam_break("small", dtype=dtype_of(prec))
    # reference lists (z-sort disabled so that array order = host order = particle id)
    ref = refsim.build_ref_scene(sc, prec, kernel=4, enableZSort=0)
    try:
        ref.search_and_density()
        n = ref.num_particles()
        want = {}
        for pid in (0, 1):
            counts = np.empty(n, dtype=np.uint32)
            ref.lib.ref_neighbor_counts(0, pid, counts.ctypes.data)
            cs = C.c_ulonglong()
            ref.lib.ref_neighbor_checksum(0, pid, C.byref(cs))
            want[pid] = (counts, cs.value)
    finally:
        ref.destroy()
    dev = refsim.build_ref_scene(sc, prec, kernel=4, b200=True, enableZSort=0)
    try:
        for pid in (0, 1):
            counts = np.empty(n, dtype=np.uint32)
            cs = C.c_ulonglong()
            rc = dev.lib.ref_b200_neighbor_counts(pid, counts.ctypes.data, C.byref(cs))
            assert rc == 0, dev.lib.ref_last_error()
            assert np.array_equal(counts, want[pid][0])
            assert cs.value == want[pid][1]
    finally:
        dev.destroy()


def test_lazy_field_mirrors_and_bulk_accessor():
    """a step downloads none of the five DFSPH fields; the first getFct(i) of a field (the way the
    reference's state writer and exporters read, SimulatorBase.cpp:2094-2123) triggers ONE bulk download of that field;
    downloadField() is the explicit bulk accessor."""
    import ctypes as C
    prec = "f64"
    if not refsim.ref_available(prec):
        pytest.skip("oracle/_ref not present")
    sc = scenes.dam_break("small", dtype=dtype_of(prec))
    dev = refsim.build_ref_scene(sc, prec, kernel=4, b200=True)
    try:
        L = dev.lib
        L.ref_b200_download_field.argtypes = [C.c_char_p, C.c_void_p]
        dev.step(3)
        assert L.ref_b200_field_downloads() == 0                    # nobody read a DFSPH field
        factor = dev.field_by_id("factor")                          # per-particle getFct reads through the reference's FieldDescription
        assert L.ref_b200_field_downloads() == 1
        dev.field_by_id("factor")
        assert L.ref_b200_field_downloads() == 1                    # still current: no second copy
        kappa = dev.field_by_id("p / rho^2")
        assert L.ref_b200_field_downloads() == 2
        dev.step(1)
        assert L.ref_b200_field_downloads() == 2
        n = dev.num_particles()
        bulk = np.empty(n, dtype=np.float64)
        assert L.ref_b200_download_field(b"factor", bulk.ctypes.data) == 0
        assert L.ref_b200_field_downloads() == 3
        assert np.array_equal(bulk, dev.field("factor"))            # same values as the per-particle path (host array order)
        assert factor.shape == kappa.shape == (n,)
        assert L.ref_b200_download_field(b"no such field", bulk.ctypes.data) == -2
    finally:
        dev.destroy()


def test_device_resident_state_matches_host_synchronised_run():
    """setHostStateSync(false): x, v stay on the device between steps; downloadState() brings FluidModel's arrays up to
    date.  The result is bitwise the host-synchronised run (the same kernels on the same data)."""
    prec = "f64"
    if not refsim.ref_available(prec):
        pytest.skip("oracle/_ref not present")
    sc = scenes.dam_break("small", dtype=dtype_of(prec))
    out = {}
    for mode in (1, 0):
        dev = refsim.build_ref_scene(sc, prec, kernel=4, b200=True)
        try:
            dev.step(1)                                              # uploads the model
            assert dev.lib.ref_b200_set_host_sync(mode) == 0
            dev.step(5)
            if mode == 0:
                stale = dev.field_by_id("position")
                assert dev.lib.ref_b200_download_state() == 0
                assert not np.array_equal(stale, dev.field_by_id("position"))   # the host copy really was behind
            out[mode] = (dev.field_by_id("position"), dev.field_by_id("velocity"), dev.field_by_id("density"), dev.iterations, dev.time)
        finally:
            dev.destroy()
    for a, b in zip(out[0][:3], out[1][:3]):
        assert np.array_equal(a, b)
    assert out[0][3:] == out[1][3:]


def test_compactnsearch_facade_surface():
    """add_point_set / set_active / find_neighbors / z_sort / sort_field of NeighborhoodSearch_B200 used the way the
    reference uses CompactNSearch; the checks live in oracle/ref_driver.cpp::ref_b200_facade_selftest."""
    prec = "f64"
    if not refsim.ref_available(prec):
        pytest.skip("oracle/_ref not present")
    sc = scenes.dam_break("small", dtype=dtype_of(prec))
    dev = refsim.build_ref_scene(sc, prec, kernel=4, b200=True, enableZSort=0)
    try:
        dev.step(2)
        rc = dev.lib.ref_b200_facade_selftest()
        assert rc == 0, (rc, dev.lib.ref_last_error())
    finally:
        dev.destroy()

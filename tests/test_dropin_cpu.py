"""CPU tests of the registration route of the drop-in (INTEGRATION.md option B): patches/register_dfsph_b200.patch applies
to the reference's Simulation.{h,cpp}, and a reference library built with it knows "simulationMethod" 7 = DFSPH_B200
(constructing the solver object needs no GPU; stepping does, see tests/test_dropin_gpu.py)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import refsim
from splishsplash_b200 import scenes
from tests.parity import ROOT

PATCH = os.path.join(ROOT, "patches", "register_dfsph_b200.patch")
REF = "/root/reference"


def test_patch_applies_to_the_reference(tmp_path):
    if not os.path.exists(os.path.join(REF, "SPlisHSPlasH", "Simulation.cpp")):
        pytest.skip("reference sources not present")
    d = tmp_path / "SPlisHSPlasH"
    d.mkdir()
    for f in ("Simulation.h", "Simulation.cpp"):
        shutil.copy(os.path.join(REF, "SPlisHSPlasH", f), d / f)
    with open(PATCH) as fh:
        r = subprocess.run(["patch", "-p1", "-d", str(tmp_path)], stdin=fh, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    src = (d / "Simulation.cpp").read_text()
    assert "new TimeStepDFSPH_B200()" in src and "ENUM_SIMULATION_DFSPH_B200" in src
    assert "DFSPH_B200, NumSimulationMethods" in (d / "Simulation.h").read_text()


@pytest.mark.parametrize("lib,expect", [("libsplish_ref_patched_f64.so", "DFSPH_B200"), ("libsplish_ref_f64.so", None)])
def test_method_id_7(lib, expect):
    path = os.path.join(ROOT, "oracle", "_ref", lib)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not present")
    sc = scenes.dam_break("tiny", dtype=np.float64)
    sim = refsim.RefSim("f64", path)
    sim.create(sc["radius"])
    sim.add_fluid(sc["fluid_x"], None)
    try:
        if expect:
            sim.configure_by_method_id(7, 4)
            assert sim.method_name == expect
        else:   # the unpatched reference maps unknown ids to DFSPH (Simulation.cpp:538-539)
            with pytest.raises(RuntimeError, match="not DFSPH_B200"):
                sim.configure_by_method_id(7, 4)
    finally:
        sim.destroy()


def test_pybind_module_of_the_drop_in_imports():
    """Next-row f4 (pybind half): splishsplash_b200/host/DFSPH_B200Module.cpp, the pySPlisHSPlasH-style module of the
    drop-in (mirror of pySPlisHSPlasH/DFSPHModule.cpp:50-66), built against the reference stack by oracle/Makefile."""
    import importlib.util
    path = os.path.join(ROOT, "oracle", "_ref", "pydfsph_b200_check.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/pydfsph_b200_check.so not present")
    spec = importlib.util.spec_from_file_location("pydfsph_b200_check", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    cls = m.TimeStepDFSPH_B200
    assert cls.METHOD_NAME == "DFSPH_B200"
    ts = cls(os.path.join(ROOT, "splishsplash_b200"))     # loads the CUDA library (no device needed until the first step)
    try:
        assert isinstance(ts, m.TimeStep) and ts.getMethodName() == "DFSPH_B200" and ts.getNumIterations() == 0
        ts.init()                                          # GenParam registration: the static handles become valid ids
        handles = [cls.SOLVER_ITERATIONS, cls.MIN_ITERATIONS, cls.MAX_ITERATIONS, cls.MAX_ERROR, cls.SOLVER_ITERATIONS_V,
                   cls.MAX_ITERATIONS_V, cls.MAX_ERROR_V, cls.USE_DIVERGENCE_SOLVER]
        assert all(h >= 0 for h in handles) and len(set(handles)) == len(handles)
        ts.setSyncAllFields(False)
    finally:
        # the constructor made Simulation::getCurrent() create a Simulation inside libsplish_ref_f64.so, which the other
        # tests of this process share: release the solver, then the Simulation
        del ts
        import ctypes
        ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsplish_ref_f64.so")).ref_destroy()

"""TEST INFRASTRUCTURE ONLY.  Driver for the plain C++ restatement oracle/dfsph_oracle.cpp (liboracle_{f32,f64}.so).
It exports the same harness ABI as oracle/ref_driver.cpp, so it is driven through refsim.RefSim."""
from __future__ import annotations

import os

from .refsim import RefSim, build_ref_scene

_HERE = os.path.dirname(os.path.abspath(__file__))


def port_lib_path(precision: str) -> str:
    return os.path.join(_HERE, f"liboracle_{precision}.so")


def port_available(precision: str) -> bool:
    return os.path.exists(port_lib_path(precision))


def build_port_scene(scene, precision="f64", kernel=4, grad_kernel=None, **params) -> RefSim:
    return build_ref_scene(scene, precision, kernel=kernel, lib_path=port_lib_path(precision), grad_kernel=grad_kernel, **params)

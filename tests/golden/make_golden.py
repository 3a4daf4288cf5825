"""Generates the golden fixtures in this directory from the REFERENCE ITSELF (oracle/_ref = the reference's unmodified
DFSPH sources compiled by oracle/Makefile; needs /root/reference at build time, not at run time).

    python tests/golden/make_golden.py

Each fixture stores, for a sequence of steps, the exact input state of the step (x, v, kappa, kappa_v by particle id,
time step size) and every per-step output field of the reference, plus the neighbour sets of the initial state and the
boundary volumes.  Tests replay each step from the stored input on the oracle port (CPU) and on the CUDA library (GPU).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refsim  # noqa: E402
from splishsplash_b200 import scenes  # noqa: E402
from tests.parity import STEP_FIELDS, neighbor_sets_by_id  # noqa: E402

FIXTURES = [  # name, scene factory, precision, kernel, steps, params
    ("dambreak_tiny_f32_k4", lambda dt: scenes.dam_break("tiny", dtype=dt), "f32", 4, 8, {}),
    ("dambreak_tiny_f64_k4", lambda dt: scenes.dam_break("tiny", dtype=dt), "f64", 4, 8, {}),
    ("dambreak_tiny_f64_k0", lambda dt: scenes.dam_break("tiny", dtype=dt), "f64", 0, 3, {}),
    ("rwstate_f64_k4", lambda dt: scenes.rw_state_scene(dtype=dt), "f64", 4, 2,
     dict(timeStepSize=0.005, cflFactor=1.0, maxError=0.05)),   # data/Scenes/ReadWriteStateTest.json settings
    ("dambreak_tiny_visc_f64_k4", lambda dt: _sheared(scenes.dam_break("tiny", dtype=dt)), "f64", 4, 3,
     dict(viscosityMethod=1, viscosity=0.05, viscosityBoundary=0.02)),   # next-row f1: Viscosity_Standard
    ("dambreak_tiny_visc_f32_k4", lambda dt: _sheared(scenes.dam_break("tiny", dtype=dt)), "f32", 4, 3,
     dict(viscosityMethod=1, viscosity=0.05, viscosityBoundary=0.0)),
    # next-row f2: the other selectable kernels of the scalar (double) build; "gradKernel" is popped from the params
    ("dambreak_tiny_f64_k1", lambda dt: scenes.dam_break("tiny", dtype=dt), "f64", 1, 2, {}),              # Wendland quintic C2
    # edge cases: a coincident pair (|r| = 0: W(0) counted, gradient 0), an isolated particle without any neighbour
    # (density = V W(0) rho0, factor 0) and one that only sees boundary particles
    ("edge_tiny_f64_k4", lambda dt: _edge(scenes.dam_break("tiny", dtype=dt)), "f64", 4, 3, {}),
    ("edge_tiny_f32_k4", lambda dt: _edge(scenes.dam_break("tiny", dtype=dt)), "f32", 4, 3, {}),
    ("dambreak_tiny_f64_k2g3", lambda dt: scenes.dam_break("tiny", dtype=dt), "f64", 2, 2, dict(gradKernel=3)),   # Poly6 / Spiky
]


def _sheared(sc):
    v = np.zeros_like(sc["fluid_x"])
    v[:, 0] = 1.5 * np.sin(6.0 * sc["fluid_x"][:, 1])
    sc["fluid_v"] = v
    return sc


def _edge(sc):
    x = sc["fluid_x"]
    tmin, tmax = np.asarray(sc["tank_min"]), np.asarray(sc["tank_max"])
    mid = 0.5 * (tmin + tmax)
    extra = np.array([
        x[len(x) // 2],                                              # coincident with an interior particle
        [mid[0] + 0.31, tmax[1] - 0.333, mid[2]],                    # mid-air, no neighbour at all
        [tmax[0] - 0.04, tmax[1] - 0.04, tmax[2] - 0.04],            # in a top corner: boundary neighbours only
    ], dtype=x.dtype)
    sc["fluid_x"] = np.ascontiguousarray(np.concatenate([x, extra]))
    return sc


def main():
    only = set(sys.argv[1:])   # optional: names of the fixtures to (re)generate
    for name, factory, prec, kernel, steps, params in FIXTURES:
        if only and name not in only:
            continue
        dt = np.float32 if prec == "f32" else np.float64
        sc = factory(dt)
        params = dict(params)
        grad_kernel = int(params.pop("gradKernel", kernel))
        sim = refsim.build_ref_scene(sc, prec, kernel=kernel, grad_kernel=grad_kernel, **params)
        out = {"fluid_x": sc["fluid_x"], "boundary_x": sc["boundary_x"], "radius": np.float64(sc["radius"]),
               **({"fluid_v": sc["fluid_v"]} if sc.get("fluid_v") is not None else {}),
               "kernel": np.int32(kernel), "grad_kernel": np.int32(grad_kernel), "steps": np.int32(steps)}
        for k, v in params.items():
            out["param_" + k] = np.float64(v)
        bx, bV = sim.boundary(0)
        key = lambda a: [tuple(r) for r in a.tolist()]
        d = dict(zip(key(bx), bV.tolist()))
        out["boundary_V"] = np.array([d[k] for k in key(np.asarray(sc["boundary_x"], dtype=dt))], dtype=dt)
        sim.search_and_density()
        rid = sim.ids()
        c, o, i = sim.neighbors(0, 0)
        nc, nl = neighbor_sets_by_id(c, o, i, rid, rid)
        out["nbr_f_counts"], out["nbr_f_ids"] = nc, nl
        ins = {k: n for n, k in enumerate(key(np.asarray(sc["boundary_x"], dtype=dt)))}
        rmap = np.array([ins[k] for k in key(bx)], dtype=np.uint32)
        c, o, i = sim.neighbors(0, 1)
        nc, nl = neighbor_sets_by_id(c, o, i, rid, rmap)
        out["nbr_b_counts"], out["nbr_b_ids"] = nc, nl
        out["density0"] = sim.field_by_id("density")
        for s in range(steps):
            for f in ("position", "velocity", "p / rho^2", "p_v / rho^2"):
                out[f"in{s}_{f}"] = sim.field_by_id(f)
            out[f"in{s}_h"] = np.float64(sim.h)
            sim.step(1)
            for f in STEP_FIELDS:
                out[f"out{s}_{f}"] = sim.field_by_id(f)
            out[f"out{s}_h"] = np.float64(sim.h)
            out[f"out{s}_iters"] = np.array([sim.iterations_v, sim.iterations], dtype=np.int32)
        sim.destroy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "N =", len(sc["fluid_x"]), "Nb =", len(sc["boundary_x"]), os.path.getsize(path) // 1024, "KiB",
              "iters", [out[f"out{s}_iters"].tolist() for s in range(steps)])


if __name__ == "__main__":
    main()

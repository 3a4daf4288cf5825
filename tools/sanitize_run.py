"""Small runs for compute-sanitizer (memcheck / racecheck / initcheck are slow: keep the scenes small).
    compute-sanitizer --tool memcheck python tools/sanitize_run.py f32"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from splishsplash_b200 import scenes
from splishsplash_b200.solver import build_b200_scene
prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
dt = np.float32 if prec == "f32" else np.float64
runs = [("small dam break", scenes.dam_break("small", dtype=dt), {}, 6),
        ("double dam break variant", scenes.double_dam_break_scene(dt), scenes.DOUBLE_DAM_BREAK_PARAMS, 40),
        ("dense block (rw state scene)", scenes.rw_state_scene(dt), dict(timeStepSize=0.005, cflFactor=1.0, maxError=0.05), 6)]
mode = sys.argv[2] if len(sys.argv) > 2 else "all"     # all | nohost | hostfirst | presized (stage sized before the first graph exists)
if mode != "all":
    runs = runs[:1]
presize = mode == "presized"
for name, sc, par, steps in runs:
    ts = build_b200_scene(sc, prec, **par)
    it = []
    if presize:
        n0 = len(sc["fluid_x"])
        px, pv, pr = ts.pinned((n0, 3)), ts.pinned((n0, 3)), ts.pinned((n0,))
        px[:] = ts.field("position"); pv[:] = ts.field("velocity")
        ts.step_host(px, pv, pr)
    for _ in range(0 if mode == "hostfirst" else steps):
        st = ts.step(1)
        it.append((int(st.iterations_v), int(st.iterations)))
    # host-buffer path and field access too
    n = len(sc["fluid_x"])
    hx, hv, hr = ts.pinned((n, 3)), ts.pinned((n, 3)), ts.pinned((n,))
    hx[:] = ts.field("position"); hv[:] = ts.field("velocity")
    if mode == "allocmid":      # new allocations after the loop graphs were instantiated, then plain (graph) steps again
        for _ in range(2):
            ts.step(1)
        print("plain steps after the allocations ok", flush=True)
    elif mode != "nohost":
        ts.step_host(hx, hv, hr)
        print("step_host ok", flush=True)
        ts.step_host(hx, hv, hr)
    c, o, i = ts.neighbors(0)
    print(name, prec, "N", n, "iterations", it[-3:], "pairs", len(i), flush=True)
    ts.close()
print("done")

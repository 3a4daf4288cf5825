"""Tuning aid (run under gpurun): per-kernel-class CUDA-event timings of alternative builds of the library.

    python tools/variant_bench.py [--particles 10M] [--warm 20] [--steps 10] [--precision f32] lib1.so lib2.so ...

Every library runs the same dam-break block from rest: `warm` untimed steps, then `steps` profiled steps."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from splishsplash_b200 import scenes, capi
from splishsplash_b200.solver import build_b200_scene

ap = argparse.ArgumentParser()
ap.add_argument("--particles", default="10M")
ap.add_argument("--warm", type=int, default=20)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--precision", default="f32")
ap.add_argument("libs", nargs="+")
a = ap.parse_args()
sc = scenes.dam_break(a.particles, dtype=np.float32 if a.precision == "f32" else np.float64)
par = dict(minIterations=2, maxIterations=100, maxError=0.01, maxIterationsV=100, maxErrorV=0.1, enableDivergenceSolver=1,
           cflMethod=1, cflFactor=0.5, cflMaxTimeStepSize=0.005, timeStepSize=0.001)
for lib in a.libs:
    os.environ[f"DFSPH_B200_LIB_{a.precision.upper()}"] = os.path.abspath(lib)
    capi._LIBS.clear()
    t0 = time.time()
    ts = build_b200_scene(sc, a.precision, **par)
    for _ in range(a.warm):
        ts.step(1)
    ts.synchronize()
    # window 1: the product path (solver loops as CUDA graphs, no per-kernel events)
    ts.timer_start()
    its = []
    for _ in range(a.steps):
        st = ts.step(1)
        its.append((st.iterations_v, st.iterations))
    ms = ts.timer_stop()
    # window 2: per-kernel-class CUDA events (host-driven loops)
    ts.set_profiling(True)
    ts.timer_start()
    its2 = []
    for _ in range(a.steps):
        st = ts.step(1)
        its2.append((st.iterations_v, st.iterations))
    ms2 = ts.timer_stop()
    prof = ts.profile()
    ts.set_profiling(False)
    out = {"lib": os.path.basename(lib), "ms_per_step": ms / a.steps, "iters": [float(np.mean([i[0] for i in its])), float(np.mean([i[1] for i in its]))],
           "profiled_ms_per_step": ms2 / a.steps, "profiled_iters": [float(np.mean([i[0] for i in its2])), float(np.mean([i[1] for i in its2]))],
           "ms_per_launch": {k: (p[0] / p[1] if p[1] else None) for k, p in prof.items()}, "setup_s": time.time() - t0}
    print(json.dumps(out), flush=True)
    ts.close()

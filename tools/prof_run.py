"""Profiling driver (run under ncu via gpurun): a few DFSPH steps of a named dam-break block."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from splishsplash_b200 import scenes
from splishsplash_b200.solver import build_b200_scene
prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
name = sys.argv[2] if len(sys.argv) > 2 else "1M"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sc = scenes.dam_break(name, dtype=np.float32 if prec == "f32" else np.float64)
ts = build_b200_scene(sc, prec)
for s in range(steps):
    st = ts.step(1)
    print(s, st.iterations_v, st.iterations, st.ms_search, st.ms_solver, flush=True)
ts.close()

"""Run the same scene twice (two contexts, same process) and compare every field bit for bit after each step."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from splishsplash_b200 import scenes
from splishsplash_b200.solver import build_b200_scene
prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
name = sys.argv[2] if len(sys.argv) > 2 else "small"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dt = np.float32 if prec == "f32" else np.float64
sc = scenes.dam_break(name, dtype=dt)
a = build_b200_scene(sc, prec); b = build_b200_scene(sc, prec)
F = ["position", "velocity", "density", "factor", "advected density", "p / rho^2", "p_v / rho^2", "pressure acceleration"]
for s in range(steps):
    sa = a.step(1); sb = b.step(1)
    bad = [f for f in F if not np.array_equal(a.field(f), b.field(f))]
    ida, idb = a.field("id", by_id=False), b.field("id", by_id=False)
    print(s, (sa.iterations_v, sa.iterations), (sb.iterations_v, sb.iterations), "order equal", np.array_equal(ida, idb), "differing fields", bad, flush=True)
    for f in bad[:3]:
        x, y = a.field(f).astype(np.float64), b.field(f).astype(np.float64)
        d = np.abs(x - y); print("    ", f, "max diff", d.max(), "count", int((d > 0).sum()))
a.close(); b.close()

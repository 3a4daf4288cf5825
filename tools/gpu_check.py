"""Development aid (run under gpurun): parity of the CUDA path against the oracle on small scenes + a quick timing."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splishsplash_b200 import scenes  # noqa: E402
from splishsplash_b200.solver import build_b200_scene  # noqa: E402
from tests.parity import compare_step, dtype_of  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"):
        for prec in ("f32", "f64"):
            for name in ("tiny", "small"):
                for kernel in ((4, 0) if prec == "f64" else (4,)):
                    sc = scenes.dam_break(name, dtype=dtype_of(prec))
                    r = compare_step(prec, sc, steps=6, kernel=kernel)
                    print(f"[{prec} {name} kernel={kernel}] {r['summary']}")
                    print("   ", {k: f"{v:.2e}" for k, v in r["max_err"].items()}, flush=True)
    if what in ("all", "perf"):
        for prec, name in (("f32", "1M"), ("f64", "1M")):
            sc = scenes.dam_break(name, dtype=dtype_of(prec))
            t0 = time.time()
            ts = build_b200_scene(sc, prec)
            print(f"[{prec} {name}] setup {time.time()-t0:.2f}s N={ts.num_particles} Nb={ts.num_boundary_particles}")
            for s in range(30):
                st = ts.step(1)
                if s % 5 == 0 or s > 26:
                    print(f"  step {s}: itV={st.iterations_v} it={st.iterations} h={st.time_step_size:.5f} search={st.ms_search:.3f}ms "
                          f"solver={st.ms_solver:.3f}ms launches={st.gpu_launches} maxnbr={st.max_neighbors}", flush=True)
            ts.close()


if __name__ == "__main__":
    main()

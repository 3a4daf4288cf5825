"""BASELINE config 1 (data/Scenes/DoubleDamBreak.json) as the labelled Akinci2012 variant: the same run on the B200 library
and on the reference's CPU build (oracle/_ref), timed side by side.  Run under gpurun:

    python tools/config1_variant.py [f32|f64] [steps] > gpurun_out/r2_config1_variant.json

4732 particles: the GPU step is launch-latency bound (one CUDA graph per solver loop), the CPU step fits the caches."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from splishsplash_b200 import scenes
from splishsplash_b200.solver import build_b200_scene
from oracle import refsim

prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
sc = scenes.double_dam_break_scene(np.float32 if prec == "f32" else np.float64)
par = scenes.DOUBLE_DAM_BREAK_PARAMS
n = len(sc["fluid_x"])

dev = build_b200_scene(sc, prec, **par)
for _ in range(20):
    dev.step(1)
dev.synchronize()
dev.timer_start()
it_d = []
for _ in range(steps):
    st = dev.step(1)
    it_d.append((int(st.iterations_v), int(st.iterations)))
ms_dev = dev.timer_stop() / steps
dev.close()

out = {"workload": "DoubleDamBreak.json geometry + settings, Akinci2012 particle box instead of the Bender2019 volume map (variant)",
       "particles": n, "precision": prec, "steps": steps, "warmup": 20,
       "b200": {"ms_per_step": ms_dev, "particle_updates_per_s": n / ms_dev * 1e3,
                "mean_iterations": [float(np.mean([i[0] for i in it_d])), float(np.mean([i[1] for i in it_d]))]}}
if refsim.ref_available(prec):
    ref = refsim.build_ref_scene(sc, prec, **par)
    ref.step(20)
    it_r = []
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.step(1)
        it_r.append((ref.iterations_v, ref.iterations))
    ms_ref = (time.perf_counter() - t0) * 1e3 / steps
    ref.destroy()
    out["reference_cpu"] = {"ms_per_step": ms_ref, "particle_updates_per_s": n / ms_ref * 1e3, "threads": os.cpu_count(),
                            "mean_iterations": [float(np.mean([i[0] for i in it_r])), float(np.mean([i[1] for i in it_r]))]}
print(json.dumps(out))

"""Phase trace of the weak-scaling bench scene (DFSPH_B200_TRACE=1): GPU-timeline time between phase marks, per rank.
    DFSPH_B200_TRACE=1 python tools/trace_run.py                      (1 GPU)
    DFSPH_B200_TRACE=1 torchrun --nproc-per-node N tools/trace_run.py (N slabs, 10 M particles each)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from splishsplash_b200 import scenes, parallel
from splishsplash_b200.solver import build_b200_scene
warm, steps, name = 5, 20, (sys.argv[1] if len(sys.argv) > 1 else "10M")
world = int(os.environ.get("WORLD_SIZE", "1"))
par = dict(minIterations=2, maxIterations=100, maxError=0.01, maxIterationsV=100, maxErrorV=0.1, enableDivergenceSolver=1,
           cflMethod=1, cflFactor=0.5, cflMaxTimeStepSize=0.005, timeStepSize=0.001)
if world > 1:
    import torch, torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = scenes.dam_break_weak(rank, world, name)
    ts = parallel.build_b200_slab(sc, "f32", rank, world, device=local, **par)
else:
    rank = 0
    ts = build_b200_scene(scenes.dam_break(name), "f32", **par)
for _ in range(warm):
    ts.step(1)
ts.synchronize()
if world > 1:
    dist.barrier()
ts.timer_start()
for _ in range(steps):
    st = ts.step(1)
ms = ts.timer_stop()
print(f"rank {rank}: {ms / steps:.3f} ms/step, iterations ({st.iterations_v}, {st.iterations})", flush=True)
ts.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

"""Multi-GPU parity check (run with torchrun --nproc-per-node 2): a dam-break block split into x-slabs must reproduce
the single-GPU run of the same global scene: per-field agreement by particle id and identical iteration counts."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splishsplash_b200 import scenes, parallel  # noqa: E402
from splishsplash_b200.solver import build_b200_scene  # noqa: E402
from tests.parity import scaled_err, dtype_of  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
    name = sys.argv[2] if len(sys.argv) > 2 else "small"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    axis = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    host = int(sys.argv[5]) if len(sys.argv) > 5 else 0   # 1: every multi-GPU step through dfsph_b200_step_host (device-order rows)
    resync = int(sys.argv[6]) if len(sys.argv) > 6 else 0  # 1: like tests/parity.py::compare_step -- every step starts from the single-GPU run's state
                                                           #    (by particle id) and every step's fields are compared at the documented tolerance
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = scenes.dam_break(name, dtype=dtype_of(prec))
    # give the block an initial velocity towards +x so that particles migrate across the slab faces
    v = np.zeros_like(sc["fluid_x"]); v[:, axis] = 0.8
    sc["fluid_v"] = v
    if os.environ.get("FUZZ_SEED"):   # randomised variant: jittered, thinned, stirred block (the same on every rank)
        rng = np.random.default_rng(int(os.environ["FUZZ_SEED"]))
        x = sc["fluid_x"].astype(np.float64)
        d = 2.0 * sc["radius"]
        x = x + float(rng.choice([0.05, 0.2, 0.35])) * d * (rng.random(x.shape) - 0.5)
        keep = rng.random(len(x)) >= float(rng.choice([0.0, 0.1, 0.4]))
        x = x[keep]
        vv = v[keep].astype(np.float64) + float(rng.choice([0.0, 0.5, 2.0])) * (rng.random(x.shape) - 0.5)
        sc["fluid_x"] = np.ascontiguousarray(x.astype(dtype_of(prec)))
        sc["fluid_v"] = np.ascontiguousarray(vv.astype(dtype_of(prec)))
    mine = parallel.select_slab(sc, rank, world, axis=axis)
    # boundary volumes must come from the global boundary (a per-rank subset would change V near the cut)
    single = build_b200_scene(sc, prec, device=local) if True else None
    bV = single.boundary_volume()
    ts = parallel.build_b200_slab(mine, prec, rank, world, device=local, boundary_V=bV[mine["boundary_keep"]])
    fields = ["position", "velocity", "density", "factor", "p / rho^2", "p_v / rho^2", "advected density"]
    ok = True
    iters_m, iters_s = [], []
    if host:
        cap = ts.capacity
        hx, hv, hrho = ts.pinned((cap, 3)), ts.pinned((cap, 3)), ts.pinned((cap,))
        m = ts.num_particles
        hx[:m] = ts.field("position", by_id=False)
        hv[:m] = ts.field("velocity", by_id=False)
    step_worst = {}
    for s in range(steps):
        if resync and s > 0:
            # slab rows <- the single-GPU run's state of the particles this rank owns now (ownership follows the migration)
            own = ts.field("id", by_id=False)
            for f in ("position", "velocity", "p / rho^2", "p_v / rho^2"):
                ts.set_field(f, single.field(f)[own], by_id=False)
            ts.setValue("timeStepSize", single.h)
        st = ts.step_host(hx, hv, hrho) if host else ts.step(1)
        ss = single.step(1)
        iters_m.append((st.iterations_v, st.iterations)); iters_s.append((ss.iterations_v, ss.iterations))
        if resync:
            own = ts.field("id", by_id=False)
            for f in fields:
                ref = single.field(f)[own]
                e = scaled_err(ts.field(f, by_id=False), ref, scale=float(np.max(np.abs(single.field(f)))) if len(own) else None)
                step_worst[f] = max(step_worst.get(f, 0.0), e)
    ids = ts.field("id", by_id=False)
    local_fields = {f: ts.field(f, by_id=False) for f in fields}
    host_ok = True
    if host:   # the host buffers hold exactly the device state of the rows this rank owns after the last step
        m = ts.num_particles
        host_ok = (np.array_equal(hx[:m], local_fields["position"]) and np.array_equal(hv[:m], local_fields["velocity"])
                   and np.array_equal(hrho[:m], local_fields["density"]) and st.num_particles == m)
        if not host_ok:
            print(f"rank {rank}: step_host buffers differ from the device state")
    oks = [None] * world
    dist.all_gather_object(oks, host_ok)
    host_ok = all(oks)
    gathered = [None] * world
    dist.all_gather_object(gathered, (ids, local_fields, ts.num_particles))
    step_worsts = [None] * world
    dist.all_gather_object(step_worsts, step_worst)
    if rank == 0:
        n = len(sc["fluid_x"])
        counts = [g[2] for g in gathered]
        all_ids = np.concatenate([g[0] for g in gathered])
        assert len(all_ids) == n and len(np.unique(all_ids)) == n, ("particles lost or duplicated", counts, n)
        worst = {}
        for f in fields:
            ref = single.field(f)
            got = np.empty_like(ref)
            for g in gathered:
                got[g[0]] = g[1][f]
            worst[f] = scaled_err(got, ref)
        if resync:   # per-step comparison from identical states: the documented per-step tolerances hold
            tol = 1e-10 if prec == "f64" else 1e-4
            worst = {f: max(sw.get(f, 0.0) for sw in step_worsts) for f in fields}
        else:
            tol = (1e-8 if prec == "f64" else 5e-4)   # free-running for `steps` steps: rounding differences accumulate
        same_iters = iters_m == iters_s
        ok = same_iters and host_ok and all(e <= tol for e in worst.values())
        print(f"[{prec} {name} world={world} axis={axis} host={host} resync={resync}] owned per rank {counts} steps={steps} iters equal={same_iters} "
              f"worst={max(worst.items(), key=lambda kv: kv[1])} ok={ok}")
        print("   ", {k: f"{e:.2e}" for k, e in worst.items()})
        if not same_iters:
            print("    multi:", iters_m, "\n    single:", iters_s)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    ts.close(); single.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

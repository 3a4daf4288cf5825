#!/usr/bin/env python
"""Benchmark of the DFSPH hot path (per-step neighbourhood search + DFSPH pressure-solver loop).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--particles 10M] [--precision f32] [--impl reference]

Metric (BASELINE.json): DFSPH particle-updates/s = fluid particles x steps / device seconds, on a synthetic dam-break
block.  One "step" = one TimeStepDFSPH::step() (search + divergence solve + pressure solve + advection).
`value`   : state resident in HBM, timed with CUDA events on the library's stream (max over ranks).
`e2e`     : the same steps through dfsph_b200_step_host with pinned HOST buffers (H2D of x,v and D2H of x,v,density
            inside the timed region).
`roofline`: dominant kernel class, algorithmic bytes / its CUDA-event launch duration / measured HBM peak.
`cpu_baseline`: the reference's own DFSPH sources (oracle/_ref, built by oracle/Makefile; neighbour search = our
            CompactNSearch-compatible stand-in) on the box's host cores, on a bounded sample of the same workload
            (1 M-particle block, the same W warm-up steps, then up to 10 timed steps).
`--impl reference` prints that CPU run as its own JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DFSPH particle-updates/s"
UNIT = "particle-updates/s"

# algorithmic bytes per fluid particle per launch, in units of R = sizeof(Real) (+ constant bytes): SURVEY.md 8d
ALGO_BYTES = {
    "sort": (19, 32),          # NS key+perm 3R+16  +  reorder of persistent fields 16R+16
    "build_neighbors": (3, 0),  # reads x once; the table it writes is an implementation artefact, not compulsory traffic
    "init_sweep": (19, 0),     # K1 4R + K2 4R + K3 11R (fused)
    "accel": (7, 0),           # pass A
    "jacobi_div": (10, 0),     # pass B
    "jacobi_press": (10, 0),
    "div_final": (28, 0),      # K5 16R + K6 12R (fused)
    "press_init": (12, 0),     # K7
    "press_final": (23, 0),    # K9 14R + K10 9R (fused)
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/),
# keyed by (kernel class, precision, particles); None when no capture exists for that configuration
NCU_TRAFFIC = {
    # profiles/r1_ncu_10M_f32_v2_summary.csv (ncu --set full, tools/prof_run.py f32 10M 2)
    ("jacobi_press", "f32", "10M"): 2.0707e9,
    ("jacobi_div", "f32", "10M"): 2.0707e9,
    ("accel", "f32", "10M"): 1.6501e9,
    ("build_neighbors", "f32", "10M"): 1.9677e9,
    ("init_sweep", "f32", "10M"): 2.3788e9,
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split("\n")[0]
                f = [t.strip() for t in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def workload_name(particles, precision, counts):
    return (f"synthetic dam-break {particles} particles ({counts[0]}x{counts[1]}x{counts[2]} lattice), DFSPH "
            f"{'fp32' if precision == 'f32' else 'fp64'}, cubic kernel, Akinci2012 box boundary, warm start, divergence solver on")


def solver_params():
    # BASELINE.md section 3 "Solver settings"
    return dict(minIterations=2, maxIterations=100, maxError=0.01, maxIterationsV=100, maxErrorV=0.1,
                enableDivergenceSolver=1, cflMethod=1, cflFactor=0.5, cflMaxTimeStepSize=0.005, timeStepSize=0.001)


def run_reference_cpu(sample, precision, steps, warmup, budget_s=25.0):
    """Time the reference's own CPU DFSPH (oracle/_ref; falls back to the C++ restatement) on `sample` particles."""
    # all the host threads this process may use -- torchrun exports OMP_NUM_THREADS=1 to its workers, which would turn
    # the reference arm at N>1 into a single-thread run
    host_threads = int(os.environ.get("BENCH_CPU_THREADS", "0")) or len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(host_threads)
    from oracle import refsim, portsim
    from splishsplash_b200 import scenes
    dt = np.float32 if precision == "f32" else np.float64
    sc = scenes.dam_break(sample, dtype=dt)
    par = solver_params()
    if refsim.ref_available(precision):
        sim = refsim.build_ref_scene(sc, precision, kernel=4, **par)
        kind = "reference"
    else:
        sim = portsim.build_port_scene(sc, precision, kernel=4, **par)
        kind = "port"
    n = sim.num_particles()
    sim.lib.ref_set_num_threads(host_threads)
    cores = sim.lib.ref_num_threads()
    sim.step(max(warmup, 1))
    sim.reset_step_seconds()
    done = 0
    t0 = time.time()
    iters = []
    while done < steps:
        sim.step(1)
        done += 1
        iters.append((sim.iterations_v, sim.iterations))
        if time.time() - t0 > budget_s and done >= 3:
            break
    secs = sim.step_seconds
    timers = {k: sim.timer_ms(k) for k in ("neighborhood_search", "precomputeValues", "computeDFSPHFactor", "divergenceSolve", "pressureSolve")}
    sim.destroy()
    value = n * done / secs
    return {"value": value, "unit": UNIT, "cores": int(cores), "kind": kind,
            "sample": f"dam-break {sample} block ({n} particles), {done} steps after {max(warmup, 1)} warm-up steps, "
                      f"{'float+AVX' if precision == 'f32' else 'double scalar'} build, OMP threads={cores}; neighbour search = "
                      "CompactNSearch-compatible stand-in (oracle/standin), not CompactNSearch a9ab7c71",
            "ms_per_step": 1000.0 * secs / done, "steps": done,
            "mean_iterations": [float(np.mean([i[0] for i in iters])), float(np.mean([i[1] for i in iters]))],
            "ref_timers_ms": timers}


PROGRESS = {"phase": "start", "step": -1}


def start_watchdog(seconds, rank):
    """A hung exchange (a rank spinning on a peer that died) must not hold N GPUs until the caller's limit: after
    `seconds` the process reports where it was and exits; CUDA tears the context (and any spinning kernel) down."""
    def run():
        time.sleep(seconds)
        sys.stderr.write(f"bench.py watchdog: rank {rank} still in phase '{PROGRESS['phase']}' step {PROGRESS['step']} after "
                         f"{seconds:.0f} s -- aborting\n")
        sys.stderr.flush()
        os._exit(3)
    threading.Thread(target=run, daemon=True).start()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--particles", default="10M")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", default="1M")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps (the e2e leg replays the window of the device-resident run)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = one block of --particles per GPU (default); strong = one block of --particles cut into N slabs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from splishsplash_b200 import scenes
    counts = scenes.NAMED_BLOCKS[args.particles]
    R = 4 if args.precision == "f32" else 8
    cfg = {"workload": workload_name(args.particles, args.precision, counts), "particles_per_gpu": int(np.prod(counts)),
           "precision": args.precision, "l2": "particle state per GPU (>= 0.8 GB at 10M) exceeds the 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        # the same W warm-up steps as the B200 arm (the cost of a DFSPH step grows with the phase of the collapse: 2
        # pressure iterations in the first steps, ~10 after 30 steps at 1 M particles), then up to K timed steps
        r = run_reference_cpu(args.cpu_sample, args.precision, args.steps, args.warmup, budget_s=40.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": cfg,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "mean_iterations": r["mean_iterations"], "ref_timers_ms": r["ref_timers_ms"]}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    # rank 0 prints ONE JSON line on stdout: everything libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        start_watchdog(float(os.environ.get("BENCH_WATCHDOG_S", 300 + 0.5 * (args.steps + args.warmup))), rank)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the DFSPH hot path has no CPU fallback)")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    from splishsplash_b200.solver import build_b200_scene

    dt = np.float32 if args.precision == "f32" else np.float64
    if world == 1:
        sc = scenes.dam_break(args.particles, dtype=dt)
        ts = build_b200_scene(sc, args.precision, device=local_rank, **solver_params())
    else:
        # weak scaling: `world` blocks of the named size side by side in one tank (strong: one block cut into `world`
        # slabs), one slab per GPU; migration and ghost exchange happen inside the library (csrc/multi_gpu.cuh)
        from splishsplash_b200 import parallel
        if args.scaling == "weak":
            sc = scenes.dam_break_weak(rank, world, args.particles, dtype=dt)
        else:
            sc = scenes.dam_break_slab(rank, world, args.particles, dtype=dt)
        ts = parallel.build_b200_slab(sc, args.precision, rank, world, device=local_rank, **solver_params())
    n = ts.num_particles
    n_global = n if world == 1 else sc["global_particles"]
    cfg["particles_per_gpu"] = int(n_global // world)
    cfg["particles_total"] = int(n_global)

    def barrier():
        ts.synchronize()
        if world > 1:
            dist.barrier()

    PROGRESS["phase"] = "warm-up (device-resident)"
    for k in range(args.warmup):
        PROGRESS["step"] = k
        ts.step(1)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None   # one nvidia-smi poller per job, not per rank
    if sampler:
        sampler.start()
    ts.set_profiling(True)
    ts.timer_start()
    launches = 0
    iters = []
    ms_search = ms_solver = 0.0
    PROGRESS["phase"] = "timed steps (device-resident)"
    for k in range(args.steps):
        PROGRESS["step"] = k
        st = ts.step(1)
        launches += st.gpu_launches
        iters.append((st.iterations_v, st.iterations))
        ms_search += st.ms_search
        ms_solver += st.ms_solver
    ms = ts.timer_stop()
    clocks = sampler.stop() if sampler else None
    prof = ts.profile()
    ts.set_profiling(False)
    barrier()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- end-to-end through host buffers (pinned): H2D of x, v and D2H of x, v, density inside the timed region
    e2e_iters = []
    if world == 1:
        # same workload, same window: a fresh context replays warm-up + timed steps, every step through step_host
        ts.close()
        ts = build_b200_scene(sc, args.precision, device=local_rank, **solver_params())
        n = ts.num_particles
        x = ts.pinned((n, 3))
        v = ts.pinned((n, 3))
        rho = ts.pinned((n,))
        x[:] = sc["fluid_x"]
        v[:] = sc["fluid_v"] if sc.get("fluid_v") is not None else 0
        e2e_steps = args.e2e_steps if args.e2e_steps > 0 else args.steps
        e2e_warm = args.warmup

        def e2e_step():
            st = ts.step_host(x, v, rho)
            e2e_iters.append((st.iterations_v, st.iterations))
    else:
        # multi-GPU: ids are global, so the host buffers are in this rank's device order (capacity rows); step_host
        # uploads the owned rows, steps (migration + ghost exchange inside) and downloads the rows owned afterwards.
        # Same workload, same window: every rank re-submits its initial slab on the same communicator and replays
        # warm-up + timed steps through step_host.
        PROGRESS["phase"] = "re-submitting the initial slab"
        barrier()
        ts.setValue("timeStepSize", solver_params()["timeStepSize"])
        ts.set_fluid(sc["fluid_x"], sc.get("fluid_v"), ids=sc["fluid_ids"])
        n = ts.num_particles
        cap = ts.capacity
        x = ts.pinned((cap, 3))
        v = ts.pinned((cap, 3))
        rho = ts.pinned((cap,))
        x[:n] = sc["fluid_x"]
        v[:n] = sc["fluid_v"] if sc.get("fluid_v") is not None else 0
        e2e_steps = args.e2e_steps if args.e2e_steps > 0 else args.steps
        e2e_warm = args.warmup

        def e2e_step():
            st = ts.step_host(x, v, rho)
            e2e_iters.append((st.iterations_v, st.iterations))
    PROGRESS["phase"] = "warm-up (host buffers)"
    for k in range(e2e_warm):
        PROGRESS["step"] = k
        e2e_step()
    barrier()
    del e2e_iters[:]
    t0 = time.perf_counter()
    ts.timer_start()
    PROGRESS["phase"] = "timed steps (host buffers)"
    for k in range(e2e_steps):
        PROGRESS["step"] = k
        e2e_step()
    ms_e2e = ts.timer_stop()
    wall_e2e = (time.perf_counter() - t0) * 1000.0
    ms_e2e = max(ms_e2e, wall_e2e)   # host-side copies are synchronous: take the larger of device and wall time
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    h2d = 2 * 3 * R * n_global            # x, v of every particle of the job, per step (all ranks together)
    d2h = 2 * 3 * R * n_global + R * n_global   # x, v, density

    if rank == 0:
        peak, peak_src = load_peaks()
        total_ms = sum(p[0] for p in prof.values())
        dom = max(prof.items(), key=lambda kv: kv[1][0])[0]
        dms, dcnt = prof[dom]
        rb, cb = ALGO_BYTES[dom]
        bytes_per_launch = (rb * R + cb) * n
        achieved = bytes_per_launch / (dms / dcnt * 1e-3) / 1e9 if dcnt else 0.0
        nv = float(np.mean([i[0] for i in iters]))
        npr = float(np.mean([i[1] for i in iters]))
        step_bytes = ((101 + 17 * (nv + npr)) * R + 32) * n
        traffic = NCU_TRAFFIC.get((dom, args.precision, args.particles))
        line = {
            "metric": METRIC, "value": n_global * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": dict(cfg, parallelism=("single GPU" if world == 1 else
                                             f"{world} {'xyz'[sc.get('slab_axis', 2)]}-slabs, one per GPU: migration + ghost x,v per step over NCCL, ghost "
                                             "kappa / a per iteration and the error all-reduce over NVLink peer memory")),
            "e2e": {"value": n_global * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "warmup": e2e_warm, "ms_per_step": ms_e2e / e2e_steps,
                    "mean_iterations": {"divergence": float(np.mean([i[0] for i in e2e_iters])),
                                        "pressure": float(np.mean([i[1] for i in e2e_iters]))}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         # actual DRAM bytes (ncu capture, mostly the neighbour-index table) over the live launch time
                         "dram_frac": (traffic / (dms / dcnt * 1e-3) / 1e9 / peak) if (traffic and dcnt) else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": dms / dcnt if dcnt else None, "share_of_step": dms / total_ms if total_ms else None},
            "roofline_step": {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                              "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak, "unit": "GB/s"},
            "kernel_ms_per_step": {k: p[0] / args.steps for k, p in prof.items()},
            "kernel_launches": {k: p[1] for k, p in prof.items()},
            "mean_iterations": {"divergence": nv, "pressure": npr},
            "ms_search_per_step": ms_search / args.steps, "ms_solver_per_step": ms_solver / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = run_reference_cpu(args.cpu_sample, args.precision, min(args.steps, 10), args.warmup, budget_s=15.0)
                line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the CPU leg must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    PROGRESS["phase"] = "shutdown"
    ts.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

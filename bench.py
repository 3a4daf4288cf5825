#!/usr/bin/env python
"""Benchmark of the DFSPH hot path (per-step neighbourhood search + DFSPH pressure-solver loop).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--particles 10M] [--precision f32] [--impl reference]

Metric (BASELINE.json): DFSPH particle-updates/s = fluid particles x steps / device seconds, on a synthetic dam-break
block.  One "step" = one TimeStepDFSPH::step() (search + divergence solve + pressure solve + advection).
`value`   : state resident in HBM, W warm-up + K timed steps from the resting lattice, product path (solver loops as CUDA
            graphs), timed with CUDA events on the library's stream (max over ranks).
`e2e`     : the same window through dfsph_b200_step_host with pinned HOST buffers (H2D of x,v and D2H of x,v,density
            inside the timed region); `e2e_plugin`: the same window through TimeStepDFSPH_B200::step() inside the
            reference's own Simulation objects (oracle/_ref as the host of the plugin).
`roofline`: dominant kernel class of the timed window, algorithmic bytes / its CUDA-event launch duration /
            measured HBM peak, from a profiled replay of the same window; `roofline_kernels`: every class.
`steady_window`: steps 60..80 of the same run (18-30 pressure iterations per step instead of 2-3), same quantities.
`cpu_baseline`: the reference's own DFSPH sources (oracle/_ref, built by oracle/Makefile; neighbour search = our
            CompactNSearch-compatible stand-in) on the box's host cores, on a bounded sample of the same workload
            (the same block, 1 warm-up + up to 3 timed steps).
`--impl reference` runs the SAME block and window on the CPU (all host threads) and prints it as its own JSON line.
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
    # profiles/r2_ncu_10M_f32_summary.csv (ncu --set full, tools/prof_run.py f32 10M 1; round-2 kernels)
    ("jacobi_press", "f32", "10M"): 2.0598e9,
    ("jacobi_div", "f32", "10M"): 2.0615e9,
    ("accel", "f32", "10M"): 1.6434e9,
    ("build_neighbors", "f32", "10M"): 1.6848e9,     # k_build_tiles (the table it writes is 1.42 GB of it)
    ("init_sweep", "f32", "10M"): 2.4320e9,
    ("div_final", "f32", "10M"): 2.0736e9,
    ("press_init", "f32", "10M"): 2.0929e9,
    ("press_final", "f32", "10M"): 2.1486e9,
    # profiles/r2_ncu_10M_f64_summary.csv (the sweeps of the double build)
    ("jacobi_press", "f64", "10M"): 2.7672e9,
    ("jacobi_div", "f64", "10M"): 2.7672e9,
    ("accel", "f64", "10M"): 1.9348e9,
    ("init_sweep", "f64", "10M"): 3.3288e9,
    ("press_init", "f64", "10M"): 2.8783e9,
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
    """Time the reference's own CPU DFSPH (oracle/_ref; falls back to the C++ restatement) on the `sample` block: `warmup`
    untimed steps from rest, then up to `steps` timed steps (at least 2; stops early once `budget_s` seconds of timed
    stepping have passed).  The step cost depends on the phase of the scene, so the window is the B200 arm's window."""
    # all the host threads this process may use -- torchrun exports OMP_NUM_THREADS=1 to its workers, which would turn
    # the reference arm at N>1 into a single-thread run
    host_threads = int(os.environ.get("BENCH_CPU_THREADS", "0")) or len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(host_threads)
    from oracle import refsim, portsim
    from splishsplash_b200 import scenes
    dt = np.float32 if precision == "f32" else np.float64
    t_setup = time.time()
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
    t_setup = time.time() - t_setup
    if warmup > 0:
        sim.step(warmup)
    sim.reset_step_seconds()
    for k in ("neighborhood_search", "precomputeValues", "computeDFSPHFactor", "divergenceSolve", "pressureSolve"):
        sim.timer_ms(k)   # (averages are cumulative over the run; see ref_timers_ms below)
    done = 0
    t0 = time.time()
    iters = []
    while done < steps:
        sim.step(1)
        done += 1
        iters.append((sim.iterations_v, sim.iterations))
        if time.time() - t0 > budget_s and done >= 2:
            break
    secs = sim.step_seconds
    timers = {k: sim.timer_ms(k) for k in ("neighborhood_search", "precomputeValues", "computeDFSPHFactor", "divergenceSolve", "pressureSolve")}
    sim.destroy()
    value = n * done / secs
    return {"value": value, "unit": UNIT, "cores": int(cores), "kind": kind, "particles": int(n),
            "sample": f"dam-break {sample} block ({n} particles), {done} timed steps after {warmup} warm-up steps from rest, "
                      f"{'float+AVX' if precision == 'f32' else 'double scalar'} build, OMP threads={cores}; neighbour search = "
                      "CompactNSearch-compatible stand-in (oracle/standin, OpenMP), not CompactNSearch a9ab7c71",
            "ms_per_step": 1000.0 * secs / done, "steps": done, "setup_s": t_setup,
            "mean_iterations": [float(np.mean([i[0] for i in iters])), float(np.mean([i[1] for i in iters]))],
            # the reference's own timers (Utilities/Timing.h), average ms per call over the whole run incl. warm-up
            "ref_timers_ms": timers,
            "solver_only_ms_per_step": timers["divergenceSolve"] + timers["pressureSolve"]}


def run_plugin_leg(particles, precision, steps, warmup):
    """The drop-in as the reference sees it: oracle/_ref's Simulation / FluidModel / BoundaryModel_Akinci2012 / TimeManager
    objects with the product's C++ class TimeStepDFSPH_B200 installed as the time step; every step is one
    TimeStepDFSPH_B200::step() (upload of the FluidModel's x, v, device step, download of x, v, density into the
    FluidModel's arrays).  The reference stack is only the HOST of the plugin here; what is timed is the plugin call."""
    from oracle import refsim
    from splishsplash_b200 import scenes
    if not refsim.ref_available(precision):
        return None
    dt = np.float32 if precision == "f32" else np.float64
    sc = scenes.dam_break(particles, dtype=dt)
    sim = refsim.build_ref_scene(sc, precision, kernel=4, b200=True, **solver_params())
    try:
        n = sim.num_particles()
        out = {}
        for mode, key in ((1, "host_state_sync"), (0, "device_resident")):
            if mode == 0:
                sim.lib.ref_b200_set_host_sync(0)
            if warmup > 0 and mode == 1:
                sim.step(warmup)
            sim.reset_step_seconds()
            its = []
            for _ in range(steps):
                sim.step(1)
                its.append(sim.iterations)
            secs = sim.step_seconds
            out[key] = {"value": n * steps / secs, "ms_per_step": 1000.0 * secs / steps, "mean_pressure_iterations": float(np.mean(its))}
        sim.lib.ref_b200_download_state()
        return {"unit": UNIT, "steps": steps, "warmup": warmup, "particles": int(n),
                "what": "TimeStepDFSPH_B200::step() inside the reference's Simulation (oracle/_ref objects as host); wall clock "
                        "around step(). host_state_sync = the default (FluidModel x, v, density exchanged every step); "
                        "device_resident = setHostStateSync(false), the following steps of the same run", **out}
    finally:
        sim.destroy()


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


def kernel_table(prof, n, R, peak, precision, particles):
    """per kernel class: launches, ms per launch, algorithmic bytes per launch, achieved GB/s, fraction of the measured
    HBM peak, share of the profiled window, and (where an ncu capture is committed) the real DRAM traffic"""
    total = sum(p[0] for p in prof.values()) or 1.0
    out = {}
    for k, (ms, cnt) in prof.items():
        if not cnt:
            continue
        rb, cb = ALGO_BYTES[k]
        bpl = (rb * R + cb) * n
        per = ms / cnt
        tr = NCU_TRAFFIC.get((k, precision, particles))
        out[k] = {"launches": int(cnt), "ms_per_launch": per, "algorithmic_bytes_per_launch": bpl,
                  "achieved_gbs": bpl / (per * 1e-3) / 1e9, "frac": bpl / (per * 1e-3) / 1e9 / peak, "share": ms / total,
                  "traffic": tr, "dram_frac": (tr / (per * 1e-3) / 1e9 / peak) if tr else None}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--particles", default="10M")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", default="", help="block of the CPU legs (default: the same block as --particles)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-plugin-leg", action="store_true")
    ap.add_argument("--no-steady", action="store_true", help="skip the steady-phase window (steps 60..80 of the collapse)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps (the e2e leg replays the window of the device-resident run)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = one block of --particles per GPU (default); strong = one block of --particles cut into N slabs")
    ap.add_argument("--slab-axis", type=int, default=2, choices=[0, 2], help="N>1 weak scaling: 2 = slabs across the flow (default), 0 = along it")
    args = ap.parse_args()
    cpu_sample = args.cpu_sample or args.particles

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from splishsplash_b200 import scenes
    counts = scenes.NAMED_BLOCKS[args.particles]
    R = 4 if args.precision == "f32" else 8
    cfg = {"workload": workload_name(args.particles, args.precision, counts), "particles_per_gpu": int(np.prod(counts)),
           "precision": args.precision, "window": f"{args.steps} timed steps after {args.warmup} warm-up steps from the resting lattice",
           "l2": "particle state per GPU (>= 0.8 GB at 10M) exceeds the 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        # the SAME workload and window as the B200 arm: the named block, W warm-up steps from rest, then K timed steps
        # (the loop stops early only if the timed steps alone exceed the budget; `steps` reports what was run)
        r = run_reference_cpu(cpu_sample, args.precision, args.steps, args.warmup, budget_s=float(os.environ.get("BENCH_REF_BUDGET_S", "150")))
        if cpu_sample != args.particles:
            cfg["workload"] = workload_name(cpu_sample, args.precision, scenes.NAMED_BLOCKS[cpu_sample])
            cfg["particles_per_gpu"] = r["particles"]
        cfg["particles_total"] = r["particles"]
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": cfg,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "mean_iterations": r["mean_iterations"], "ref_timers_ms": r["ref_timers_ms"],
                "solver_only_ms_per_step": r["solver_only_ms_per_step"], "setup_s": r["setup_s"]}
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
    from splishsplash_b200 import capi
    # one rank = one process = one GPU: run on the CPUs of the GPU's NUMA node and take the pinned host buffers of the
    # end-to-end leg from that node's memory (8 ranks staging through one socket do not scale)
    numa_node = capi.load(args.precision).dfsph_b200_bind_host_numa(local_rank)

    dt = np.float32 if args.precision == "f32" else np.float64
    if world == 1:
        sc = scenes.dam_break(args.particles, dtype=dt)
        ts = build_b200_scene(sc, args.precision, device=local_rank, **solver_params())
    else:
        # weak scaling: `world` blocks of the named size side by side in one tank (strong: one block cut into `world`
        # slabs), one slab per GPU; migration and ghost exchange happen inside the library (csrc/multi_gpu.cuh)
        from splishsplash_b200 import parallel
        if args.scaling == "weak":
            sc = scenes.dam_break_weak(rank, world, args.particles, dtype=dt, axis=args.slab_axis)
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

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def resubmit():
        """the initial state again, on the same context (and communicator)"""
        barrier()
        ts.setValue("timeStepSize", solver_params()["timeStepSize"])
        if world == 1:
            ts.set_fluid(sc["fluid_x"], sc.get("fluid_v"))
        else:
            ts.set_fluid(sc["fluid_x"], sc.get("fluid_v"), ids=sc["fluid_ids"])

    def run_window(warm, steps, phase, profile=False):
        """`warm` untimed + `steps` timed device-resident steps; returns (ms, iteration list, launches, search ms, solver ms, profile)"""
        PROGRESS["phase"] = f"{phase}: warm-up"
        PROGRESS["step"] = -1
        if warm > 0:
            ts.step(warm, sync=False)   # no statistics, no host synchronisation between the steps
        barrier()
        if profile:
            ts.set_profiling(True)
        ts.timer_start()
        launches, iters, ms_search, ms_solver = 0, [], 0.0, 0.0
        PROGRESS["phase"] = f"{phase}: timed steps"
        for k in range(steps):
            PROGRESS["step"] = k
            st = ts.step(1)
            launches += st.gpu_launches
            iters.append((st.iterations_v, st.iterations))
            ms_search += st.ms_search
            ms_solver += st.ms_solver
        ms = ts.timer_stop()
        prof = None
        if profile:
            prof = ts.profile()
            ts.set_profiling(False)
        barrier()
        return max_over_ranks(ms), iters, launches, ms_search, ms_solver, prof

    # ---- window 1 (the headline): W warm-up + K timed steps from rest, product path (solver loops as CUDA graphs, no
    # per-kernel events), state resident in HBM
    sampler = ClockSampler(local_rank) if rank == 0 else None   # one nvidia-smi poller per job, not per rank
    if sampler:
        sampler.start()
    ms, iters, launches, ms_search, ms_solver, _ = run_window(args.warmup, args.steps, "device-resident window")
    clocks = sampler.stop() if sampler else None

    # ---- the same window again with per-kernel-class CUDA events (host-driven loops): kernel table and roofline
    resubmit()
    ms_prof, iters_prof, _, _, _, prof = run_window(args.warmup, args.steps, "profiled replay", profile=True)

    # ---- steady phase of the collapse: steps 60..80 (the solver needs an order of magnitude more iterations there)
    steady = None
    if not args.no_steady:
        done = args.warmup + args.steps
        first = max(60, done)
        s_ms, s_iters, _, s_search, s_solver, _ = run_window(first - done, 20, "steady window")
        _, sp_iters, _, _, _, s_prof = run_window(0, 10, "steady window, profiled", profile=True)
        steady = (first, s_ms, s_iters, s_search, s_solver, s_prof, sp_iters)

    # ---- end-to-end through host buffers (pinned): H2D of x, v and D2H of x, v, density inside the timed region.
    # Same workload, same window: the initial state is re-submitted and warm-up + timed steps are replayed through step_host
    # (multi-GPU: ids are global, so the host buffers are in this rank's device order, capacity rows).
    e2e_iters = []
    PROGRESS["phase"] = "re-submitting the initial state"
    resubmit()
    n = ts.num_particles
    rows = n if world == 1 else ts.capacity
    x = ts.pinned((rows, 3))
    v = ts.pinned((rows, 3))
    rho = ts.pinned((rows,))
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
    ms_e2e = max_over_ranks(max(ms_e2e, wall_e2e))   # host-side copies are synchronous: take the larger of device and wall time
    numa_nodes = [numa_node]
    if world > 1:
        numa_nodes = [None] * world
        dist.all_gather_object(numa_nodes, numa_node)
    h2d = 2 * 3 * R * n_global            # x, v of every particle of the job, per step (all ranks together)
    d2h = 2 * 3 * R * n_global + R * n_global   # x, v, density

    if rank == 0:
        peak, peak_src = load_peaks()
        n_loc = int(n_global // world)
        table = kernel_table(prof, n_loc, R, peak, args.precision, args.particles)
        dom = max(table.items(), key=lambda kv: kv[1]["share"])[0]
        d = table[dom]
        nv = float(np.mean([i[0] for i in iters]))
        npr = float(np.mean([i[1] for i in iters]))
        step_bytes = ((101 + 17 * (nv + npr)) * R + 32) * n_loc
        line = {
            "metric": METRIC, "value": n_global * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": dict(cfg, parallelism=("single GPU" if world == 1 else
                                             f"{world} {'xyz'[sc.get('slab_axis', 2)]}-slabs, one per GPU: migration + ghost x,v per step over NCCL, ghost "
                                             "kappa / a per iteration and the error all-reduce over NVLink peer memory")),
            "e2e": {"value": n_global * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "warmup": e2e_warm, "ms_per_step": ms_e2e / e2e_steps,
                    "api": "dfsph_b200_step_host (C ABI, pinned host buffers)",
                    "host_numa_node_per_rank": numa_nodes,   # dfsph_b200_bind_host_numa: CPUs + pinned memory next to each rank's GPU (-1: unknown)
                    "mean_iterations": {"divergence": float(np.mean([i[0] for i in e2e_iters])),
                                        "pressure": float(np.mean([i[1] for i in e2e_iters]))}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # dominant kernel class of the timed window (largest share of the kernel time), from the profiled replay of
            # the same window: algorithmic bytes / CUDA-event launch duration / measured HBM peak
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                         "frac": d["frac"], "traffic": d["traffic"], "dram_frac": d["dram_frac"],
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
                         "avg_launch_ms": d["ms_per_launch"], "share_of_step": d["share"],
                         "measured_in": f"profiled replay of the timed window ({ms_prof / args.steps:.3f} ms/step with per-kernel events and "
                                        f"host-driven loops, {float(np.mean([i[1] for i in iters_prof])):.2f} pressure iterations)"},
            "roofline_kernels": table,
            "roofline_step": {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                              "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak, "unit": "GB/s"},
            "mean_iterations": {"divergence": nv, "pressure": npr},
            "ms_search_per_step": ms_search / args.steps, "ms_solver_per_step": ms_solver / args.steps,
        }
        if steady:
            first, s_ms, s_iters, s_search, s_solver, s_prof, sp_iters = steady
            s_nv = float(np.mean([i[0] for i in s_iters]))
            s_np = float(np.mean([i[1] for i in s_iters]))
            s_bytes = ((101 + 17 * (s_nv + s_np)) * R + 32) * n_loc
            s_table = kernel_table(s_prof, n_loc, R, peak, args.precision, args.particles)
            s_dom = max(s_table.items(), key=lambda kv: kv[1]["share"])[0]
            line["steady_window"] = {
                "window": f"steps {first}..{first + 20} of the same run (device-resident, product path)",
                "value": n_global * 20 / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms / 20,
                "mean_iterations": {"divergence": s_nv, "pressure": s_np},
                "ms_search_per_step": s_search / 20, "ms_solver_per_step": s_solver / 20,
                "roofline_step": {"algorithmic_bytes_per_step": s_bytes, "achieved": s_bytes / (s_ms / 20 * 1e-3) / 1e9,
                                  "frac": s_bytes / (s_ms / 20 * 1e-3) / 1e9 / peak, "unit": "GB/s"},
                "dominant_kernel": s_dom, "roofline_kernels": s_table,
                "profiled": f"steps {first + 20}..{first + 30}, {float(np.mean([i[1] for i in sp_iters])):.2f} pressure iterations",
            }
    PROGRESS["phase"] = "shutdown"
    ts.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        if world == 1 and not args.no_plugin_leg:
            try:
                pl = run_plugin_leg(args.particles, args.precision, args.steps, args.warmup)
                if pl:
                    line["e2e_plugin"] = pl
            except Exception as e:
                line["e2e_plugin"] = {"unavailable": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                # bounded sample of the same workload: the same block, the first steps of the window only
                r = run_reference_cpu(cpu_sample, args.precision, 3, min(args.warmup, 1), budget_s=20.0)
                line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
                line["cpu_baseline"]["solver_only_ms_per_step"] = r["solver_only_ms_per_step"]
                line["cpu_baseline"]["ms_per_step"] = r["ms_per_step"]
                line["cpu_baseline"]["mean_iterations"] = r["mean_iterations"]
            except Exception as e:  # the CPU leg must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
        os.write(json_fd, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()

"""Python host-side mirror of the reference's DFSPH solver interface over the C ABI.

``TimeStepDFSPH_B200`` exposes what ``SPH::TimeStepDFSPH`` exposes to its callers -- ``step()``, ``reset()``,
``getMethodName()``, ``getNumIterations()``, the GenParam parameter names of SPlisHSPlasH/DFSPH/TimeStepDFSPH.cpp:73-115
(``minIterations``, ``maxIterations``, ``maxError``, ``maxIterationsV``, ``maxErrorV``, ``enableDivergenceSolver``,
read-only ``iterations`` / ``iterationsV``) and the particle fields of :49-53 -- plus the Simulation / TimeManager
parameters the hot path reads (SPlisHSPlasH/Simulation.cpp:163-277).  The C++ drop-in with the same name lives in
splishsplash_b200/host/TimeStepDFSPH_B200.{h,cpp}; this module is what tests/ and bench.py drive.

Everything here is plumbing: all arithmetic happens in the CUDA library (splishsplash_b200/csrc).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

_TS_PARAMS = {  # reference parameter name -> Params field
    "minIterations": "min_iterations", "maxIterations": "max_iterations", "maxError": "max_error",
    "maxIterationsV": "max_iterations_v", "maxErrorV": "max_error_v",
    "enableDivergenceSolver": "enable_divergence_solver",
    # Simulation / TimeManager
    "timeStepSize": "time_step_size", "cflMethod": "cfl_method", "cflFactor": "cfl_factor",
    "cflMinTimeStepSize": "cfl_min_time_step_size", "cflMaxTimeStepSize": "cfl_max_time_step_size",
    # FluidModel / Viscosity_Standard (next-row f1)
    "viscosityMethod": "viscosity_method", "viscosity": "viscosity", "viscosityBoundary": "viscosity_boundary",
}


class TimeStepDFSPH_B200:
    METHOD_NAME = "DFSPH_B200"

    def __init__(self, precision="f32", particle_radius=0.025, kernel=capi.KERNEL_PRECOMPUTED_CUBIC, device=0,
                 max_fluid_neighbors=64, max_boundary_neighbors=64, max_fluid_particles=0, domain=None, grad_kernel=None):
        self.precision = precision
        self.dtype = np.float32 if precision == "f32" else np.float64
        self.lib = capi.load(precision)
        cfg = capi.Config()
        self.lib.dfsph_b200_default_config(C.byref(cfg))
        cfg.device = device
        cfg.kernel = kernel
        cfg.grad_kernel = -1 if grad_kernel is None else grad_kernel   # Simulation "gradKernel"; default: same as "kernel"
        cfg.particle_radius = particle_radius
        cfg.max_fluid_neighbors = max_fluid_neighbors
        cfg.max_boundary_neighbors = max_boundary_neighbors
        cfg.max_fluid_particles = max_fluid_particles
        if domain is not None:   # explicit cell-grid extent (mandatory and identical on all ranks in multi-GPU runs)
            for k in range(3):
                cfg.domain_min[k] = float(domain[0][k])
                cfg.domain_max[k] = float(domain[1][k])
        self.cfg = cfg
        self.ctx = C.c_void_p()
        rc = self.lib.dfsph_b200_create(C.byref(cfg), C.byref(self.ctx))
        if rc != 0:
            msg = self.lib.dfsph_b200_last_error(None).decode()
            self.ctx = None
            raise capi.DFSPHError(rc, msg)
        self.params = capi.Params()
        self.lib.dfsph_b200_default_params(C.byref(self.params))
        self.stats = capi.StepStats()
        self.particle_radius = particle_radius
        self._pinned = []
        self._nb_added = 0

    # ---- error handling ----
    def _check(self, rc):
        if rc != 0:
            raise capi.DFSPHError(rc, self.lib.dfsph_b200_last_error(self.ctx).decode())

    def close(self):
        if getattr(self, "ctx", None):
            for p in self._pinned:
                self.lib.dfsph_b200_free_pinned(p)
            self._pinned = []
            self.lib.dfsph_b200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- reference-style interface ----
    def getMethodName(self):
        return self.METHOD_NAME

    def getNumIterations(self):
        return int(self.stats.iterations)

    @property
    def iterations(self):
        return int(self.stats.iterations)

    @property
    def iterationsV(self):
        return int(self.stats.iterations_v)

    @property
    def h(self):
        return float(self.stats.time_step_size) if self.stats.num_particles else float(self.params.time_step_size)

    @property
    def time(self):
        return float(self.stats.time)

    def setValue(self, name, value):
        """GenParam-style setter by reference parameter name."""
        if name == "gravitation":
            for k in range(3):
                self.params.gravitation[k] = float(value[k])
        elif name in _TS_PARAMS:
            f = _TS_PARAMS[name]
            cur = getattr(self.params, f)
            setattr(self.params, f, type(cur)(value))
        else:
            raise KeyError(name)
        self._check(self.lib.dfsph_b200_set_params(self.ctx, C.byref(self.params)))
        self._check(self.lib.dfsph_b200_get_params(self.ctx, C.byref(self.params)))

    def getValue(self, name):
        if name == "iterations":
            return self.iterations
        if name == "iterationsV":
            return self.iterationsV
        if name == "gravitation":
            return [self.params.gravitation[k] for k in range(3)]
        return getattr(self.params, _TS_PARAMS[name])

    def set(self, **kw):
        for k, v in kw.items():
            self.setValue(k, v)

    # ---- model set-up ----
    def set_fluid(self, x, v=None, ids=None, state=None, density0=1000.0, volume=None):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        n = x.shape[0]
        if volume is None:   # FluidModel::initMasses (FluidModel.cpp:234-243), arithmetic in Real
            dt = self.dtype
            diam = dt(2.0) * dt(self.particle_radius)
            volume = float(dt(0.8) * diam * diam * diam)
        vp = ip = sp = None
        if v is not None:
            v = np.ascontiguousarray(v, dtype=self.dtype)
            vp = v.ctypes.data
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            ip = ids.ctypes.data
        if state is not None:
            state = np.ascontiguousarray(state, dtype=np.uint32)
            sp = state.ctypes.data
        self._check(self.lib.dfsph_b200_set_params(self.ctx, C.byref(self.params)))
        self._check(self.lib.dfsph_b200_set_fluid(self.ctx, n, x.ctypes.data if n else None, vp, ip, sp,
                                                  float(density0), float(volume)))

    def add_boundary(self, x, V=None, is_dynamic=False):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        Vp = None
        if V is not None:
            V = np.ascontiguousarray(V, dtype=self.dtype)
            Vp = V.ctypes.data
        self._check(self.lib.dfsph_b200_add_boundary(self.ctx, x.shape[0], x.ctypes.data if x.shape[0] else None, Vp,
                                                     1 if is_dynamic else 0))
        self._nb_added += x.shape[0]

    def compute_boundary_volume(self):
        self._check(self.lib.dfsph_b200_compute_boundary_volume(self.ctx))

    # ---- stepping ----
    def step(self, n=1, sync=True):
        for k in range(n):
            last = (k == n - 1)
            self._check(self.lib.dfsph_b200_step(self.ctx, C.byref(self.stats) if (sync or last) else None))
        return self.stats

    def step_host(self, x, v, density=None):
        """One step through host buffers (x, v in id order; pinned buffers from ``pinned`` recommended).  Multi-GPU: rows in
        device order, buffers of ``capacity`` rows (see include/dfsph_b200.h)."""
        dp = density.ctypes.data if density is not None else None
        self._check(self.lib.dfsph_b200_step_host(self.ctx, x.ctypes.data, v.ctypes.data, dp, C.byref(self.stats)))
        return self.stats

    def pinned(self, shape, dtype=None):
        arr, p = capi.pinned_array(self.lib, shape, dtype or self.dtype)
        self._pinned.append(p)
        return arr

    def set_profiling(self, on=True):
        self._check(self.lib.dfsph_b200_set_profiling(self.ctx, 1 if on else 0))

    def profile(self):
        """{kernel class: (total ms, launches)} accumulated since set_profiling(True)."""
        n = len(capi.PROF_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_uint64 * n)()
        self._check(self.lib.dfsph_b200_get_profile(self.ctx, ms, cnt))
        return {capi.PROF_CLASSES[k]: (float(ms[k]), int(cnt[k])) for k in range(n)}

    def timer_start(self):
        self._check(self.lib.dfsph_b200_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self.lib.dfsph_b200_timer_stop(self.ctx, C.byref(ms)))
        return float(ms.value)

    def synchronize(self):
        self._check(self.lib.dfsph_b200_synchronize(self.ctx))

    def search_and_density(self):
        self._check(self.lib.dfsph_b200_search_and_density(self.ctx))

    # ---- state access ----
    @property
    def num_particles(self):
        return int(self.lib.dfsph_b200_num_particles(self.ctx))

    @property
    def capacity(self):
        """Fluid rows the device arrays hold (size of the step_host buffers in multi-GPU runs)."""
        return int(self.lib.dfsph_b200_capacity(self.ctx))

    @property
    def num_boundary_particles(self):
        return int(self.lib.dfsph_b200_num_boundary_particles(self.ctx))

    def field(self, name, by_id=True):
        """Particle field by the reference's FieldDescription name.  by_id: row k = particle with original index k."""
        fid, dim, is_uint = capi.FIELDS[name]
        n = self.num_particles
        dt = np.uint32 if is_uint else self.dtype
        out = np.empty((n, dim) if dim > 1 else (n,), dtype=dt)
        self._check(self.lib.dfsph_b200_download(self.ctx, fid, out.ctypes.data, out.nbytes, 1 if by_id else 0))
        return out

    def set_field(self, name, arr, by_id=True):
        fid, dim, is_uint = capi.FIELDS[name]
        arr = np.ascontiguousarray(arr, dtype=np.uint32 if is_uint else self.dtype)
        self._check(self.lib.dfsph_b200_upload(self.ctx, fid, arr.ctypes.data, arr.nbytes, 1 if by_id else 0))

    def boundary_volume(self):
        out = np.empty(self._nb_added, dtype=self.dtype)
        self._check(self.lib.dfsph_b200_download(self.ctx, capi.FIELD_BOUNDARY_VOLUME, out.ctypes.data, out.nbytes, 0))
        return out

    def neighbors(self, other=0, lists=True):
        """(counts, offsets, idx): neighbours of fluid particles in set ``other`` (0 fluid, 1 boundary), rows in current
        device order, fluid indices in current device order, boundary indices in insertion order."""
        n = self.num_particles
        counts = np.zeros(n, dtype=np.uint32)
        if not lists:
            self._check(self.lib.dfsph_b200_neighbors(self.ctx, other, counts.ctypes.data, None, None, 0))
            return counts, None, None
        self._check(self.lib.dfsph_b200_neighbors(self.ctx, other, counts.ctypes.data, None, None, 0))
        total = int(counts.sum(dtype=np.uint64))
        offsets = np.zeros(n + 1, dtype=np.uint64)
        idx = np.zeros(max(total, 1), dtype=np.uint32)
        self._check(self.lib.dfsph_b200_neighbors(self.ctx, other, counts.ctypes.data, offsets.ctypes.data,
                                                  idx.ctypes.data, total))
        return counts, offsets, idx[:total]

    def eval_kernel(self, r, kernel=-1):
        r = np.ascontiguousarray(r, dtype=self.dtype)
        n = r.shape[0]
        W = np.empty(n, dtype=self.dtype)
        g = np.empty((n, 3), dtype=self.dtype)
        self._check(self.lib.dfsph_b200_eval_kernel(self.ctx, kernel, n, r.ctypes.data, W.ctypes.data, g.ctypes.data))
        return W, g


def build_b200_scene(scene, precision="f32", kernel=capi.KERNEL_PRECOMPUTED_CUBIC, boundary_V=None, device=0, grad_kernel=None,
                     **params):
    """Create a TimeStepDFSPH_B200 for a ``splishsplash_b200.scenes`` scene dict (same call order as the reference
    harness oracle/refsim.build_ref_scene)."""
    cap = {k: params.pop(k) for k in ("max_fluid_neighbors", "max_boundary_neighbors") if k in params}   # table capacities (config)
    ts = TimeStepDFSPH_B200(precision, scene["radius"], kernel, device=device, grad_kernel=grad_kernel, **cap)
    if params:
        ts.set(**params)
    ts.set_fluid(scene["fluid_x"], scene.get("fluid_v"))
    bx = scene.get("boundary_x")
    if bx is not None and len(bx):
        ts.add_boundary(bx, boundary_V)
        if boundary_V is None:
            ts.compute_boundary_volume()
    return ts

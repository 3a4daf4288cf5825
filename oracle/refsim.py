"""TEST INFRASTRUCTURE ONLY.  ctypes wrapper around oracle/_ref/libsplish_ref_{f32,f64}.so, i.e. the reference's own
unmodified DFSPH sources (see oracle/Makefile, oracle/ref_driver.cpp).  The reference is a process-wide singleton
(Simulation::getCurrent()), so one RefSim may exist per process and per precision at a time.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def ref_lib_path(precision: str) -> str:
    return os.path.join(_HERE, "_ref", f"libsplish_ref_{precision}.so")


def ref_available(precision: str) -> bool:
    return os.path.exists(ref_lib_path(precision))


FIELDS = {  # reference FieldDescription name -> dim
    "position": 3, "velocity": 3, "acceleration": 3, "density": 1, "factor": 1, "advected density": 1,
    "p / rho^2": 1, "p_v / rho^2": 1, "pressure acceleration": 3,
}


class RefSim:
    """The reference's Simulation + TimeStepDFSPH driven through the C harness."""

    def __init__(self, precision: str = "f64", lib_path: str | None = None):
        assert precision in ("f32", "f64")
        self.precision = precision
        self.dtype = np.float32 if precision == "f32" else np.float64
        self.lib = C.CDLL(lib_path or ref_lib_path(precision))
        L = self.lib
        for name in ("ref_step_seconds", "ref_time", "ref_time_step_size", "ref_w_zero", "ref_avg_timer_ms"):
            getattr(L, name).restype = C.c_double
        L.ref_fluid_volume.restype = C.c_double
        L.ref_fluid_volume.argtypes = [C.c_int]
        L.ref_avg_timer_ms.argtypes = [C.c_char_p]
        L.ref_create.argtypes = [C.c_double]
        L.ref_add_fluid.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_double]
        L.ref_add_boundary.argtypes = [C.c_void_p, C.c_uint]
        L.ref_configure.argtypes = [C.c_int, C.c_int]
        L.ref_set_real.argtypes = [C.c_char_p, C.c_double]
        L.ref_set_int.argtypes = [C.c_char_p, C.c_int]
        L.ref_set_gravity.argtypes = [C.c_double] * 3
        L.ref_step.argtypes = [C.c_int]
        L.ref_num_particles.restype = C.c_uint
        L.ref_num_particles.argtypes = [C.c_int]
        L.ref_num_boundary_particles.restype = C.c_uint
        L.ref_num_boundary_particles.argtypes = [C.c_int]
        L.ref_get_field.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_int]
        L.ref_get_ids.argtypes = [C.c_int, C.c_void_p]
        L.ref_set_field.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_int]
        L.ref_set_state.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_get_boundary.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_neighbor_counts.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.ref_neighbor_lists.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_set_num_threads.argtypes = [C.c_int]
        L.ref_set_viscosity.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        if hasattr(L, "ref_configure_b200"):   # only the reference build carries the drop-in solver
            L.ref_configure_b200.argtypes = [C.c_int, C.c_int, C.c_char_p]
            L.ref_last_error.restype = C.c_char_p
            L.ref_method_name.restype = C.c_char_p
        assert L.ref_sizeof_real() == np.dtype(self.dtype).itemsize
        self._alive = False
        self.n_fluid_models = 0
        self.n_boundary_models = 0

    # ---- scene construction (order mirrors SimulatorBase::initSimulation / deferredInit) ----
    def create(self, radius: float):
        rc = self.lib.ref_create(float(radius))
        if rc != 0:
            raise RuntimeError("a reference Simulation already exists in this process")
        self._alive = True

    def add_fluid(self, x, v=None, density0=1000.0):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        vp = None
        if v is not None:
            v = np.ascontiguousarray(v, dtype=self.dtype)
            vp = v.ctypes.data
        self.n_fluid_models += 1
        return self.lib.ref_add_fluid(x.ctypes.data, vp, x.shape[0], float(density0))

    def configure(self, kernel=4, grad_kernel=None):
        self.lib.ref_configure(int(kernel), int(kernel if grad_kernel is None else grad_kernel))

    def configure_b200(self, kernel=4, libdir=None, grad_kernel=None):
        """Install the product's C++ drop-in TimeStepDFSPH_B200 as the solver of the reference Simulation."""
        libdir = libdir or os.path.join(os.path.dirname(_HERE), "splishsplash_b200")
        rc = self.lib.ref_configure_b200(int(kernel), int(kernel if grad_kernel is None else grad_kernel), libdir.encode())
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def eval_kernel(self, r, kind):
        """W[n], gradW[n,3] of the kernel with Simulation id `kind` (0 cubic .. 4 precomputed cubic) at the points r."""
        r = np.ascontiguousarray(r, dtype=self.dtype)
        W = np.empty(len(r), dtype=self.dtype)
        g = np.empty((len(r), 3), dtype=self.dtype)
        self.lib.ref_eval_kernel.restype = C.c_int
        rc = self.lib.ref_eval_kernel(int(kind), C.c_uint(len(r)), C.c_void_p(r.ctypes.data), C.c_void_p(W.ctypes.data), C.c_void_p(g.ctypes.data))
        if rc != 0:
            raise ValueError(f"unknown kernel id {kind}")
        return W, g

    def configure_by_method_id(self, method=7, kernel=4, grad_kernel=None, libdir=None):
        """Select the solver through the reference's own "simulationMethod" enum (needs the library built with
        patches/register_dfsph_b200.patch: oracle/_ref/libsplish_ref_patched_f64.so)."""
        os.environ["DFSPH_B200_LIB_DIR"] = libdir or os.path.join(os.path.dirname(_HERE), "splishsplash_b200")
        self.lib.ref_configure_by_method_id.restype = C.c_int
        rc = self.lib.ref_configure_by_method_id(int(method), int(kernel), int(kernel if grad_kernel is None else grad_kernel))
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())

    @property
    def method_name(self):
        return self.lib.ref_method_name().decode()

    def add_boundary(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        self.n_boundary_models += 1
        return self.lib.ref_add_boundary(x.ctypes.data, x.shape[0])

    def finalize(self):
        self.lib.ref_finalize()

    def set_viscosity(self, method=1, viscosity=0.01, viscosity_boundary=0.0, fluid=0):
        if self.lib.ref_set_viscosity(fluid, int(method), float(viscosity), float(viscosity_boundary)) != 0:
            raise ValueError("unsupported viscosity method")

    def set(self, **kw):
        visc = {k: kw.pop(k) for k in ("viscosityMethod", "viscosity", "viscosityBoundary") if k in kw}
        if visc:
            self.set_viscosity(visc.get("viscosityMethod", 1), visc.get("viscosity", 0.01), visc.get("viscosityBoundary", 0.0))
        for k, v in kw.items():
            if k == "gravitation":
                self.lib.ref_set_gravity(*[float(c) for c in v])
            elif isinstance(v, (bool, int, np.integer)) and k not in ("maxError", "maxErrorV", "timeStepSize", "cflFactor",
                                                                        "cflMinTimeStepSize", "cflMaxTimeStepSize"):
                if self.lib.ref_set_int(k.encode(), int(v)) != 0:
                    raise KeyError(k)
            else:
                if self.lib.ref_set_real(k.encode(), float(v)) != 0:
                    raise KeyError(k)

    # ---- stepping ----
    def step(self, n=1):
        self.lib.ref_step(int(n))

    def search_and_density(self):
        self.lib.ref_search_and_density()

    @property
    def iterations(self):
        return self.lib.ref_iterations()

    @property
    def iterations_v(self):
        return self.lib.ref_iterations_v()

    @property
    def h(self):
        return self.lib.ref_time_step_size()

    @property
    def time(self):
        return self.lib.ref_time()

    @property
    def step_seconds(self):
        return self.lib.ref_step_seconds()

    def reset_step_seconds(self):
        self.lib.ref_reset_step_seconds()

    def timer_ms(self, name):
        return self.lib.ref_avg_timer_ms(name.encode())

    def num_particles(self, fluid=0):
        return self.lib.ref_num_particles(fluid)

    # ---- state access ----
    def field(self, name, fluid=0):
        dim = FIELDS[name]
        n = self.num_particles(fluid)
        out = np.empty((n, dim) if dim > 1 else (n,), dtype=self.dtype)
        self.lib.ref_get_field(fluid, name.encode(), out.ctypes.data, dim)
        return out

    def ids(self, fluid=0):
        out = np.empty(self.num_particles(fluid), dtype=np.uint32)
        self.lib.ref_get_ids(fluid, out.ctypes.data)
        return out

    def field_by_id(self, name, fluid=0):
        """Field re-ordered so that row k belongs to the particle whose original id is k."""
        f = self.field(name, fluid)
        out = np.empty_like(f)
        out[self.ids(fluid)] = f
        return out

    def set_field_by_id(self, name, arr, fluid=0):
        """Overwrite position / velocity / p / rho^2 / p_v / rho^2; row k of ``arr`` belongs to the particle with id k."""
        arr = np.ascontiguousarray(np.asarray(arr, dtype=self.dtype)[self.ids(fluid)])
        if self.lib.ref_set_field(fluid, name.encode(), arr.ctypes.data, FIELDS[name]) != 0:
            raise KeyError(name)

    def set_state(self, x=None, v=None, fluid=0):
        xp = vp = None
        if x is not None:
            x = np.ascontiguousarray(x, dtype=self.dtype)
            xp = x.ctypes.data
        if v is not None:
            v = np.ascontiguousarray(v, dtype=self.dtype)
            vp = v.ctypes.data
        self.lib.ref_set_state(fluid, xp, vp)

    def boundary(self, b=0):
        n = self.lib.ref_num_boundary_particles(b)
        x = np.empty((n, 3), dtype=self.dtype)
        V = np.empty(n, dtype=self.dtype)
        self.lib.ref_get_boundary(b, x.ctypes.data, V.ctypes.data)
        return x, V

    def neighbors(self, fluid=0, pid=0):
        """(counts, offsets, idx) CSR in current array order; idx holds array indices into point set ``pid``."""
        n = self.num_particles(fluid)
        counts = np.empty(n, dtype=np.uint32)
        self.lib.ref_neighbor_counts(fluid, pid, counts.ctypes.data)
        offsets = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(counts, out=offsets[1:])
        idx = np.empty(int(offsets[-1]), dtype=np.uint32)
        self.lib.ref_neighbor_lists(fluid, pid, offsets.ctypes.data, idx.ctypes.data)
        return counts, offsets, idx

    def destroy(self):
        if self._alive:
            self.lib.ref_destroy()
            self._alive = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.destroy()


def build_ref_scene(scene, precision="f64", kernel=4, lib_path=None, b200=False, grad_kernel=None, **params):
    """Create a RefSim for a ``splishsplash_b200.scenes`` scene dict.  b200=True swaps the reference's TimeStepDFSPH
    for the drop-in TimeStepDFSPH_B200 (everything else stays the reference's own code)."""
    sim = RefSim(precision, lib_path)
    sim.create(scene["radius"])
    sim.add_fluid(scene["fluid_x"], scene.get("fluid_v"))
    if b200 == "enum":
        sim.configure_by_method_id(7, kernel, grad_kernel)
    elif b200:
        sim.configure_b200(kernel, grad_kernel=grad_kernel)
    else:
        sim.configure(kernel, grad_kernel)
    if scene.get("boundary_x") is not None and len(scene["boundary_x"]):
        sim.add_boundary(scene["boundary_x"])
    if params:
        sim.set(**params)
    sim.finalize()
    return sim

"""ctypes binding of the C ABI in include/dfsph_b200.h (libdfsph_b200_{f32,f64}.so, built in-tree by
``__graft_entry__.build()``).  This is the only way Python reaches the CUDA kernels; there is no CPU fallback: if the
shared object is missing, or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))

OK = 0
ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED, ERR_COMM = -1, -2, -3, -4, -5
KERNEL_CUBIC, KERNEL_WENDLAND_QUINTIC_C2, KERNEL_POLY6, KERNEL_SPIKY, KERNEL_PRECOMPUTED_CUBIC = 0, 1, 2, 3, 4

# dfsph_b200_field
FIELD_POSITION, FIELD_VELOCITY, FIELD_DENSITY, FIELD_FACTOR, FIELD_DENSITY_ADV = 0, 1, 2, 3, 4
FIELD_KAPPA, FIELD_KAPPA_V, FIELD_PRESSURE_ACCEL, FIELD_ID, FIELD_STATE = 5, 6, 7, 8, 9
FIELD_BOUNDARY_VOLUME, FIELD_NUM_NEIGHBORS = 10, 11

# reference FieldDescription names -> (field id, components, is_uint)
FIELDS = {
    "position": (FIELD_POSITION, 3, False), "velocity": (FIELD_VELOCITY, 3, False),
    "density": (FIELD_DENSITY, 1, False), "factor": (FIELD_FACTOR, 1, False),
    "advected density": (FIELD_DENSITY_ADV, 1, False), "p / rho^2": (FIELD_KAPPA, 1, False),
    "p_v / rho^2": (FIELD_KAPPA_V, 1, False), "pressure acceleration": (FIELD_PRESSURE_ACCEL, 3, False),
    "id": (FIELD_ID, 1, True), "state": (FIELD_STATE, 1, True), "num_neighbors": (FIELD_NUM_NEIGHBORS, 1, True),
}


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("kernel", C.c_int32), ("particle_radius", C.c_double),
                ("max_fluid_particles", C.c_uint64), ("max_fluid_neighbors", C.c_int32),
                ("max_boundary_neighbors", C.c_int32), ("domain_min", C.c_double * 3), ("domain_max", C.c_double * 3),
                ("rank", C.c_int32), ("world_size", C.c_int32), ("grad_kernel", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("time_step_size", C.c_double), ("gravitation", C.c_double * 3),
                ("min_iterations", C.c_uint32), ("max_iterations", C.c_uint32), ("max_error", C.c_double),
                ("max_iterations_v", C.c_uint32), ("max_error_v", C.c_double),
                ("enable_divergence_solver", C.c_int32), ("cfl_method", C.c_int32), ("cfl_factor", C.c_double),
                ("cfl_min_time_step_size", C.c_double), ("cfl_max_time_step_size", C.c_double),
                ("viscosity_method", C.c_int32), ("viscosity", C.c_double), ("viscosity_boundary", C.c_double)]


class StepStats(C.Structure):
    _fields_ = [("iterations", C.c_uint32), ("iterations_v", C.c_uint32), ("avg_density_error", C.c_double),
                ("avg_density_error_v", C.c_double), ("time_step_size", C.c_double), ("time", C.c_double),
                ("num_particles", C.c_uint32), ("max_neighbors", C.c_uint32), ("gpu_launches", C.c_uint32),
                ("ms_search", C.c_float), ("ms_solver", C.c_float)]


# every symbol include/dfsph_b200.h declares
EXPORTS = [
    "dfsph_b200_sizeof_real", "dfsph_b200_version", "dfsph_b200_default_config", "dfsph_b200_default_params",
    "dfsph_b200_create", "dfsph_b200_destroy", "dfsph_b200_last_error", "dfsph_b200_set_fluid",
    "dfsph_b200_add_boundary", "dfsph_b200_compute_boundary_volume", "dfsph_b200_set_params", "dfsph_b200_get_params",
    "dfsph_b200_step", "dfsph_b200_step_host", "dfsph_b200_download", "dfsph_b200_upload", "dfsph_b200_neighbors",
    "dfsph_b200_search_and_density", "dfsph_b200_num_particles", "dfsph_b200_num_boundary_particles", "dfsph_b200_capacity",
    "dfsph_b200_eval_kernel", "dfsph_b200_bind_host_numa", "dfsph_b200_alloc_pinned", "dfsph_b200_free_pinned", "dfsph_b200_host_register",
    "dfsph_b200_host_unregister", "dfsph_b200_synchronize",
    "dfsph_b200_set_profiling", "dfsph_b200_get_profile", "dfsph_b200_timer_start", "dfsph_b200_timer_stop",
    "dfsph_b200_comm_get_unique_id", "dfsph_b200_comm_init", "dfsph_b200_p2p_export", "dfsph_b200_p2p_import",
    "dfsph_b200_p2p_disable",
]

PROF_CLASSES = ["sort", "build_neighbors", "init_sweep", "accel", "jacobi_div", "jacobi_press", "div_final",
                "press_init", "press_final"]


def lib_path(precision: str) -> str:
    # DFSPH_B200_LIB_<PRECISION> overrides the in-tree library (tuning experiments with alternative builds)
    override = os.environ.get(f"DFSPH_B200_LIB_{precision.upper()}")
    if override:
        return override
    return os.path.join(_PKG, f"libdfsph_b200_{precision}.so")


_LIBS = {}


def load(precision: str = "f32"):
    """Load libdfsph_b200_<precision>.so and declare the prototypes.  Raises if the extension is not built."""
    if precision in _LIBS:
        return _LIBS[precision]
    path = lib_path(precision)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the DFSPH hot path has no CPU fallback)")
    L = C.CDLL(path)
    P = C.c_void_p
    L.dfsph_b200_sizeof_real.restype = C.c_int
    L.dfsph_b200_version.restype = C.c_char_p
    L.dfsph_b200_default_config.argtypes = [C.POINTER(Config)]
    L.dfsph_b200_default_params.argtypes = [C.POINTER(Params)]
    L.dfsph_b200_create.argtypes = [C.POINTER(Config), C.POINTER(P)]
    L.dfsph_b200_destroy.argtypes = [P]
    L.dfsph_b200_last_error.argtypes = [P]
    L.dfsph_b200_last_error.restype = C.c_char_p
    L.dfsph_b200_set_fluid.argtypes = [P, C.c_uint64, P, P, P, P, C.c_double, C.c_double]
    L.dfsph_b200_add_boundary.argtypes = [P, C.c_uint64, P, P, C.c_int]
    L.dfsph_b200_compute_boundary_volume.argtypes = [P]
    L.dfsph_b200_set_params.argtypes = [P, C.POINTER(Params)]
    L.dfsph_b200_get_params.argtypes = [P, C.POINTER(Params)]
    L.dfsph_b200_step.argtypes = [P, C.POINTER(StepStats)]
    L.dfsph_b200_step_host.argtypes = [P, P, P, P, C.POINTER(StepStats)]
    L.dfsph_b200_download.argtypes = [P, C.c_int, P, C.c_size_t, C.c_int]
    L.dfsph_b200_upload.argtypes = [P, C.c_int, P, C.c_size_t, C.c_int]
    L.dfsph_b200_neighbors.argtypes = [P, C.c_int, P, P, P, C.c_uint64]
    L.dfsph_b200_search_and_density.argtypes = [P]
    L.dfsph_b200_num_particles.argtypes = [P]
    L.dfsph_b200_num_particles.restype = C.c_uint64
    L.dfsph_b200_num_boundary_particles.argtypes = [P]
    L.dfsph_b200_num_boundary_particles.restype = C.c_uint64
    L.dfsph_b200_capacity.argtypes = [P]
    L.dfsph_b200_capacity.restype = C.c_uint64
    L.dfsph_b200_eval_kernel.argtypes = [P, C.c_int, C.c_uint64, P, P, P]
    L.dfsph_b200_bind_host_numa.argtypes = [C.c_int]
    L.dfsph_b200_bind_host_numa.restype = C.c_int
    L.dfsph_b200_alloc_pinned.argtypes = [C.c_size_t]
    L.dfsph_b200_alloc_pinned.restype = P
    L.dfsph_b200_free_pinned.argtypes = [P]
    L.dfsph_b200_host_register.argtypes = [P, C.c_size_t]
    L.dfsph_b200_host_unregister.argtypes = [P]
    L.dfsph_b200_synchronize.argtypes = [P]
    L.dfsph_b200_set_profiling.argtypes = [P, C.c_int]
    L.dfsph_b200_get_profile.argtypes = [P, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.dfsph_b200_timer_start.argtypes = [P]
    L.dfsph_b200_timer_stop.argtypes = [P, C.POINTER(C.c_float)]
    L.dfsph_b200_comm_get_unique_id.argtypes = [P]
    L.dfsph_b200_comm_init.argtypes = [P, P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    L.dfsph_b200_p2p_export.argtypes = [P, P]
    L.dfsph_b200_p2p_import.argtypes = [P, P]
    L.dfsph_b200_p2p_disable.argtypes = [P]
    want = 4 if precision == "f32" else 8
    if L.dfsph_b200_sizeof_real() != want:
        raise RuntimeError(f"{path}: sizeof(Real) mismatch")
    _LIBS[precision] = L
    return L


class DFSPHError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dfsph_b200 error {code}: {msg}")
        self.code = code


def pinned_array(lib, shape, dtype):
    """numpy array backed by page-locked host memory (for the host-buffer path)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.dfsph_b200_alloc_pinned(max(n, 1))
    if not p:
        raise MemoryError("cudaMallocHost failed")
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr, p

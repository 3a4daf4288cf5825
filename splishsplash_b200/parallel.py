"""Host-side logic of the multi-GPU slab decomposition (one process per GPU, launched with torchrun): slab bounds,
per-rank scene selection, and the NCCL bootstrap.  torch.distributed is plumbing only (rendezvous + handing the NCCL
unique id to every rank + barriers / timing reductions in bench.py); the particle exchange itself runs inside the CUDA
library over NCCL send/recv on the library's own stream (csrc/multi_gpu.cuh).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

BIG = 1.0e300


def slab_bounds(x_min: float, x_max: float, world: int):
    """Equal-width slabs [lo, hi) along x covering [x_min, x_max); the outermost faces are open (+-1e300) so that
    particles leaving the initial extent stay owned by the first / last rank."""
    edges = np.linspace(float(x_min), float(x_max), world + 1)
    out = []
    for r in range(world):
        lo = -BIG if r == 0 else float(edges[r])
        hi = BIG if r == world - 1 else float(edges[r + 1])
        out.append((lo, hi))
    return out


def owner_of(x, bounds):
    """Rank owning each x coordinate."""
    x = np.asarray(x, dtype=np.float64)
    his = np.array([b[1] for b in bounds[:-1]], dtype=np.float64)
    return np.searchsorted(his, x, side="right")


def select_slab(scene, rank: int, world: int, bounds=None, halo_cells: float = 2.0, axis: int = 0):
    """Per-rank view of a global scene dict: owned fluid particles (with global ids), the boundary particles within
    `halo_cells` cells of the slab, the slab itself and the global cell-grid domain every rank must share."""
    x = np.asarray(scene["fluid_x"])
    bx = scene.get("boundary_x")
    if bounds is None:
        bounds = slab_bounds(float(x[:, axis].min()), float(np.nextafter(x[:, axis].max(), np.inf)), world)
    lo, hi = bounds[rank]
    own = np.nonzero((x[:, axis].astype(np.float64) >= lo) & (x[:, axis].astype(np.float64) < hi))[0]
    cell = 4.0 * scene["radius"] * (1.0 + 1.0e-5)
    out = dict(scene)
    out["fluid_x"] = np.ascontiguousarray(x[own])
    if scene.get("fluid_v") is not None:
        out["fluid_v"] = np.ascontiguousarray(np.asarray(scene["fluid_v"])[own])
    out["fluid_ids"] = own.astype(np.uint32)
    pts = [x.min(axis=0), x.max(axis=0)]
    if bx is not None and len(bx):
        b = np.asarray(bx)
        keep = (b[:, axis].astype(np.float64) >= lo - halo_cells * cell) & (b[:, axis].astype(np.float64) < hi + halo_cells * cell)
        out["boundary_x"] = np.ascontiguousarray(b[keep])
        out["boundary_keep"] = np.nonzero(keep)[0]
        pts += [b.min(axis=0), b.max(axis=0)]
    pts = np.asarray(pts, dtype=np.float64)
    out["domain"] = (pts.min(axis=0) - cell, pts.max(axis=0) + cell)
    out["slab"] = (lo, hi)
    out["slab_axis"] = axis
    out["bounds"] = bounds
    return out


def bootstrap_comm(ts, rank: int, world: int, slab, axis: int = 0):
    """Create the NCCL communicator of a TimeStepDFSPH_B200 context: rank 0 draws the unique id, torch.distributed
    broadcasts the 256 bytes, every rank calls dfsph_b200_comm_init.  Must run before set_fluid."""
    import torch.distributed as dist
    buf = (C.c_char * 256)()
    if rank == 0:
        ts._check(ts.lib.dfsph_b200_comm_get_unique_id(buf))
    obj = [bytes(buf.raw) if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(obj, src=0)
    idb = (C.c_char * 256).from_buffer_copy(obj[0])
    ts._check(ts.lib.dfsph_b200_comm_init(ts.ctx, idb, rank, world, int(axis), float(slab[0]), float(slab[1])))


def enable_p2p(ts, rank: int, world: int):
    """Switch the ghost refresh of a slab context to direct NVLink peer stores: all_gather the CUDA IPC blobs and hand
    every rank its neighbours'.  Must run after set_fluid (the buffers must exist)."""
    import torch.distributed as dist
    blob = (C.c_char * 512)()
    ok = ts.lib.dfsph_b200_p2p_export(ts.ctx, blob) == 0
    blobs = [None] * world
    dist.all_gather_object(blobs, (ok, bytes(blob.raw)))
    if all(b[0] for b in blobs):
        allb = (C.c_char * (512 * world)).from_buffer_copy(b"".join(b[1] for b in blobs))
        ok = ts.lib.dfsph_b200_p2p_import(ts.ctx, allb) == 0
    else:
        ok = False
    # every rank must end up on the same path: if peer mapping failed anywhere (no NVLink peer access, IPC disabled in
    # the container, more than 16 ranks) all ranks stay on NCCL send/recv + ncclAllReduce
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if not all(oks):
        ts.lib.dfsph_b200_p2p_disable(ts.ctx)
        return False
    return True


def build_b200_slab(scene_rank, precision, rank, world, kernel=4, device=0, boundary_V=None, p2p=None, **params):
    """TimeStepDFSPH_B200 for one rank of a slab-decomposed scene (see select_slab)."""
    from .solver import TimeStepDFSPH_B200
    ts = TimeStepDFSPH_B200(precision, scene_rank["radius"], kernel, device=device, domain=scene_rank["domain"])
    if params:
        ts.set(**params)
    bootstrap_comm(ts, rank, world, scene_rank["slab"], scene_rank.get("slab_axis", 0))
    ts.set_fluid(scene_rank["fluid_x"], scene_rank.get("fluid_v"), ids=scene_rank["fluid_ids"])
    bx = scene_rank.get("boundary_x")
    if bx is not None and len(bx):
        ts.add_boundary(bx, boundary_V)
        if boundary_V is None:
            ts.compute_boundary_volume()
    import os
    if p2p is None:
        p2p = os.environ.get("DFSPH_B200_P2P", "1") != "0"
    if p2p and world > 1:
        enable_p2p(ts, rank, world)
    return ts

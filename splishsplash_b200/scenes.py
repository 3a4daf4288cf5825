"""Synthetic scene generators for the DFSPH hot path.

No RNG anywhere: the particle distribution is the lattice itself, so a scene is fully described by its integer
lattice counts, the particle radius and the dtype.

* ``fluid_block``   restates the reference's ``SimulatorBase::createFluidBlocks`` (Simulator/SimulatorBase.cpp:1418-1526)
  for denseMode 0 (regular lattice) and denseMode 1: spacing d = 2r, first particle at ``min + d``,
  ``steps = round(L/d) - 1`` per axis, loop order x outer, y, z inner, arithmetic in ``Real``.
* ``box_boundary``  single layer of Akinci2012 boundary particles on the faces of an axis-aligned box
  (regular grid, default spacing 1.5 r).  The reference samples meshes (RegularTriangleSampling); here the box is
  sampled directly, which is the same kind of input for ``BoundaryModel_Akinci2012::initModel``
  (SPlisHSPlasH/BoundaryModel_Akinci2012.cpp:77-110).
* ``dam_break``     the measurement scene of BASELINE.json configs 2-5: fluid block in the first third of a closed tank.
"""
from __future__ import annotations

import numpy as np

# BASELINE.json / SURVEY.md 8d named particle counts -> lattice counts (x, y, z)
NAMED_BLOCKS = {
    "tiny": (10, 12, 10),          # 1 200      unit-test size
    "small": (20, 24, 20),         # 9 600
    "64k": (40, 40, 40),           # 64 000
    "1M": (100, 100, 100),         # 1 000 000  config 2
    "10M": (216, 216, 215),        # 10 031 040 configs 3 and 5 (per GPU)
    "50M": (464, 232, 464),        # 49 948 672 config 4
}


def fluid_lattice(counts, radius, start, dtype=np.float32):
    """Regular lattice with ``counts`` = (nx, ny, nz) particles, spacing 2r, first particle at ``start``.

    Position arithmetic follows SimulatorBase.cpp:1484 (``Vector3r(j*xshift, k*yshift, l*diam) + start`` in Real)."""
    dt = np.dtype(dtype).type
    nx, ny, nz = (int(c) for c in counts)
    diam = dt(2.0) * dt(radius)
    ax = [(np.arange(n, dtype=dtype) * diam + dt(s)).astype(dtype) for n, s in zip((nx, ny, nz), start)]
    x = np.empty((nx, ny, nz, 3), dtype=dtype)
    x[..., 0] = ax[0][:, None, None]
    x[..., 1] = ax[1][None, :, None]
    x[..., 2] = ax[2][None, None, :]
    return x.reshape(-1, 3)


def fluid_block(box_min, box_max, radius, dtype=np.float32, dense_mode=0):
    """Restatement of ``createFluidBlocks`` (Simulator/SimulatorBase.cpp:1418-1526), denseMode 0 or 1, no transform."""
    dt = np.dtype(dtype).type
    r = dt(radius)
    diam = dt(2.0) * r
    xshift = diam
    yshift = diam
    if dense_mode == 1:
        yshift = dt(np.sqrt(dt(3.0)) * r + dt(1.0e-9))
    elif dense_mode != 0:
        raise ValueError("dense_mode 0 or 1 only")
    bmin = np.asarray(box_min, dtype=dtype)
    bmax = np.asarray(box_max, dtype=dtype)
    diff = (bmax - bmin).astype(dtype)
    if dense_mode == 1:
        diff[0] -= diam
        diff[2] -= diam
    # C round(): half away from zero
    cround = lambda v: int(np.floor(abs(float(v)) + 0.5) * (1 if v >= 0 else -1))
    sx = cround(diff[0] / xshift) - 1
    sy = cround(diff[1] / yshift) - 1
    sz = cround(diff[2] / diam) - 1
    if sx <= 1 or sy <= 1 or sz <= 1:
        return np.zeros((0, 3), dtype=dtype)
    start = (bmin + diam).astype(dtype)
    j = np.arange(sx, dtype=dtype)[:, None, None]
    k = np.arange(sy, dtype=dtype)[None, :, None]
    l = np.arange(sz, dtype=dtype)[None, None, :]
    x = np.empty((sx, sy, sz, 3), dtype=dtype)
    x[..., 0] = (j * xshift).astype(dtype) + start[0]
    x[..., 1] = (k * yshift).astype(dtype) + start[1]
    x[..., 2] = (l * diam).astype(dtype) + start[2]
    if dense_mode == 1:
        even = (np.arange(sy) % 2 == 0)[None, :, None]
        x[..., 2] = np.where(even, x[..., 2] + r, x[..., 2]).astype(dtype)
        x[..., 0] = np.where(~even, x[..., 0] + r, x[..., 0]).astype(dtype)
    return x.reshape(-1, 3)


def box_boundary(box_min, box_max, radius, dtype=np.float32, spacing_factor=1.5, x_range=None, axis_range=None):
    """One layer of boundary particles on the six faces of an axis-aligned box, no duplicates on edges.
    x_range = (a, b) / axis_range = (axis, a, b): only the particles with a <= coordinate < b are generated (the per-rank
    portion of a long tank)."""
    if axis_range is not None and axis_range[0] != 0:
        # permute so that the filtered axis becomes x, generate, permute back
        ax_, a, b = axis_range
        perm = [ax_] + [k for k in range(3) if k != ax_]
        inv = np.argsort(perm)
        pts = box_boundary(np.asarray(box_min, dtype=np.float64)[perm], np.asarray(box_max, dtype=np.float64)[perm], radius, dtype,
                           spacing_factor, x_range=(a, b))
        return np.ascontiguousarray(pts[:, inv])
    if axis_range is not None:
        x_range = (axis_range[1], axis_range[2])
    bmin = np.asarray(box_min, dtype=np.float64)
    bmax = np.asarray(box_max, dtype=np.float64)
    s = spacing_factor * float(radius)
    n = [max(int(round((bmax[k] - bmin[k]) / s)), 1) for k in range(3)]
    ax = [np.linspace(bmin[k], bmax[k], n[k] + 1) for k in range(3)]
    if x_range is not None:
        a, b = x_range
        xs = ax[0]
        first, last = xs[0], xs[-1]
        inner = xs[1:-1]
        inner = inner[(inner >= a) & (inner < b)]
        ends = np.array([v for v in (first, last) if a <= v < b])
        parts = []

        def grid_(p, q, r):
            g = np.empty((len(p), len(q), len(r), 3), dtype=np.float64)
            g[..., 0] = p[:, None, None]
            g[..., 1] = q[None, :, None]
            g[..., 2] = r[None, None, :]
            return g.reshape(-1, 3)

        if len(ends):
            parts.append(grid_(ends, ax[1], ax[2]))
        if len(inner):
            parts.append(grid_(inner, ax[1][[0, -1]], ax[2]))
            parts.append(grid_(inner, ax[1][1:-1], ax[2][[0, -1]]))
        if not parts:
            return np.zeros((0, 3), dtype=dtype)
        return np.ascontiguousarray(np.concatenate(parts, axis=0).astype(dtype))

    def grid(a, b, c):
        g = np.empty((len(a), len(b), len(c), 3), dtype=np.float64)
        g[..., 0] = a[:, None, None]
        g[..., 1] = b[None, :, None]
        g[..., 2] = c[None, None, :]
        return g.reshape(-1, 3)

    parts = [
        grid(ax[0][[0, -1]], ax[1], ax[2]),                 # x faces (own all their edges)
        grid(ax[0][1:-1], ax[1][[0, -1]], ax[2]),           # y faces minus x edges
        grid(ax[0][1:-1], ax[1][1:-1], ax[2][[0, -1]]),     # z faces minus x and y edges
    ]
    return np.ascontiguousarray(np.concatenate(parts, axis=0).astype(dtype))


def dam_break(counts="tiny", radius=0.025, dtype=np.float32, tank_x_factor=3.0, tank_y_factor=1.5,
              spacing_factor=1.5, x_offset=0.0):
    """Dam-break block in a closed tank.

    ``counts``: key of NAMED_BLOCKS or an (nx, ny, nz) tuple.  The block's bounding box starts at the tank corner, so the
    nearest fluid particle sits one diameter from each adjacent wall.  Returns a dict with ``fluid_x`` (N,3),
    ``boundary_x`` (Nb,3), ``radius``, ``counts``, ``tank_min``, ``tank_max``."""
    if isinstance(counts, str):
        counts = NAMED_BLOCKS[counts]
    nx, ny, nz = counts
    d = 2.0 * radius
    block = np.array([(nx + 1) * d, (ny + 1) * d, (nz + 1) * d])
    tmin = np.array([x_offset, 0.0, 0.0])
    tmax = tmin + np.array([tank_x_factor * block[0], tank_y_factor * block[1], block[2]])
    fluid = fluid_lattice(counts, radius, tmin + d, dtype)
    bnd = box_boundary(tmin, tmax, radius, dtype, spacing_factor)
    return {"fluid_x": fluid, "boundary_x": bnd, "radius": float(radius), "counts": tuple(counts),
            "tank_min": tmin, "tank_max": tmax}


def dam_break_slab(rank, world, global_counts, radius=0.025, dtype=np.float32, tank_x_factor=3.0, tank_y_factor=1.5,
                   spacing_factor=1.5, halo_cells=2.0, axis=2):
    """Per-rank portion of ONE dam-break block of `global_counts` lattice particles cut into `world` slabs along `axis`
    (strong scaling, BASELINE config 4; also the building block of the weak-scaling scene).  Rank r generates ONLY its own
    lattice planes (global particle ids), its slab [lo, hi) and the boundary particles within `halo_cells` cells of the
    slab; the global scene is never materialised."""
    if isinstance(global_counts, str):
        global_counts = NAMED_BLOCKS[global_counts]
    if axis not in (0, 2):
        raise ValueError("axis 0 (x) or 2 (z)")
    g = [int(c) for c in global_counts]
    dt = np.dtype(dtype).type
    d = 2.0 * radius
    block = np.array([(g[0] + 1) * d, (g[1] + 1) * d, (g[2] + 1) * d])
    tmin = np.zeros(3)
    tmax = np.array([tank_x_factor * block[0], tank_y_factor * block[1], block[2]])
    # lattice planes of this rank along the slab axis (as even as possible)
    cuts = [(g[axis] * r) // world for r in range(world + 1)]
    idx = [np.arange(g[0]), np.arange(g[1]), np.arange(g[2])]
    idx[axis] = np.arange(cuts[rank], cuts[rank + 1])
    # the same position arithmetic as fluid_lattice for the global lattice (index * diam + start, in Real)
    diam = dt(2.0) * dt(radius)
    ax = [(i.astype(dtype) * diam + dt(d)).astype(dtype) for i in idx]
    x = np.empty((len(idx[0]), len(idx[1]), len(idx[2]), 3), dtype=dtype)
    x[..., 0] = ax[0][:, None, None]
    x[..., 1] = ax[1][None, :, None]
    x[..., 2] = ax[2][None, None, :]
    ids = ((idx[0][:, None, None].astype(np.int64) * g[1] + idx[1][None, :, None]) * g[2] + idx[2][None, None, :])
    # slab faces half-way between lattice planes
    face = lambda plane: d + (plane - 0.5) * d
    lo = -1.0e300 if rank == 0 else face(cuts[rank])
    hi = 1.0e300 if rank == world - 1 else face(cuts[rank + 1])
    cell = 4.0 * radius * (1.0 + 1.0e-5)
    bnd = box_boundary(tmin, tmax, radius, dtype, spacing_factor, axis_range=(axis, lo - halo_cells * cell, hi + halo_cells * cell))
    return {"fluid_x": x.reshape(-1, 3), "fluid_ids": ids.reshape(-1).astype(np.uint32), "boundary_x": bnd,
            "radius": float(radius), "counts": tuple(len(i) for i in idx), "slab": (lo, hi), "slab_axis": axis,
            "domain": (tmin - cell, tmax + cell), "tank_min": tmin, "tank_max": tmax,
            "global_particles": int(g[0]) * g[1] * g[2], "global_counts": tuple(g)}


def dam_break_weak(rank, world, counts="10M", radius=0.025, dtype=np.float32, axis=2, **kw):
    """Weak-scaling scene: `world` blocks of the named size side by side along `axis` in one tank, one
    slab per rank (see dam_break_slab).

    axis = 2 (default): the dam is replicated across the tank (z), slabs are perpendicular to the flow, so every GPU
    sees the same dam-break physics for the whole run (same iteration counts as the single-GPU block, no load drift).
    axis = 0: blocks side by side along the flow direction (SURVEY.md 8d wording): one long block; the larger global
    system needs more Jacobi iterations and the fluid drifts towards the last ranks as the dam collapses."""
    if isinstance(counts, str):
        counts = NAMED_BLOCKS[counts]
    g = list(counts)
    g[axis] *= world
    return dam_break_slab(rank, world, g, radius, dtype, axis=axis, **kw)


def rw_state_scene(dtype=np.float32):
    """Geometry of the reference's only shipped DFSPH + Akinci2012 + deterministic-sampling scene,
    data/Scenes/ReadWriteStateTest.json: box 1 x 1.5 x 1 centred at (0, 0.75, 0), fluid block
    [-0.25,0,-0.25]-[0.25,1,0.25] in denseMode 1, r = 0.025 (viscosity/vorticity off)."""
    r = 0.025
    fluid = fluid_block([-0.25, 0.0, -0.25], [0.25, 1.0, 0.25], r, dtype, dense_mode=1)
    bnd = box_boundary([-0.5, 0.0, -0.5], [0.5, 1.5, 0.5], r, dtype)
    return {"fluid_x": fluid, "boundary_x": bnd, "radius": r, "counts": None,
            "tank_min": np.array([-0.5, 0.0, -0.5]), "tank_max": np.array([0.5, 1.5, 0.5])}


# Solver settings of data/Scenes/DoubleDamBreak.json:15-33 (the values the scene file overrides; everything else default)
DOUBLE_DAM_BREAK_PARAMS = dict(minIterations=2, maxIterations=100, maxError=0.05, maxIterationsV=100, maxErrorV=0.1,
                               enableDivergenceSolver=1, cflMethod=1, cflFactor=1.0, cflMaxTimeStepSize=0.005,
                               viscosityMethod=1, viscosity=0.01)


def double_dam_break_scene(dtype=np.float32):
    """BASELINE config 1 as an Akinci2012 VARIANT: the fluid blocks, particle radius and box of
    data/Scenes/DoubleDamBreak.json (two blocks [-1.5,0,-1.5]-[-0.8,0.75,-0.8] and [0.8,0,0.8]-[1.5,0.75,1.5] ->
    2 x 13*14*13 = 4732 particles; UnitBox scaled 3.1 at (0,1.5,0)), with the box walls sampled by boundary particles
    instead of the shipped Bender2019 volume map (needs Discregrid, which is not in the image: SURVEY.md section 0 item 7).
    Solver settings: DOUBLE_DAM_BREAK_PARAMS."""
    r = 0.025
    a = fluid_block([-1.5, 0.0, -1.5], [-0.8, 0.75, -0.8], r, dtype, dense_mode=0)
    b = fluid_block([0.8, 0.0, 0.8], [1.5, 0.75, 1.5], r, dtype, dense_mode=0)
    lo, hi = np.array([-1.55, -0.05, -1.55]), np.array([1.55, 3.05, 1.55])
    bnd = box_boundary(lo, hi, r, dtype)
    return {"fluid_x": np.ascontiguousarray(np.concatenate([a, b], axis=0)), "boundary_x": bnd, "radius": r, "counts": None,
            "tank_min": lo, "tank_max": hi}

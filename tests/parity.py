"""Parity helpers shared by tests/, __graft_entry__.smoke() and tools/.  The oracle (oracle/_ref = the reference's own
DFSPH sources, or the C++ restatement oracle/liboracle_*.so) is the CHECKER; the thing checked is the CUDA library
reached through the C ABI (splishsplash_b200.solver.TimeStepDFSPH_B200).

Tolerances (BASELINE.json north_star): fields within 1e-4 (float) / 1e-10 (double) relative error, neighbour sets
bit-exact.  "Relative" is taken against the field's scale (max |reference| over the particles): element-wise
relative error is meaningless for fields that cancel to ~0 in the bulk (pressure acceleration, velocity at rest),
and the reference itself does not reproduce those element-wise when its neighbour order changes.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TOL = {"f32": 1.0e-4, "f64": 1.0e-10}

STEP_FIELDS = ["density", "factor", "advected density", "p / rho^2", "p_v / rho^2", "velocity", "position",
               "pressure acceleration"]


def dtype_of(precision):
    return np.float32 if precision == "f32" else np.float64


def make_oracle(scene, precision, kernel=4, grad_kernel=None, **params):
    """Reference-side simulation: the reference's own sources if oracle/_ref is built, else the C++ restatement."""
    from oracle import refsim
    if refsim.ref_available(precision):
        return refsim.build_ref_scene(scene, precision, kernel=kernel, grad_kernel=grad_kernel, **params), "reference"
    from oracle import portsim
    return portsim.build_port_scene(scene, precision, kernel=kernel, grad_kernel=grad_kernel, **params), "port"


def scaled_err(a, b, scale=None):
    """max |a-b| / max |b|  (0 if both are identically zero).  scale: the field's scale when b is only a part of it."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if scale is None:
        scale = float(np.max(np.abs(b))) if b.size else 0.0
    diff = float(np.max(np.abs(a - b))) if b.size else 0.0
    if scale == 0.0:
        return diff
    return diff / scale


def conditioned_errors(name, dev_f, ref_f, ref, h, keep=None):
    """Error measures for the ill-conditioned solver fields (see module docstring of tests/test_parity_gpu.py).

    kappa ("p / rho^2") is produced by  kappa <- max(kappa - 0.5 (s - A p) alpha/h^2, 0)  with s = 1 - rho_adv, so an
    absolute perturbation e of rho_adv (which is O(1)) moves the stored, h^2-scaled kappa by 0.5 alpha e.  When only a
    handful of particles are compressed at all, max|kappa| is itself rounding noise and the scale-relative error is
    meaningless; the well-conditioned statement is the error in DENSITY units  |d kappa| / alpha  (alpha = stored
    factor * h^2), which must be <= tol relative to the rest density (1).  The pressure acceleration is linear in
    kappa / h^2 and is judged through the velocity increment it causes, h |d a| / max|v|."""
    d = np.abs(np.asarray(dev_f, dtype=np.float64) - np.asarray(ref_f, dtype=np.float64))
    if keep is not None:      # particles in reach of a warm-start branch flip are left out (warm_start_flips)
        d = np.where(keep.reshape((-1,) + (1,) * (d.ndim - 1)), d, 0.0)
    if name == "p / rho^2":
        alpha = np.asarray(ref.field_by_id("factor"), dtype=np.float64) * h * h
        m = alpha > 0
        return float(np.max(d[m] / alpha[m])) if m.any() else 0.0
    if name == "pressure acceleration":
        vmax = float(np.max(np.abs(ref.field_by_id("velocity"))))
        return float(h * d.max() / vmax) if vmax > 0 else float(d.max())
    return None


# Fields that depend on the branch the pressure warm start takes (everything the pressure solve produces)
FLIP_FIELDS = ("p / rho^2", "velocity", "position", "pressure acceleration")


def warm_start_flips(ref, dev, scene, precision, pressure_iterations):
    """The reference's pressure warm start is a step function of rho_adv (TimeStepDFSPH.cpp:296-299:
    `if (densityAdv > 1.0) kappa = 0.5 min(kappa, 0.00025) / h^2; else kappa = 0`).  A particle whose rho_adv lies within
    rounding of 1.0 takes one branch or the other depending on the summation order of its neighbour sum (the reference's own
    AVX order follows CompactNSearch's list order), and the jump in kappa_0 is finite -- no implementation that does not
    reproduce the reference's float summation order bit for bit can follow it, and the reference itself would not under a
    different neighbour order.  This helper finds such particles: the branch differs between the reference and the device
    AND the reference's rho_adv is within 8 ulp of 1.0 (anything else is a real error and is NOT excused).  The disturbance
    spreads one support radius per sweep, two sweeps per iteration; the solver-produced fields are then compared on the
    particles outside that range only.  Returns (keep mask, number of flipped particles)."""
    ra = np.asarray(ref.field_by_id("advected density"), dtype=np.float64)
    da = np.asarray(dev.field("advected density"), dtype=np.float64)
    eps = float(np.finfo(dtype_of(precision)).eps)
    flipped = ((ra > 1.0) != (da > 1.0)) & (np.abs(ra - 1.0) <= 8.0 * eps)
    n = int(flipped.sum())
    if n == 0:
        return np.ones(len(ra), dtype=bool), 0
    from scipy.spatial import cKDTree
    x = np.asarray(ref.field_by_id("position"), dtype=np.float64)
    reach = 4.0 * float(scene["radius"]) * 2.0 * (max(int(pressure_iterations), 1) + 1)
    dist, _ = cKDTree(x[flipped]).query(x)
    return dist > reach, n


def neighbor_sets_by_id(counts, offsets, idx, row_ids, col_ids=None):
    """CSR in array order -> canonical form: (counts ordered by particle id, per-id ascending neighbour ids concatenated)."""
    counts = np.asarray(counts)
    cols = np.asarray(idx if col_ids is None else np.asarray(col_ids)[idx], dtype=np.uint32)
    rows = np.repeat(np.asarray(row_ids, dtype=np.uint32), counts)      # CSR rows are contiguous: offsets = cumsum(counts)
    order = np.lexsort((cols, rows))
    out_counts = np.zeros(len(counts), dtype=counts.dtype)
    out_counts[np.asarray(row_ids)] = counts
    if len(row_ids) and np.array_equal(np.sort(np.asarray(row_ids)), np.arange(len(row_ids))):
        return out_counts, cols[order]
    return counts[np.argsort(row_ids, kind="stable")], cols[order]


def sync_state(ref, dev):
    """Make the device start the next step from exactly the reference's current state."""
    dev.set_field("position", ref.field_by_id("position"))
    dev.set_field("velocity", ref.field_by_id("velocity"))
    dev.set_field("p / rho^2", ref.field_by_id("p / rho^2"))
    dev.set_field("p_v / rho^2", ref.field_by_id("p_v / rho^2"))
    dev.setValue("timeStepSize", ref.h)


def compare_step(precision, scene, steps=1, kernel=4, resync=True, tol=None, check_neighbors=True, grad_kernel=None, preroll=0,
                 collect=None, **params):
    """Run `steps` steps on the oracle and on the device from identical input states; compare every per-step field,
    the iteration counts, the new time step size and (first step) the neighbour sets.  Returns a result dict.

    preroll: the oracle first advances this many steps on its own (the device then starts from the oracle's state), so
    that the compared steps lie in a later phase of the scene (more solver iterations per step).
    collect: optional callable(step, ref, dev, stats) -> dict, stored per step under "extra" (run statistics).
    Where a conditioned measure replaces the scale-relative error of kappa / the pressure acceleration, the step record
    keeps both numbers under "conditioned" and the result counts the fallbacks in "conditioned_fallbacks"."""
    from splishsplash_b200.solver import build_b200_scene
    tol = TOL[precision] if tol is None else tol
    dev_only = {k: params.pop(k) for k in ("max_fluid_neighbors", "max_boundary_neighbors") if k in params}   # device table capacities
    ref, kind = make_oracle(scene, precision, kernel=kernel, grad_kernel=grad_kernel, **params)
    res = {"ok": True, "oracle": kind, "precision": precision, "steps": [], "max_err": {}}
    try:
        bx, bV = (None, None)
        if scene.get("boundary_x") is not None and len(scene["boundary_x"]):
            bx, bV = ref.boundary(0)
        # boundary volumes: device-computed, checked against the reference's (then the reference's are used so that
        # the step comparison starts from identical inputs)
        dev = build_b200_scene(scene, precision, kernel=kernel, grad_kernel=grad_kernel, **params, **dev_only)
        try:
            if bx is not None:
                # the reference z-sorts its boundary arrays once; match particles by position
                key = lambda a: [tuple(r) for r in a.tolist()]
                pos_to_V = dict(zip(key(bx), bV.tolist()))
                refV = np.array([pos_to_V[k] for k in key(np.asarray(scene["boundary_x"], dtype=dtype_of(precision)))])
                devV = dev.boundary_volume()
                res["max_err"]["boundary volume"] = scaled_err(devV, refV)
                if res["max_err"]["boundary volume"] > tol:
                    res["ok"] = False
            if check_neighbors:
                ref.search_and_density()
                rc, ro, ri = ref.neighbors(0, 0)
                rid = ref.ids()
                dc, do, di = dev.neighbors(0)
                did = dev.field("id", by_id=False)
                a = neighbor_sets_by_id(rc, ro, ri, rid, rid)
                b = neighbor_sets_by_id(dc, do, di, did, did)
                same = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
                res["neighbors_fluid_equal"] = bool(same)
                res["neighbor_pairs"] = int(len(ri))
                if bx is not None:
                    rcb, rob, rib = ref.neighbors(0, 1)
                    dcb, dob, dib = dev.neighbors(1)
                    # map reference boundary array indices -> insertion indices via positions
                    ins = {k: i for i, k in enumerate(key(np.asarray(scene["boundary_x"], dtype=dtype_of(precision))))}
                    rmap = np.array([ins[k] for k in key(bx)], dtype=np.uint32)
                    a = neighbor_sets_by_id(rcb, rob, rib, rid, rmap)
                    b = neighbor_sets_by_id(dcb, dob, dib, did, None)
                    sameb = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
                    res["neighbors_boundary_equal"] = bool(sameb)
                    same = same and sameb
                if not same:
                    res["ok"] = False
            if preroll:
                ref.step(int(preroll))
            for s in range(steps):
                if resync or s == 0:
                    sync_state(ref, dev)
                ref.step(1)
                st = dev.step(1)
                rec = {"ref_iter": (ref.iterations_v, ref.iterations), "dev_iter": (int(st.iterations_v), int(st.iterations)),
                       "ref_h": ref.h, "dev_h": float(st.time_step_size), "err": {}}
                keep, flips = warm_start_flips(ref, dev, scene, precision, rec["ref_iter"][1])
                if flips:
                    rec["warm_start_flips"] = flips
                    rec["compared_fraction"] = float(keep.mean())
                    res["warm_start_flips"] = res.get("warm_start_flips", 0) + flips
                for name in STEP_FIELDS:
                    df, rf = dev.field(name), ref.field_by_id(name)
                    if flips and name in FLIP_FIELDS:
                        scale = float(np.max(np.abs(np.asarray(rf, dtype=np.float64)))) if len(rf) else 0.0
                        e = scaled_err(np.asarray(df)[keep], np.asarray(rf)[keep], scale)
                    else:
                        e = scaled_err(df, rf)
                    if e > tol:
                        ce = conditioned_errors(name, df, rf, ref, ref.h, keep if flips and name in FLIP_FIELDS else None)
                        if ce is not None:
                            rec.setdefault("conditioned", {})[name] = (e, ce)
                            res["conditioned_fallbacks"] = res.get("conditioned_fallbacks", 0) + 1
                            e = min(e, ce)
                    rec["err"][name] = e
                    res["max_err"][name] = max(res["max_err"].get(name, 0.0), e)
                    if not (e <= tol):
                        res["ok"] = False
                if rec["ref_iter"] != rec["dev_iter"]:
                    res["ok"] = False
                if abs(rec["ref_h"] - rec["dev_h"]) > tol * abs(rec["ref_h"]):
                    res["ok"] = False
                if collect is not None:
                    rec["extra"] = collect(s, ref, dev, st)
                res["steps"].append(rec)
        finally:
            dev.close()
    finally:
        ref.destroy()
    worst = max(res["max_err"].items(), key=lambda kv: kv[1]) if res["max_err"] else ("-", 0.0)
    res["summary"] = (f"oracle={kind} N={len(scene['fluid_x'])} steps={steps} worst={worst[0]}:{worst[1]:.3e} "
                      f"iters={[(r['ref_iter'], r['dev_iter']) for r in res['steps']]} "
                      f"nbr_equal={res.get('neighbors_fluid_equal')}/{res.get('neighbors_boundary_equal')} "
                      f"conditioned_fallbacks={res.get('conditioned_fallbacks', 0)} warm_start_flips={res.get('warm_start_flips', 0)} "
                      f"ok={res['ok']}")
    return res

"""CPU tests of the multi-GPU host logic (world_size 2, gloo): slab partition, ghost-layer completeness, id plumbing.
The exchange itself (NCCL inside the CUDA library) is covered on GPUs by tests/test_multigpu_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from splishsplash_b200 import parallel, scenes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = scenes.dam_break("tiny", dtype=np.float64)
        mine = parallel.select_slab(sc, rank, world)
        lo, hi = mine["slab"]
        R = 4.0 * sc["radius"]
        cell = R * (1.0 + 1.0e-5)
        x = mine["fluid_x"]
        # what csrc/multi_gpu.cuh::k_select_exports exports: owned particles within one cell of a face
        exp = {}
        if rank > 0:
            exp[rank - 1] = (mine["fluid_ids"][x[:, 0] < lo + cell], x[x[:, 0] < lo + cell])
        if rank < world - 1:
            exp[rank + 1] = (mine["fluid_ids"][x[:, 0] >= hi - cell], x[x[:, 0] >= hi - cell])
        gathered = [None] * world
        dist.all_gather_object(gathered, exp)
        ghosts_ids = np.concatenate([g[rank][0] for g in gathered if rank in g] + [np.zeros(0, dtype=np.uint32)])
        # every true neighbour (|xi - xj| < R, brute force over the GLOBAL scene) of an owned particle is owned or a ghost
        gx = sc["fluid_x"]
        known = np.zeros(len(gx), dtype=bool)
        known[mine["fluid_ids"]] = True
        known[ghosts_ids] = True
        d2 = ((x[:, None, :] - gx[None, :, :]) ** 2).sum(-1)
        need = (d2 < R * R).any(axis=0)
        ok_cover = bool(known[need].all())
        # id plumbing used by parallel.bootstrap_comm
        obj = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        ok_id = obj[0] == bytes(range(128))
        counts = [None] * world
        dist.all_gather_object(counts, len(x))
        out.put((rank, ok_cover, ok_id, counts, len(gx), len(mine["boundary_x"]), len(sc["boundary_x"])))
    finally:
        dist.destroy_process_group()


def test_two_rank_slab_logic_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_cover, ok_id, counts, n, nb_mine, nb in res:
        assert ok_cover, f"rank {rank}: ghost layer misses a neighbour"
        assert ok_id
        assert sum(counts) == n            # partition: nobody lost, nobody duplicated
        assert 0 < nb_mine <= nb


def test_slab_bounds_and_owner():
    b = parallel.slab_bounds(0.0, 8.0, 4)
    assert b[0][0] == -parallel.BIG and b[-1][1] == parallel.BIG
    assert [round(v[1], 6) for v in b[:-1]] == [2.0, 4.0, 6.0]
    x = np.array([-5.0, 0.0, 1.999, 2.0, 5.5, 7.9, 100.0])
    assert parallel.owner_of(x, b).tolist() == [0, 0, 0, 1, 2, 3, 3]
    # every particle of a scene has exactly one owner and select_slab agrees with owner_of
    sc = scenes.dam_break("small", dtype=np.float32)
    world = 3
    parts = [parallel.select_slab(sc, r, world) for r in range(world)]
    ids = np.concatenate([p["fluid_ids"] for p in parts])
    assert len(ids) == len(sc["fluid_x"]) and len(np.unique(ids)) == len(ids)
    own = parallel.owner_of(sc["fluid_x"][:, 0], parts[0]["bounds"])
    for r, p in enumerate(parts):
        assert np.array_equal(np.sort(p["fluid_ids"]), np.nonzero(own == r)[0])
        assert np.array_equal(p["domain"][0], parts[0]["domain"][0])   # shared cell grid


@pytest.mark.parametrize("axis", [0, 2])
def test_weak_scaling_scene_tiles_the_global_block(axis):
    world = 4
    parts = [scenes.dam_break_weak(r, world, "tiny", dtype=np.float32, axis=axis) for r in range(world)]
    ids = np.concatenate([p["fluid_ids"] for p in parts])
    assert len(np.unique(ids)) == len(ids) == parts[0]["global_particles"] == 4 * 1200
    g = scenes.fluid_lattice(parts[0]["global_counts"], 0.025, (0.05, 0.05, 0.05), np.float32)
    x = np.empty_like(g)
    for p in parts:
        x[p["fluid_ids"]] = p["fluid_x"]
        lo, hi = p["slab"]
        assert p["slab_axis"] == axis
        assert ((p["fluid_x"][:, axis].astype(np.float64) >= lo) & (p["fluid_x"][:, axis].astype(np.float64) < hi)).all()
    assert np.array_equal(x, g)
    # slabs are contiguous and the per-rank boundary portions cover the whole tank boundary
    for a, b in zip(parts[:-1], parts[1:]):
        assert a["slab"][1] == b["slab"][0]
    full = scenes.box_boundary(parts[0]["tank_min"], parts[0]["tank_max"], 0.025, np.float32)
    got = np.unique(np.concatenate([p["boundary_x"] for p in parts]), axis=0)
    assert np.array_equal(got, np.unique(full, axis=0))


def test_strong_scaling_slabs_tile_one_block():
    world = 3
    parts = [scenes.dam_break_slab(r, world, (10, 12, 10), dtype=np.float64, axis=2) for r in range(world)]
    ids = np.concatenate([p["fluid_ids"] for p in parts])
    assert len(np.unique(ids)) == len(ids) == 1200 == parts[0]["global_particles"]
    g = scenes.fluid_lattice((10, 12, 10), 0.025, (0.05, 0.05, 0.05), np.float64)
    x = np.empty_like(g)
    for p in parts:
        x[p["fluid_ids"]] = p["fluid_x"]
    assert np.array_equal(x, g)
    assert [p["counts"][2] for p in parts] == [3, 3, 4]

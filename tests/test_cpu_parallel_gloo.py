"""CPU, world_size 2, gloo: the host-side routing logic of the multi-GPU grid (split sizes, the two
all-to-alls, (source rank, pose) run bookkeeping).  The device-side partition kernel is covered by
tests/test_gpu_parallel.py."""
import os
import socket

import numpy as np
import pytest

from octreelib_b200 import _native
from octreelib_b200.parallel import exchange_points, pose_starts, routing_layout, segments_from_counts, slab_boundaries


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _owner_numpy(points, edge, world):
    """numpy restatement of the device owner rule: hash of the integer cell coordinates."""
    lib = _native.lib()
    q = np.floor_divide(points, edge).astype(np.int64)
    return np.array([lib.ol_host_cell_owner(int(a), int(b), int(c), world) for a, b, c in q], dtype=np.int64)


def _worker(rank, world, port, n_poses, out_queue):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    # this rank holds poses rank, rank + world, ... (so the received runs are NOT pose-monotone)
    my_poses = list(range(rank, n_poses, world))
    clouds = [rng.random((50 + 10 * p, 3)) * 8 - 4 for p in my_poses]
    local = np.vstack(clouds)
    owner = _owner_numpy(local, 1.0, world)
    pose_of = np.concatenate([np.full(len(c), j) for j, c in enumerate(clouds)])
    order = np.lexsort((np.arange(len(local)), pose_of, owner))  # stable by (owner, run)
    send = torch.from_numpy(local[order].copy())
    counts = np.zeros((world, len(my_poses)), dtype=np.int64)
    np.add.at(counts, (owner, pose_of), 1)
    send_counts = routing_layout(counts, my_poses, n_poses)
    recv, recv_counts = exchange_points(send, send_counts)
    sizes, poses, first = segments_from_counts(recv_counts)
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(poses=my_poses, clouds=clouds))
    # expectation: for every source rank in order, for every pose it holds, the points I own, in input order
    exp, exp_sizes, exp_poses = [], [], []
    for src in range(world):
        for p, c in sorted(zip(gathered[src]["poses"], gathered[src]["clouds"]), key=lambda t: t[0]):
            mine = c[_owner_numpy(c, 1.0, world) == rank]
            if len(mine):
                exp.append(mine)
                exp_sizes.append(len(mine))
                exp_poses.append(p)
    ok = (np.vstack(exp) == recv.numpy()).all() and exp_sizes == sizes.tolist() and exp_poses == poses.tolist()
    ok = ok and all(f == 0 for f in first)  # every pose lives on exactly one source rank here
    total = torch.tensor([recv.shape[0]], dtype=torch.int64)
    dist.all_reduce(total)
    out_queue.put((rank, bool(ok), int(total.item()), sum(len(c) for g in gathered for c in g["clouds"])))
    dist.destroy_process_group()


def test_exchange_points_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, total, expected_total in results:
        assert ok, f"rank {rank}: received runs differ from the expectation"
        assert total == expected_total  # no point lost or duplicated


def test_segments_and_layout():
    counts = np.array([[3, 0, 2], [0, 4, 1]], dtype=np.int64)  # [src][pose]
    sizes, poses, first = segments_from_counts(counts)
    assert sizes.tolist() == [3, 2, 4, 1] and poses.tolist() == [0, 2, 1, 2] and first.tolist() == [0, 0, 0, 2]
    local = np.array([[1, 2], [3, 4]], dtype=np.int64)  # [owner][local run]; runs are poses 4 and 1
    assert routing_layout(local, [4, 1], 6).tolist() == [[0, 2, 0, 0, 1, 0], [0, 4, 0, 0, 3, 0]]


def test_owner_hash_is_balanced_and_deterministic():
    lib = _native.lib()
    rng = np.random.default_rng(0)
    q = rng.integers(-500, 500, size=(20000, 3))
    for world in (2, 4, 8):
        o = np.array([lib.ol_host_cell_owner(int(a), int(b), int(c), world) for a, b, c in q])
        share = np.bincount(o, minlength=world) / len(o)
        assert share.min() > 0.8 / world and share.max() < 1.2 / world
    assert lib.ol_host_cell_owner(1, 2, 3, 8) == lib.ol_host_cell_owner(1, 2, 3, 8)


def test_slab_boundaries_balance_and_are_rank_independent():
    """count quantiles of the leading cell coordinate from the per-rank (min, max, equal-width histogram) summaries"""
    rng = np.random.default_rng(1)
    world, bins = 8, 1024
    ix = [rng.integers(-50 + 120 * r, 260 + 120 * r, 30000) for r in range(world)]
    ix[3] = np.empty(0, dtype=np.int64)  # a rank without points
    g = np.zeros((world, 2 + bins), dtype=np.int64)
    for r, v in enumerate(ix):
        if len(v) == 0:
            g[r, 0], g[r, 1] = np.iinfo(np.int64).max, np.iinfo(np.int64).min
            continue
        lo, hi = v.min(), v.max()
        w = max(1, -(-(hi - lo + 1) // bins))
        g[r, 0], g[r, 1] = lo, hi
        g[r, 2:] = np.bincount((v - lo) // w, minlength=bins)
    b = slab_boundaries(g, world)
    assert len(b) == world - 1 and (np.diff(b) >= 0).all()
    allx = np.concatenate(ix)
    share = np.bincount(np.searchsorted(b, allx, side="right"), minlength=world) / len(allx)
    assert share.min() > 0.8 / world and share.max() < 1.25 / world, share
    assert (slab_boundaries(g.copy(), world) == b).all()
    # every point in ONE column of cells: no split is possible, every boundary coincides (documented limitation)
    g1 = np.zeros((2, 2 + bins), dtype=np.int64)
    g1[:, 0] = g1[:, 1] = 5
    g1[:, 2] = 1000
    assert len(set(slab_boundaries(g1, 2).tolist())) == 1


def test_pose_starts_restate_the_reference_batch_layout():
    """cuda_ransac.py:65-67 + grid.py:149-157: exclusive cumulative sum over the poses of a batch, cells in rank-major order"""
    sizes = np.array([[3, 1, 2, 0, 4], [1, 1, 1, 5, 0], [0, 2, 0, 1, 1]])  # [rank][pose]
    for ppb in (1, 2, 5):
        flat = []  # the reference's layout: batch, pose, then (rank-major) cells
        for first in range(0, 5, ppb):
            pos = 0
            for p in range(first, min(first + ppb, 5)):
                for r in range(3):
                    flat.append((p, r, pos))
                    pos += sizes[r, p]
        for r in range(3):
            got = pose_starts(sizes, r, ppb)
            assert got.tolist() == [next(pos for (p2, r2, pos) in flat if p2 == p and r2 == r) for p in range(5)]

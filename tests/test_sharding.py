"""CPU tests of the multi-GPU host logic: round-robin sharding of independent pairs and the single all-gather of
poses, exercised with a real 2-process gloo group (no GPU)."""
import os
import socket

import numpy as np
import pytest

from cvo_rgbd_b200 import sharding


def test_shard_pairs_partitions_every_pair_exactly_once():
    for n, w in [(500, 8), (7, 2), (3, 8), (0, 4), (148, 1)]:
        seen = np.concatenate([sharding.shard_pairs(n, w, r) for r in range(w)]) if w else np.array([])
        assert sorted(seen.tolist()) == list(range(n))
        assert max(len(sharding.shard_pairs(n, w, r)) for r in range(w)) == (sharding.max_shard_len(n, w) if n else 0)


def test_compose_trajectory_is_prefix_product():
    rng = np.random.default_rng(0)
    Ts = []
    for _ in range(5):
        T = np.eye(4)
        T[:3, 3] = rng.normal(0, 0.01, 3)
        Ts.append(T)
    traj = sharding.compose_trajectory(Ts)
    assert np.allclose(traj[-1][:3, 3], np.sum([T[:3, 3] for T in Ts], axis=0))


def _fake_pose(p):
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [p, 2 * p, -p]
    return T


def _worker(rank, world, port, n_pairs, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_pairs(n_pairs, world, rank)
        poses = np.stack([_fake_pose(p) for p in mine]) if len(mine) else np.zeros((0, 4, 4), np.float32)
        iters = (mine % 7).astype(np.int32)
        all_poses, all_iters = sharding.gather_poses(poses, iters, n_pairs)
        ok = all(np.array_equal(all_poses[p], _fake_pose(p)) for p in range(n_pairs))
        ok = ok and np.array_equal(all_iters, np.arange(n_pairs) % 7)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [9, 2, 1])
def test_gather_poses_world_size_2_gloo(n_pairs):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(2))
    assert got == [(0, True), (1, True)]


def _oracle_worker(rank, world, port, ids, q):
    import sys
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        poses = bench.oracle_poses_over_ranks(ids, 1, dist, torch, rank, world)
        q.put((rank, poses))
    finally:
        dist.destroy_process_group()


def test_parity_sample_dealt_over_ranks_world_size_2_gloo():
    """bench.py at N > 1: the oracle runs of the parity sample are dealt over the ranks (pair k of the sample -> rank
    k mod W) and summed into place by one all-reduce; every rank ends with every pose, in sample order."""
    import sys
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import bench
    ids = [0, 3, 5]  # BASELINE config 1 pairs (500 x 500 points, stock cvo schedule): seconds on the CPU
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_oracle_worker, args=(r, 2, port, ids, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = bench.cpu_reference_run(ids, threads=2, cfg=1)["poses"]
    for r in range(2):
        assert got[r].shape == (3, 4, 4)
        assert np.array_equal(got[r], np.asarray(want, np.float32))

"""Multi-GPU plumbing: independent frame pairs are dealt round-robin to ranks; the only
exchange is ONE all-gather of the 4x4 poses (+ iteration counts) at the end.  torch.distributed is used for the
collective only (backend nccl on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_pairs(n_pairs, world_size, rank):
    """Indices of the pairs rank `rank` registers: pair p -> rank p mod world_size."""
    return np.arange(rank, n_pairs, world_size, dtype=np.int64)


def max_shard_len(n_pairs, world_size):
    return (n_pairs + world_size - 1) // world_size


def gather_poses(local_poses, local_iters, n_pairs, device="cpu"):
    """All-gathers per-rank results into pair order.

    local_poses: [n_local, 4, 4] f32 (numpy), local_iters: [n_local] i32; every rank holds shard_pairs(n_pairs, W, r).
    Returns (poses [n_pairs, 4, 4], iters [n_pairs]) on every rank.  Message size: 64 B + 4 B per pair."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(local_poses, np.float32).reshape(-1, 4, 4), np.asarray(local_iters, np.int32)
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = max_shard_len(n_pairs, world)
    buf = torch.zeros((cap, 17), dtype=torch.float32)
    mine = shard_pairs(n_pairs, world, rank)
    assert len(mine) == len(local_poses) == len(local_iters)
    if len(mine):
        buf[:len(mine), :16] = torch.from_numpy(np.ascontiguousarray(local_poses, np.float32).reshape(-1, 16))
        buf[:len(mine), 16] = torch.from_numpy(np.asarray(local_iters, np.float32))
    buf = buf.to(device)
    out = torch.empty((world, cap, 17), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(out.view(-1), buf.view(-1))
    out = out.cpu().numpy()
    poses = np.zeros((n_pairs, 4, 4), np.float32)
    iters = np.zeros(n_pairs, np.int32)
    for r in range(world):
        idx = shard_pairs(n_pairs, world, r)
        poses[idx] = out[r, :len(idx), :16].reshape(-1, 4, 4)
        iters[idx] = out[r, :len(idx), 16].astype(np.int32)
    return poses, iters


def compose_trajectory(relative_poses):
    """accum_transform *= transform over a sequence (src/cvo.cpp:414): prefix product on the host."""
    acc = np.eye(4)
    out = []
    for T in relative_poses:
        acc = acc @ np.asarray(T, np.float64)
        out.append(acc.copy())
    return np.array(out)

"""Multi-GPU batching of the front end: one process per GPU, frames sharded by rank.

The per-keyframe path is embarrassingly parallel given the odometry poses; the only
coupling between neighbouring frames is the scan-matching window (frame i is matched against frames
i-window .. i-1), so every shard carries `window` halo frames in front of it whose clouds are recomputed
locally and whose results are dropped.  torch.distributed (NCCL on GPUs, gloo on CPU for tests) is used
only at the edges: scatter the backlog from rank 0, gather the 64-byte-per-frame results.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous, near-equal blocks: list of (start, end) for every rank."""
    per, extra = divmod(n, world)
    out, s = [], 0
    for r in range(world):
        e = s + per + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def shard_with_halo(n, world, window):
    """(halo_start, start, end) per rank: rank r processes frames [halo_start, end) and reports [start, end)."""
    return [(max(0, s - window), s, e) for s, e in shard_bounds(n, world)]


def scatter_backlog(frames, poses, window, device="cpu", src=0):
    """Rank `src` holds frames uint8 [n,R,B] and poses float64 [n,3] (others pass None); every rank
    returns (frames_local, poses_local, n_halo) as torch tensors on `device`."""
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = [None]
    if rank == src:
        meta = [(tuple(frames.shape), window)]
    dist.broadcast_object_list(meta, src=src)
    (n, R, B), window = meta[0]
    plan = shard_with_halo(n, world, window)
    h, s, e = plan[rank]
    f_local = torch.empty((e - h, R, B), dtype=torch.uint8, device=device)
    p_local = torch.empty((e - h, 3), dtype=torch.float64, device=device)
    if rank == src:
        ft = torch.as_tensor(frames).to(device)
        pt = torch.as_tensor(poses, dtype=torch.float64).to(device)
        reqs = []
        for r, (hh, ss, ee) in enumerate(plan):
            if r == src:
                f_local.copy_(ft[hh:ee])
                p_local.copy_(pt[hh:ee])
            else:
                reqs.append(dist.isend(ft[hh:ee].contiguous(), dst=r))
                reqs.append(dist.isend(pt[hh:ee].contiguous(), dst=r))
        for q in reqs:
            q.wait()
    else:
        dist.recv(f_local, src=src)
        dist.recv(p_local, src=src)
    return f_local, p_local, s - h


def gather_results(local, n_halo, n_total, dst=0):
    """local: dict of per-frame numpy arrays for the rank's [halo_start, end) frames.  Rank `dst` gets the
    concatenation over ranks in frame order with halos removed (others get None).  The first frame of
    every shard but the first keeps its halo-informed result; frame 0 of the backlog stays "skipped"."""
    rank, world = dist.get_rank(), dist.get_world_size()
    own = {k: np.ascontiguousarray(v[n_halo:]) for k, v in local.items()}
    gathered = [None] * world if rank == dst else None
    dist.gather_object(own, gathered, dst=dst)
    if rank != dst:
        return None
    out = {k: np.concatenate([g[k] for g in gathered]) for k in own}
    assert all(len(v) == n_total for v in out.values())
    return out

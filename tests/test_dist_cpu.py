"""CPU-only, world_size 2 over gloo: the multi-GPU sharding plumbing (scatter with halo, gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sonar_slam_b200 import dist as sdist


def test_shard_bounds_and_halo():
    assert sdist.shard_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert sdist.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    plan = sdist.shard_with_halo(100, 4, 3)
    assert plan[0] == (0, 0, 25) and plan[1] == (22, 25, 50) and plan[3] == (72, 75, 100)
    covered = sorted(i for _, s, e in plan for i in range(s, e))
    assert covered == list(range(100))


def _fake_frontend(frames, poses, window):
    """Stand-in with the same dependency structure as the scan matcher: frame i's result depends on
    frames i-window..i (here: a checksum), so a missing halo would change the answer."""
    f = frames.numpy().astype(np.int64).reshape(len(frames), -1).sum(1)
    out = np.zeros(len(f), np.int64)
    for i in range(len(f)):
        out[i] = f[max(0, i - window):i + 1].sum() * 7 + int(poses[i, 0].item() * 1000)
    return {"checksum": out, "index": np.round(poses[:, 1].numpy()).astype(np.int64)}


def _worker(rank, world, port, n, window, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 255, (n, 8, 16)).astype(np.uint8) if rank == 0 else None
    poses = np.c_[rng.random(n), np.arange(n), np.zeros(n)] if rank == 0 else None
    fl, pl, n_halo = sdist.scatter_backlog(frames, poses, window)
    local = _fake_frontend(fl, pl, window)
    res = sdist.gather_results(local, n_halo, n)
    if rank == 0:
        want = _fake_frontend(torch.as_tensor(frames), torch.as_tensor(poses), window)
        q.put((np.array_equal(res["checksum"], want["checksum"]), np.array_equal(res["index"], np.arange(n))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,window", [(37, 3), (5, 3), (64, 1)])
def test_scatter_compute_gather_world2(n, window):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, window, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(60)
    assert ok == (True, True)

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


# ------------------------------------------------------------------ config 5: backlog of independent scan matches
def _fake_icp(src, tgt, guess, prm):
    """Stand-in solver with the real one's signature: the packed 48-byte record is a checksum of the problem's own
    clouds and guess, so a record that reached the wrong rank, chunk or slot changes the answer."""
    n = src.shape[0]
    out = torch.zeros((n, sdist.RESULT_WORDS), dtype=torch.int32)
    if n == 0:
        return out
    out[:, 0] = (src.reshape(n, -1).sum(1) * 8).round().to(torch.int32)
    out[:, 1] = (tgt.reshape(n, -1).sum(1) * 8).round().to(torch.int32)
    out[:, 2] = (guess.reshape(n, -1)[:, 2] * 1000).round().to(torch.int32)
    out[:, 9] = prm
    return out


def _pairs_worker(rank, world, port, P, chunks, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ns, nt = 5, 11
    src = tgt = guess = None
    if rank == 0:
        g = torch.Generator().manual_seed(4)
        src = torch.randint(0, 64, (P, ns, 2), generator=g).float() / 8
        tgt = torch.randint(0, 64, (P, nt, 2), generator=g).float() / 8
        guess = torch.eye(3).repeat(P, 1, 1)
        guess[:, 0, 2] = torch.arange(P).float()
    a = sdist.run_pair_backlog(P, ns, nt, 7, src, tgt, guess, chunks=chunks, icp_fn=_fake_icp, device="cpu")
    shard = sdist.scatter_pairs(P, ns, nt, src, tgt, guess, device="cpu")
    b = sdist.gather_pair_results(_fake_icp(*shard, 7), P)
    if rank == 0:
        want = _fake_icp(src, tgt, guess, 7)
        q.put((bool(torch.equal(a, want)), bool(torch.equal(b, want)),
               bool(torch.equal(sdist.unpack_results(a)["iterations"], want[:, 9]))))
    else:
        assert a is None and b is None
    dist.destroy_process_group()


@pytest.mark.parametrize("P,chunks", [(37, 3), (3, 4), (64, 8), (1, 2)])
def test_pair_backlog_scatter_solve_gather_world2(P, chunks):
    plan = sdist.chunk_plan(P, 2, chunks)
    assert sorted(i for r in plan for a, b in r for i in range(a, b)) == list(range(P))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pairs_worker, args=(r, 2, port, P, chunks, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(60)
    assert ok == (True, True, True)


def test_pack_unpack_results_roundtrip():
    T = torch.randn(5, 3, 3)
    res = dict(T=T, iterations=torch.arange(5, dtype=torch.int32), inliers=torch.arange(5, dtype=torch.int32) * 3,
               status=torch.zeros(5, dtype=torch.int32))
    back = sdist.unpack_results(sdist.pack_results(res))
    assert torch.equal(back["T"], T) and torch.equal(back["inliers"], res["inliers"])

"""Global-initialisation cost (row N2) timing: grid build, single evaluations (what scipy.shgo issues one by one)
and dense batches of candidate poses, next to the CPU restatement of the reference closure (slam.py:541-568).
Run on the GPU box:  PYTHONPATH=. python tools/bench_globalinit.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import globalinit_ref as gref  # CPU leg only
from sonar_slam_b200 import _lib, synth

src, tgt, _ = synth.make_icp_pair(3)  # 2 000 / 20 000 points (config 3)
rng = np.random.default_rng(0)
ctx = _lib.default_context()
res = {}

t0 = time.perf_counter()
grid, xmin, ymin, resolution, hs = gref.target_grid(tgt, 0.5)
res["cpu_grid_build_ms"] = (time.perf_counter() - t0) * 1e3
for _ in range(3):
    cm = _lib.CostMap(ctx, tgt, xmin, ymin, resolution, grid.shape[0], grid.shape[1], hs)
t0 = time.perf_counter()
for _ in range(20):
    cm = _lib.CostMap(ctx, tgt, xmin, ymin, resolution, grid.shape[0], grid.shape[1], hs)
res["gpu_grid_build_ms_incl_h2d_sync"] = (time.perf_counter() - t0) / 20 * 1e3
res["grid"] = list(grid.shape)
assert np.array_equal(cm.grid(), grid)
cm.set_source(src)


def rows(K):
    x = rng.uniform(-1, 1, (K, 3)) * [1.0, 1.0, 0.1]
    c, s = np.cos(x[:, 2]), np.sin(x[:, 2])
    return np.stack([c, -s, s, c, x[:, 0], x[:, 1]], 1).astype(np.float32), x


# single evaluations through the host entry point (H2D 24 B + launch + D2H 4 B + sync each)
tf, xs = rows(2000)
for i in range(50):
    cm.score(tf[i:i + 1])
t0 = time.perf_counter()
for i in range(2000):
    cm.score(tf[i:i + 1])
res["gpu_single_eval_us"] = (time.perf_counter() - t0) / 2000 * 1e6
t0 = time.perf_counter()
for i in range(200):
    gref.cost_of_transform(grid, xmin, ymin, resolution, src, gref.Pose2(*xs[i]))
res["cpu_single_eval_us"] = (time.perf_counter() - t0) / 200 * 1e6

import torch  # device timing of the batch kernel
torch.cuda.set_device(0)
for K in (4096, 65536, 1048576):
    tf, _ = rows(K)
    d_tf = torch.from_numpy(tf).cuda()
    d_src = torch.from_numpy(src).cuda()
    d_cost = torch.empty(K, dtype=torch.int32, device="cuda")
    tctx = _lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        cm.score_dev(d_src.data_ptr(), len(src), d_tf.data_ptr(), K, d_cost.data_ptr(), ctx=tctx)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cm.score_dev(d_src.data_ptr(), len(src), d_tf.data_ptr(), K, d_cost.data_ptr(), ctx=tctx)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    host = cm.score(tf[:4096])
    assert np.array_equal(host, d_cost[:4096].cpu().numpy())
    res[f"gpu_batch_{K}"] = dict(ms=ts[len(ts) // 2], poses_per_s=K / ts[len(ts) // 2] * 1e3,
                                 point_lookups_per_s=K * len(src) / ts[len(ts) // 2] * 1e3)
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_globalinit.json", "w"), indent=1)

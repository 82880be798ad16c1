"""BASELINE config 5 / config 3 at scale: independent (source, target) ICP problems sharded over the GPUs of one box.

    python tools/bench_pairs.py [--pairs P] [--ns 2000] [--nt 20000] [--iters 20] [--checkers]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_pairs.py --pairs P

Every rank solves its own P problems (weak scaling; `--pairs 10000` on 8 GPUs is the 80 k-pair configuration), built
from 16 seeded scenes of sonar_slam_b200.synth.make_icp_pair repeated across the batch (the kernel keeps no state
between problems, so repetition does not help it); no data-path collective, NCCL only for the barrier and the max over
ranks of the device time.  Prints one JSON line on rank 0.  No CPU path: needs a GPU."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1184, help="problems per GPU (default: 8 waves of 148)")
    ap.add_argument("--ns", type=int, default=2000)
    ap.add_argument("--nt", type=int, default=20000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--checkers", action="store_true", help="shipped icp.yaml checkers instead of a fixed count")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    from sonar_slam_b200 import _lib, ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench_pairs.py: no CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    P = a.pairs
    scenes = [synth.make_icp_pair(1000 * rank + s, n_source=a.ns, n_target=a.nt)[:2] for s in range(16)]
    src = np.concatenate([scenes[i % 16][0] for i in range(P)])
    tgt = np.concatenate([scenes[i % 16][1] for i in range(P)])
    so = np.zeros(P + 1, np.int32)
    so[1:] = np.cumsum([len(scenes[i % 16][0]) for i in range(P)])
    to = np.zeros(P + 1, np.int32)
    to[1:] = np.cumsum([len(scenes[i % 16][1]) for i in range(P)])
    dev = lambda x: torch.from_numpy(x).cuda()
    sp, so_d, tp, to_d = dev(src), dev(so), dev(tgt), dev(to)
    guess = torch.eye(3, device="cuda").repeat(P, 1, 1).contiguous()
    prm = _lib.IcpParams() if a.checkers else _lib.IcpParams(smooth_length=0, max_iterations=a.iters)
    ns_max = max(len(s[0]) for s in scenes)
    nt_max = max(len(s[1]) for s in scenes)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        out = ops.icp(sp, so_d, tp, to_d, guess, ns_max, nt_max, prm)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = ops.icp(sp, so_d, tp, to_d, guess, ns_max, nt_max, prm)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    stats = torch.tensor([float((out["status"] == 0).sum()), float(out["iterations"].float().mean())],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    if rank == 0:
        t = float(ms.item()) * 1e-3
        print(json.dumps({
            "metric": "ICP pairs/sec", "value": world * P * a.steps / t, "unit": "pairs/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": t * 1e3 / a.steps, "higher_is_better": True,
            "scaling": "weak", "data": "synthetic", "dtype": "f32",
            "config": {"workload": "config5: independent ICP problems sharded by rank", "pairs_per_gpu": P,
                       "source_points": a.ns, "target_points": a.nt,
                       "mode": "icp.yaml checkers" if a.checkers else f"fixed {a.iters} iterations",
                       "converged": int(stats[0].item()), "mean_iterations": float(stats[1].item() / world)}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

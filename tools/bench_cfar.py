"""Quick CFAR-only timing (config 2: 4096 x 512 x 512 frames resident in HBM)."""
import json
import sys

import torch

from sonar_slam_b200 import ops

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
DATA = sys.argv[2] if len(sys.argv) > 2 else "replay"   # replay: bench.py's frames (speckle + echoes); noise: speckle
TAU = 2.749063720096473
torch.cuda.set_device(0)
if DATA == "replay":
    from sonar_slam_b200 import synth
    imgs = synth.make_trajectory_frames(F, seed=0, device="cuda")["frames"]
else:
    g = torch.Generator(device="cuda").manual_seed(0)
    imgs = torch.empty((F, 512, 512), dtype=torch.uint8, device="cuda")
    for i in range(0, F, 256):
        u = torch.rand((min(256, F - i), 512, 512), device="cuda", generator=g).clamp_min(1e-7)
        imgs[i:i + 256] = torch.clamp(torch.round(18.0 * torch.sqrt(-2.0 * torch.log(u))), 0, 255).to(torch.uint8)
print("data:", DATA, "cells >= 66:", float((imgs[:64] >= 66).float().mean()))
res = {}
for name, x in (("f32", imgs.float()), ("u8", imgs)):
    for outs in (dict(want_mask=True), dict(want_mask=False, want_bits=True), dict(want_mask=True, want_bits=True)):
        for alg in ("SOCA", "CA"):
            for _ in range(3):
                ops.cfar(x, alg, 20, 5, TAU, gate=65, **outs)
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.cfar(x, alg, 20, 5, TAU, gate=65, **outs)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            es = 4 if name == "f32" else 1
            byts = F * 512 * 512 * (es + (1 if outs.get("want_mask") else 0) + (0.125 if outs.get("want_bits") else 0))
            key = f"{name}_{alg}_{'mask' if outs.get('want_mask') else ''}{'bits' if outs.get('want_bits') else ''}"
            res[key] = dict(ms_best=ts[0], ms_med=ts[5], GBs_best=byts / ts[0] / 1e6, frames_per_s=F / ts[0] * 1e3)
            print(key, res[key], flush=True)
json.dump(res, open("gpurun_out/bench_cfar.json", "w"), indent=1)

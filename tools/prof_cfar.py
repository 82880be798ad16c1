"""Minimal CFAR launch sequence for ncu:  python tools/prof_cfar.py [F] [f32|u8] [mask|bits|maskbits] [noise|replay]
`replay` = frames of the synthetic bag replay bench.py uses (speckle + wall echoes); `noise` = speckle only."""
import sys
import torch
from sonar_slam_b200 import ops, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dt = sys.argv[2] if len(sys.argv) > 2 else "f32"
mode = sys.argv[3] if len(sys.argv) > 3 else "mask"
data = sys.argv[4] if len(sys.argv) > 4 else "replay"
torch.cuda.set_device(0)
if data == "replay":
    imgs = synth.make_trajectory_frames(F, seed=0, device="cuda")["frames"]
else:
    g = torch.Generator(device="cuda").manual_seed(0)
    imgs = torch.empty((F, 512, 512), dtype=torch.uint8, device="cuda")
    for i in range(0, F, 256):
        u = torch.rand((min(256, F - i), 512, 512), device="cuda", generator=g).clamp_min(1e-7)
        imgs[i:i + 256] = torch.clamp(torch.round(18.0 * torch.sqrt(-2.0 * torch.log(u))), 0, 255).to(torch.uint8)
x = imgs.float() if dt == "f32" else imgs
kw = dict(want_mask="mask" in mode, want_bits="bits" in mode)
for _ in range(3):
    ops.cfar(x, "SOCA", 20, 5, 2.749063720096473, gate=65, **kw)
torch.cuda.synchronize()

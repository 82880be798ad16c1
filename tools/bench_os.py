import torch
from sonar_slam_b200 import ops
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(0)
F = 1024
u = torch.rand((F, 512, 512), device="cuda", generator=g).clamp_min(1e-7)
imgs = torch.clamp(torch.round(18.0 * torch.sqrt(-2.0 * torch.log(u))), 0, 255).to(torch.uint8)
for name, x in (("u8", imgs), ("f32", imgs[:64].float())):
    for _ in range(2): ops.cfar(x, "OS", 20, 5, 9.137608674642355, k=10, gate=65)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.cfar(x, "OS", 20, 5, 9.137608674642355, k=10, gate=65); e1.record(); torch.cuda.synchronize()
    print(name, "frames", x.shape[0], "ms", e0.elapsed_time(e1), "frames/s", x.shape[0] / e0.elapsed_time(e1) * 1e3)

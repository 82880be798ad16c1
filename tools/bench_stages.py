"""Per-stage timings of the front end on synthetic data (development aid, not the bench contract)."""
import sys
import numpy as np
import torch
from sonar_slam_b200 import _lib, ops, synth
from sonar_slam_b200.bruce_slam.feature_extraction import FeatureExtraction

F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.cuda.set_device(0)
ctx = ops.context(0)

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), r

# frames: speckle + arcs generated on device
g = torch.Generator(device="cuda").manual_seed(0)
imgs = torch.empty((F, 512, 512), dtype=torch.uint8, device="cuda")
for i in range(0, F, 256):
    n = min(256, F - i)
    u = torch.rand((n, 512, 512), device="cuda", generator=g).clamp_min(1e-7)
    x = 18.0 * torch.sqrt(-2.0 * torch.log(u))
    for f in range(n):
        for _ in range(6):
            r0 = int(torch.randint(30, 479, (1,)).item()); w = int(torch.randint(30, 111, (1,)).item()); b0 = int(torch.randint(0, 512 - w, (1,)).item())
            x[f, r0:r0 + 3, b0:b0 + w] += 90 + 110 * float(torch.rand(1).item())
    imgs[i:i + n] = torch.clamp(torch.round(x), 0, 255).to(torch.uint8)
geo = FeatureExtraction()
geo.generate_map_xy(synth.Ping(0, None, 30.0 / 512, 512, synth.bearings_oculus(512)))
maps = _lib.Maps(ctx, geo.map_x, geo.map_y, 512, 512, geo.width, geo.height)
TAU = 2.749063720096473
t, det = timeit(lambda: ops.cfar(imgs, "SOCA", 20, 5, TAU, gate=65, want_mask=False, want_bits=True))
print(f"cfar u8 bits      {t:8.3f} ms  {F / t * 1e3:12.0f} frames/s")
CAP = 8192
t, cp = timeit(lambda: ops.cart_points(maps, bits=det["bits"], capacity=CAP))
cnt = cp["count"]
print(f"cart_points       {t:8.3f} ms  {F / t * 1e3:12.0f} frames/s   pts/frame mean {cnt.float().mean().item():.0f} max {cnt.max().item()}")
# pack clouds
def pack():
    c = torch.clamp(cnt, max=CAP)
    off = torch.zeros(F + 1, dtype=torch.int32, device="cuda"); off[1:] = torch.cumsum(c, 0)
    sel = torch.arange(CAP, device="cuda")[None, :] < c[:, None]
    return cp["xy"][sel].contiguous(), off, int(c.max().item())
t, (pts, off, nmax) = timeit(pack)
print(f"pack (torch)      {t:8.3f} ms")
t, ds = timeit(lambda: ops.downsample(pts, off, nmax, 0.5))
print(f"downsample        {t:8.3f} ms  {F / t * 1e3:12.0f} frames/s   mean out {ds['count'].float().mean().item():.0f}")
def pack2(res):
    c = res["count"]; o2 = torch.zeros(F + 1, dtype=torch.int32, device="cuda"); o2[1:] = torch.cumsum(c, 0)
    idx = torch.arange(pts.shape[0], device="cuda"); start = off[:-1].long()
    owner = torch.bucketize(idx, off[1:].long(), right=True); keep = (idx - start[owner]) < c[owner]
    return res["pts"][keep].contiguous(), o2, int(c.max().item())
p2, o2, n2 = pack2(ds)
t, ro = timeit(lambda: ops.remove_outlier(p2, o2, n2, 1.0, 5))
print(f"remove_outlier    {t:8.3f} ms  {F / t * 1e3:12.0f} frames/s   mean out {ro['count'].float().mean().item():.0f}")
# ICP config 3
pairs = [synth.make_icp_pair(s)[:2] for s in range(8)]
src = np.concatenate([pairs[i % 8][0] for i in range(P)]); tgt = np.concatenate([pairs[i % 8][1] for i in range(P)])
so = np.zeros(P + 1, np.int32); so[1:] = np.cumsum([len(pairs[i % 8][0]) for i in range(P)])
to = np.zeros(P + 1, np.int32); to[1:] = np.cumsum([len(pairs[i % 8][1]) for i in range(P)])
sp, tp = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
so, to = torch.from_numpy(so).cuda(), torch.from_numpy(to).cuda()
gs = torch.eye(3, device="cuda").repeat(P, 1, 1).contiguous()
for name, prm in (("fixed20", _lib.IcpParams(smooth_length=0, max_iterations=20)), ("checkers", _lib.IcpParams())):
    t, out = timeit(lambda: ops.icp(sp, so, tp, to, gs, 2000, 20000, prm), n=3, warm=1)
    print(f"icp 2k/20k {name:9s} {t:8.3f} ms  {P / t * 1e3:12.0f} pairs/s   iters mean {out['iterations'].float().mean().item():.1f}")
# ICP small (pipeline-like): 400 vs 1500
pairs = [synth.make_icp_pair(100 + s, n_source=400, n_target=1500)[:2] for s in range(8)]
P2 = 4 * P
src = np.concatenate([pairs[i % 8][0] for i in range(P2)]); tgt = np.concatenate([pairs[i % 8][1] for i in range(P2)])
so = np.zeros(P2 + 1, np.int32); so[1:] = np.cumsum([len(pairs[i % 8][0]) for i in range(P2)])
to = np.zeros(P2 + 1, np.int32); to[1:] = np.cumsum([len(pairs[i % 8][1]) for i in range(P2)])
sp, tp = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
so, to = torch.from_numpy(so).cuda(), torch.from_numpy(to).cuda()
gs = torch.eye(3, device="cuda").repeat(P2, 1, 1).contiguous()
for name, prm in (("fixed20", _lib.IcpParams(smooth_length=0, max_iterations=20)), ("checkers", _lib.IcpParams())):
    t, out = timeit(lambda: ops.icp(sp, so, tp, to, gs, 400, 1500, prm), n=3, warm=1)
    print(f"icp 400/1500 {name:9s} {t:8.3f} ms  {P2 / t * 1e3:12.0f} pairs/s   iters mean {out['iterations'].float().mean().item():.1f}")

import numpy as np, torch
from oracle import featx_ref, pipeline_ref, oracle as orc
from sonar_slam_b200 import _lib, ops, pipeline, synth
d = synth.make_trajectory_frames(6, seed=2)
frames, poses = d["frames"].numpy(), d["poses_odom"]
geo = featx_ref.Geometry(30.0 / 512, 512, d["bearings"])
ctx = ops.context(0)
maps = _lib.Maps(ctx, geo.map_x, geo.map_y, 512, 512, geo.width, geo.height)
fe = pipeline.FrontEnd(ctx, maps, max_frames=8, min_points=30)
res = fe.run_host(frames, poses)
clouds, want = pipeline_ref.run(frames, poses, geo, min_points=30)
print("gpu", res["npoints"].tolist(), "oracle", [len(c) for c in clouds])
# stage by stage
img = torch.from_numpy(frames).cuda()
det = ops.cfar(img, "SOCA", 20, 5, 2.749063720096473, gate=65, want_bits=True)
cp = ops.cart_points(maps, bits=det["bits"], capacity=4096)
for i in range(6):
    mask = orc.cfar_u8("SOCA", frames[i], 20, 5, 0, 2.749063720096473, 65)
    print(i, "mask eq", np.array_equal(det["mask"][i].cpu().numpy(), mask))
    locs, pts = featx_ref.cart_points(mask, geo)
    k = int(cp["count"][i])
    print("   cart", k, len(locs), np.array_equal(cp["ij"][i,:k].cpu().numpy(), locs))
    p32 = pts.astype(np.float32)
    ds, di = orc.downsample(p32, 0.5)
    from sonar_slam_b200.bruce_slam import pcl
    g = pcl.downsample(p32, 0.5)
    print("   ds", len(ds), len(g), np.array_equal(ds, g) if len(ds)==len(g) else None)
    ro, _ = orc.remove_outlier(ds, 1.0, 5)
    g2 = pcl.remove_outlier(ds, 1.0, 5)
    print("   ro", len(ro), len(g2))

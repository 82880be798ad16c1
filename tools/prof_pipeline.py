"""Front-end launch sequence for ncu (config 4, 1024 frames)."""
import sys
import torch
from sonar_slam_b200 import _lib, ops, pipeline, synth
from sonar_slam_b200.bruce_slam.feature_extraction import FeatureExtraction
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.cuda.set_device(0)
d = synth.make_trajectory_frames(n, seed=0, device="cuda")
fx = FeatureExtraction()
fx.generate_map_xy(synth.Ping(0, None, 30.0 / 512, 512, d["bearings"]))
ctx = ops.context(0)
maps = _lib.Maps(ctx, fx.map_x, fx.map_y, 512, 512, fx.width, fx.height)
fe = pipeline.FrontEnd(ctx, maps, max_frames=n, icp=_lib.IcpParams(smooth_length=0, max_iterations=20,
                                                              minimizer=int(sys.argv[2]) if len(sys.argv) > 2 else 0))
for _ in range(3):
    fe.run_dev(d["frames"].data_ptr(), d["poses_odom"], n)
torch.cuda.synchronize()

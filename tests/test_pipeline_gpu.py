"""GPU parity of the batched front end (one C-ABI call, host buffers) against the chained CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import featx_ref, oracle as orc, pipeline_ref
from sonar_slam_b200 import _lib, pipeline, synth

pytestmark = pytest.mark.gpu


def _pose(T):
    return np.array([T[0, 2], T[1, 2], np.arctan2(T[1, 0], T[0, 0])], np.float64)


@pytest.mark.parametrize("mode", ["checkers", "fixed20"])
def test_frontend_host_call_equals_oracle_chain(gpu_ctx, mode):
    n = 14
    d = synth.make_trajectory_frames(n, seed=3)
    frames, poses = d["frames"].numpy(), d["poses_odom"]
    geo = featx_ref.Geometry(30.0 / 512, 512, d["bearings"])
    maps = _lib.Maps(gpu_ctx, geo.map_x, geo.map_y, 512, 512, geo.width, geo.height)
    kw = dict(smooth_length=0, max_iterations=20) if mode == "fixed20" else {}
    fe = pipeline.FrontEnd(gpu_ctx, maps, max_frames=32, icp=_lib.IcpParams(**kw), min_points=30)
    got = fe.run_host(frames, poses, chunk_frames=5)
    clouds, want = pipeline_ref.run(frames, poses, geo, min_points=30, icp_params=orc.IcpParams(**kw))
    # clouds (bit-exact): read them back from the device
    r = fe.results_dev()
    stride = r["cloud_stride"]
    xy = gpu_ctx.to_host(r["cloud_xy"], (n, stride, 2), np.float32)
    for i, c in enumerate(clouds):
        assert np.array_equal(xy[i, :len(c)], c), i
    assert np.array_equal(got["npoints"], [len(c) for c in clouds])
    n_matched = 0
    for i in range(n):
        assert got["status"][i] == want[i]["status"], (i, got["status"][i], want[i]["status"])
        if want[i]["status"] == 0:
            n_matched += 1
            assert got["iterations"][i] == want[i]["iterations"], i
            assert got["inliers"][i] == want[i]["inliers"], i
            dlt = np.abs(_pose(got["T"][i]) - _pose(want[i]["T"]))
            assert dlt[:2].max() < 1e-3 and dlt[2] < 1e-3, (i, dlt)
        else:
            assert np.allclose(got["T"][i], want[i]["T"])      # failed / skipped: the guess comes back
    assert got["status"][0] == 7 and n_matched >= n - 4
    # device-resident flavour gives the same answers
    dev_frames = torch.from_numpy(frames).cuda()
    fe.run_dev(dev_frames.data_ptr(), poses, n)
    gpu_ctx.sync()
    again = fe.run_host(frames, poses, chunk_frames=64)
    for k in ("status", "iterations", "inliers", "npoints"):
        assert np.array_equal(again[k], got[k])
    assert np.array_equal(again["T"], got["T"])


def test_frontend_both_icp_size_classes(gpu_ctx):
    """The front end serves ICP with two launches by problem size (pipeline.cu: FE_ICP_SMALL_*).  A finer voxel
    grid gives clouds on both sides of the 640-point class boundary; every frame must still equal the oracle."""
    n = 10
    d = synth.make_trajectory_frames(n, seed=5)
    frames, poses = d["frames"].numpy(), d["poses_odom"]
    geo = featx_ref.Geometry(30.0 / 512, 512, d["bearings"])
    maps = _lib.Maps(gpu_ctx, geo.map_x, geo.map_y, 512, 512, geo.width, geo.height)
    prm = dict(smooth_length=0, max_iterations=20)
    fe = pipeline.FrontEnd(gpu_ctx, maps, max_frames=16, icp=_lib.IcpParams(**prm), min_points=30, resolution=0.35,
                           submap_resolution=0.35, cap_source=4096, cap_target=12288)
    got = fe.run_host(frames, poses, chunk_frames=4)
    clouds, want = pipeline_ref.run(frames, poses, geo, min_points=30, submap_resolution=0.35,
                                    icp_params=orc.IcpParams(**prm), resolution=0.35)
    sizes = np.array([len(c) for c in clouds])
    assert np.array_equal(got["npoints"], sizes)
    assert (sizes > 640).any() and (sizes <= 640).any(), sizes  # both launches had work
    for i in range(n):
        assert got["status"][i] == want[i]["status"], (i, got["status"][i], want[i]["status"])
        if want[i]["status"] == 0:
            assert got["iterations"][i] == want[i]["iterations"] and got["inliers"][i] == want[i]["inliers"], i
            dlt = np.abs(_pose(got["T"][i]) - _pose(want[i]["T"]))
            assert dlt[:2].max() < 1e-3 and dlt[2] < 1e-3, (i, dlt)

"""GPU: point-to-plane scan matching (X1: icp.yaml:18-19 `PointToPlaneErrorMinimizer force2D`, normals from a
SurfaceNormalDataPointsFilter{knn} on the reference) against oracle/icp_ref.c (minimizer 1).

Default mode = the oracle's float32 operation order, normals included: status, iterations, inliers and the 3x3 result
are bit-identical.  Float64-accumulation mode (flags bit 1) and the independent float64 numpy arm of
tests/test_oracle_plane.py are compared within the north star's 1e-3 m / 1e-3 rad on well-conditioned scenes."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from sonar_slam_b200 import _lib, ops, synth
from sonar_slam_b200.bruce_slam import pcl

from test_icp_parity_gpu import _gpu, _pose
from test_oracle_plane import plane_arm

pytestmark = pytest.mark.gpu
N = int(os.environ.get("SFE_ICP_PLANE_SWEEP", "64"))


def _problems(n, rng, big=False):
    pairs, guesses = [], []
    for s in range(n):
        if big:
            src, tgt, Tgt = synth.make_icp_pair(7000 + s)
            guesses.append(np.eye(3))
        else:
            ns, nt = int(rng.integers(250, 640)), int(rng.integers(700, 1536))
            src, tgt, Tgt = synth.make_icp_pair(6000 + s, n_source=ns, n_target=nt)
            guesses.append(Tgt @ synth.se2(*rng.normal(0, [0.1, 0.1, 0.01])))
        pairs.append((src, tgt))
    return pairs, guesses


@pytest.mark.parametrize("mode", ["fixed20", "checkers"])
@pytest.mark.parametrize("size", ["pipeline", "config3"])
def test_plane_mode_is_bit_identical_to_the_oracle(gpu_ctx, mode, size):
    kw = dict(minimizer=1, normals_knn=5)
    if mode == "fixed20":
        kw.update(smooth_length=0, max_iterations=20)
    pairs, guesses = _problems(N if size == "pipeline" else max(N // 8, 4), np.random.default_rng(3), size == "config3")
    got = _gpu(pairs, guesses, **kw)
    f64 = _gpu(pairs, guesses, flags=2, **kw)
    dev = []
    for i, ((s, t), g) in enumerate(zip(pairs, guesses)):
        w = orc.icp(s, t, g.astype(np.float32), orc.IcpParams(**kw))
        assert got["status"][i] == w["status"] and got["iterations"][i] == w["iterations"], (size, mode, i)
        assert got["inliers"][i] == w["inliers"], (size, mode, i)
        assert np.array_equal(got["T"][i].view(np.uint32), w["T"].view(np.uint32)), (size, mode, i, got["T"][i], w["T"])
        assert f64["status"][i] == w["status"]
        if w["status"] == 0 and f64["iterations"][i] == w["iterations"]:
            dev.append(np.abs(_pose(f64["T"][i]) - _pose(w["T"])))
    dev = np.array(dev)
    assert len(dev) >= len(pairs) * 0.8
    # the float64-accumulating mode follows the float32 chain on all but the ill-conditioned scenes (a wall seen
    # end-on: the normal equations amplify the last bit of the sums and the two chains pick different matches)
    q = np.percentile(dev, [50, 90, 100], axis=0)
    print("plane", size, mode, "float64 mode vs oracle [x y theta] p50/p90/max:", q.tolist())
    assert q[1, :2].max() < 1e-3 and q[1, 2] < 1e-3, q


def test_plane_mode_against_the_float64_arm(gpu_ctx):
    kw = dict(minimizer=1, normals_knn=5, smooth_length=0, max_iterations=20)
    pairs, guesses = _problems(16, np.random.default_rng(9))
    got = _gpu(pairs, guesses, **kw)
    worst = 0.0
    for i, ((s, t), g) in enumerate(zip(pairs, guesses)):
        assert got["status"][i] == 0
        d = np.abs(_pose(got["T"][i]) - _pose(plane_arm(s, t, g.astype(np.float32), 20)))
        worst = max(worst, d[:2].max())
        assert d[:2].max() < 5e-3 and d[2] < 1e-3, (i, d)
    print("plane mode vs float64 arm: worst translation deviation", worst)


def test_knn_values_and_small_references(gpu_ctx):
    rng = np.random.default_rng(11)
    for knn in (3, 8, 16):
        pairs, guesses = _problems(6, rng)
        kw = dict(minimizer=1, normals_knn=knn)
        got = _gpu(pairs, guesses, **kw)
        for i, ((s, t), g) in enumerate(zip(pairs, guesses)):
            w = orc.icp(s, t, g.astype(np.float32), orc.IcpParams(**kw))
            assert got["status"][i] == w["status"] and got["iterations"][i] == w["iterations"]
            assert np.array_equal(got["T"][i].view(np.uint32), w["T"].view(np.uint32)), (knn, i)
    # a reference with fewer points than knn, a straight wall (rank-deficient normal equations), coincident points
    wall = np.c_[np.linspace(0, 40, 800), np.zeros(800)].astype(np.float32)
    cases = [(wall[::4] + np.float32([0.3, 0.2]), wall), (wall[:3] + np.float32([0.1, 0.1]), wall[:4]),
             (np.ones((5, 2), np.float32), np.ones((9, 2), np.float32))]
    for s, t in cases:
        got = _gpu([(s, t)], [np.eye(3)], minimizer=1)
        w = orc.icp(s, t, None, orc.IcpParams(minimizer=1))
        assert got["status"][0] == w["status"] and got["iterations"][0] == w["iterations"]
        assert np.array_equal(got["T"][0].view(np.uint32), w["T"].view(np.uint32)), (got["T"][0], w["T"])
    with pytest.raises(_lib.SonarFEError, match="knn"):
        _gpu([cases[0]], [np.eye(3)], minimizer=1, normals_knn=40)


def test_mirror_loads_the_point_to_plane_yaml(gpu_ctx, icp_yaml):
    """bruce_slam.pcl.ICP with the shipped chain, the commented-out minimiser switched on and the reference filter
    that gives it normals."""
    text = open(icp_yaml).read().replace("PointToPointErrorMinimizer", "PointToPlaneErrorMinimizer:\n    force2D: 1")
    text = "referenceDataPointsFilters:\n  - SurfaceNormalDataPointsFilter:\n      knn: 5\n      keepNormals: 1\n" + text
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        f.write(text)
    icp = pcl.ICP()
    icp.loadFromYaml(f.name)
    os.unlink(f.name)
    assert icp.params.minimizer == 1
    src, tgt, Tgt = synth.make_icp_pair(42, n_source=500, n_target=1400)
    msg, T = icp.compute(src, tgt, Tgt.astype(np.float32))
    w = orc.icp(src, tgt, Tgt.astype(np.float32), orc.IcpParams(minimizer=1))
    assert msg == w["message"] == "success" and np.array_equal(T.view(np.uint32), w["T"].view(np.uint32))

"""GPU: the drop-in classes used the way the reference's nodes use them."""
import numpy as np
import pytest

from oracle import featx_ref, oracle as orc, pipeline_ref
from sonar_slam_b200 import synth

pytestmark = pytest.mark.gpu

FEATURE_YAML = {"CFAR": {"Ntc": 40, "Ngc": 10, "Pfa": 0.1, "rank": 10, "alg": "SOCA"},
                "filter": {"threshold": 65, "resolution": 0.5, "radius": 1.0, "min_points": 5, "skip": 1},
                "compressed_images": False}


@pytest.mark.parametrize("tag", ["uniform", "oculus"])
def test_feature_extraction_callback(gpu_ctx, tag, golden_dir):
    from sonar_slam_b200.bruce_slam.feature_extraction import FeatureExtraction
    g = np.load(f"{golden_dir}/featx_config1.npz")
    bearings = synth.bearings_uniform(512) if tag == "uniform" else synth.bearings_oculus(512)
    fe = FeatureExtraction()
    fe.init_node(FEATURE_YAML)
    ping = synth.Ping(ping_id=0, image=synth.make_frame(1), range_resolution=30.0 / 512, num_ranges=512, bearings=bearings)
    pts = fe.callback(ping)
    # geometry and Cartesian pixel list are those of the reference's own callback (fixture)
    assert [fe.rows, fe.cols] == list(g[tag + "_rows_cols"])
    assert np.array_equal(np.array([fe.width, fe.height, fe.res]), g[tag + "_width_height_res"])
    assert np.array_equal(fe.locs, g[tag + "_locs"])
    # filtered cloud == oracle chain on the fixture's points
    want = g[tag + "_points"].astype(np.float32)
    want, _ = orc.downsample(want, 0.5)
    want, _ = orc.remove_outlier(want, 1.0, 5)
    assert pts.dtype == np.float32 and np.array_equal(pts, want)
    # a different alg / no filters
    fe.alg, fe.resolution, fe.outlier_filter_min_points = "CA", 0, 1
    pts2 = fe.callback(ping)
    mask = orc.cfar_u8("CA", ping.image, 20, 5, 0, fe.detector.threshold_factor_CA, 65)
    _, w2 = featx_ref.cart_points(mask, featx_ref.Geometry(ping.range_resolution, 512, bearings))
    assert np.array_equal(pts2, w2.astype(np.float32))
    # skipped ping -> the NaN sentinel cloud (feature_extraction.py:201-207)
    fe.skip = 5
    ping.ping_id = 3
    out = fe.callback(ping)
    assert out.shape == (1, 2) and np.isnan(out).all()
    # geometry change rebuilds the maps
    ping2 = synth.Ping(ping_id=5, image=synth.make_frame(2)[:256], range_resolution=0.1, num_ranges=256, bearings=bearings)
    pts3 = fe.callback(ping2)
    geo2 = featx_ref.Geometry(0.1, 256, bearings)
    assert fe.rows == 256 and fe.cols == geo2.cols
    mask = orc.cfar_u8("CA", ping2.image, 20, 5, 0, fe.detector.threshold_factor_CA, 65)
    _, w3 = featx_ref.cart_points(mask, geo2)
    assert np.array_equal(pts3, w3.astype(np.float32))


def _pose3(p):
    return np.array([p.x(), p.y(), p.theta()])


def test_slam_scan_matching_surface(gpu_ctx, tmp_path):
    from sonar_slam_b200.bruce_slam.slam import SLAM, Pose2, transform_points
    cfg = tmp_path / "icp.yaml"
    cfg.write_text("matcher:\n  KDTreeMatcher:\n    knn: 1\n    epsilon: 0\n    maxDist: 10.0\noutlierFilters:\n"
                   "  - MaxDistOutlierFilter:\n      maxDist: 3.0\n  - TrimmedDistOutlierFilter:\n      ratio: 0.8\n"
                   "errorMinimizer:\n  PointToPointErrorMinimizer\ntransformationCheckers:\n"
                   "  - CounterTransformationChecker:\n      maxIterationCount: 40\n"
                   "  - DifferentialTransformationChecker:\n      minDiffRotErr: 0.01\n      minDiffTransErr: 0.1\n"
                   "      smoothLength: 4\ninspector:\n  NullInspector\n")
    slam = SLAM()
    slam.icp.loadFromYaml(str(cfg))
    src, tgt, _ = synth.make_icp_pair(11, n_source=600, n_target=4000)
    guess = Pose2(0.1, -0.05, 0.01)
    msg, pose = slam.compute_icp(src, tgt, guess)
    want = orc.icp(src, tgt, guess.matrix().astype(np.float32))
    assert msg == "success" == want["message"]
    wp = np.array([want["T"][0, 2], want["T"][1, 2], np.arctan2(want["T"][1, 0], want["T"][0, 0])])
    assert np.abs(_pose3(pose) - wp).max() < 1e-3
    # overlap = number of source points with a target point within point_noise (slam.py:389-424)
    ov, idx = slam.get_overlap(src, tgt, source_pose=pose, return_indices=True)
    wi, _ = orc.match(tgt, transform_points(src, pose), 0.5)
    assert ov == int((wi != -1).sum()) and np.array_equal(idx, wi)
    # 30 initial guesses in one batch + MinCovDet (slam.py:325-387)
    rng = np.random.default_rng(0)
    guesses = [Pose2(*(rng.normal(0, [0.2, 0.2, 0.02]))) for _ in range(30)]
    msg, m, cov, samples = slam.compute_icp_with_cov(src, tgt, guesses)
    assert msg == "success" and cov.shape == (3, 3) and len(samples) >= 5
    one = [orc.icp(src, tgt, g.matrix().astype(np.float32)) for g in guesses[:6]]
    for s, w in zip(samples[:6], one):
        assert np.abs(s - np.array([w["T"][0, 2], w["T"][1, 2], np.arctan2(w["T"][1, 0], w["T"][0, 0])])).max() < 1e-3
    assert slam.compute_icp_with_cov(src, tgt, guesses[:3])[0] == "Too few samples for covariance computation"
    # get_points: transform the window into the reference frame, concatenate, voxel down-sample
    class KF:  # the two Keyframe fields get_points reads
        def __init__(self, points, pose):
            self.points, self.pose = points, pose
    slam.keyframes = [KF(tgt[i::3][:800], Pose2(0.3 * i, 0.1 * i, 0.02 * i)) for i in range(3)]
    got = slam.get_points([0, 1, 2], 2)
    parts = [pipeline_ref.transform_points(k.points, slam.keyframes[2].pose.between(k.pose).matrix().astype(np.float32))
             for k in slam.keyframes]
    want_pts, _ = orc.downsample(np.concatenate([transform_points(k.points, slam.keyframes[2].pose.between(k.pose))
                                                  for k in slam.keyframes]).astype(np.float32), 0.5)
    assert np.array_equal(got, want_pts)
    pts_k, keys = slam.get_points([0, 1, 2], 2, return_keys=True)
    assert np.array_equal(pts_k, want_pts) and set(np.unique(keys)) <= {0.0, 1.0, 2.0}


def test_fov_prefilter_equals_reference_lines(gpu_ctx, golden_dir):
    """Row N3: SLAM.select_targets_in_fov (sfe_fov_select_host) against the fixture produced by exec'ing the
    reference's own lines slam.py:876-899, and the device entry point against the numpy restatement on a large
    random cloud (integer/index work: exact)."""
    import torch
    from oracle import fov_ref
    from sonar_slam_b200 import _lib
    from sonar_slam_b200.bruce_slam.slam import SLAM, Pose2
    g = np.load(f"{golden_dir}/fov_select.npz")
    slam = SLAM()
    slam.oculus.max_range, slam.oculus.horizontal_aperture = float(g["max_range"]), float(g["horizontal_aperture"])

    class KF:
        def __init__(self, pose, cov):
            self.pose, self.cov = pose, cov

    slam.keyframes = {int(k): KF(Pose2(*p), c) for k, p, c in zip(g["source_frames"], g["poses"], g["covs"])}
    pts, keys, sel = slam.select_targets_in_fov(g["target_points"], g["target_keys"], [int(k) for k in g["source_frames"]])
    assert np.array_equal(sel, g["sel"])
    assert np.array_equal(pts, g["kept_points"]) and np.array_equal(keys, g["kept_keys"])
    # no source frames -> nothing selected; empty cloud -> empty
    assert slam.select_targets_in_fov(g["target_points"], None, [])[2].sum() == 0
    assert len(slam.select_targets_in_fov(np.zeros((0, 2), np.float32), None, [17])[0]) == 0
    # device flavour, 200 k points, 5 keyframes
    rng = np.random.default_rng(4)
    big = rng.uniform(-60, 60, (200000, 2)).astype(np.float32)
    T = [Pose2(*(rng.uniform(-1, 1, 3) * [20, 20, 3.1])).inverse().matrix().astype(np.float32) for _ in range(5)]
    rb, bb = rng.uniform(25, 40, 5), rng.uniform(0.9, 1.6, 5)
    want = fov_ref.fov_select(big, T, rb, bb)
    rows = np.array([[t[0, 0], t[0, 1], t[1, 0], t[1, 1], t[0, 2], t[1, 2]] for t in T], np.float32)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pd, Td, rd, bd = d(big), d(rows), d(rb), d(bb)
    out = torch.empty(len(big), dtype=torch.uint8, device="cuda")
    ctx = _lib.default_context()
    _lib.check(ctx.lib.sfe_fov_select_dev(ctx.handle, pd.data_ptr(), len(big), Td.data_ptr(), rd.data_ptr(),
                                          bd.data_ptr(), 5, out.data_ptr()), "sfe_fov_select_dev")
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(bool)
    # float32 atan2 of numpy (SVML / libm) and of CUDA may differ by an ulp: only points within 4 ulp of a bound may differ
    diff = np.flatnonzero(got != want)
    assert len(diff) <= 2, len(diff)

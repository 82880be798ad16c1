"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the loop-closure target pre-filter
(bruce_slam/src/bruce_slam/slam.py:876-899, inside SLAM.initialize_nonsequential_scan_matching), with numpy like the
reference.  Pinned: tests/golden/fov_select.npz was produced by tools/make_golden.py by exec'ing the reference's own
source lines 876-899 (read from /root/reference at generation time) on seeded inputs;
tests/test_oracle_fov.py checks this restatement against it."""
import numpy as np


def transform_points(points, T):
    """Keyframe.transform_points (slam_objects.py:178-198) with a ready float32 3x3."""
    return points.dot(T[:2, :2].T) + T[:2, 2]


def fov_select(target_points, inv_T, range_bound, bearing_bound):
    """target_points float32 [n,2]; inv_T [K] float32 3x3 = pose.inverse().matrix().astype(float32);
    range_bound / bearing_bound [K] float64.  Returns the boolean selection of slam.py:879-895."""
    sel = np.zeros(len(target_points), bool)
    for T, rb, bb in zip(inv_T, range_bound, bearing_bound):
        local_points = transform_points(target_points, T)
        ranges = np.linalg.norm(local_points, axis=1)
        bearings = np.arctan2(local_points[:, 1], local_points[:, 0])
        sel |= (ranges < rb) & (abs(bearings) < bb)
    return sel

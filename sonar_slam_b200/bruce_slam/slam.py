"""Scan-matching helpers of the reference's `SLAM` class, ROS- and gtsam-free
(bruce_slam/src/bruce_slam/slam.py: compute_icp :294-323, compute_icp_with_cov :325-387,
get_overlap :389-424, get_points :229-292; Keyframe.transform_points slam_objects.py:178-198).

The pose-graph part of SLAM (ISAM2, PCM, keyframe logic) is out of this library's scope and stays
with the caller; `Pose2` below is a minimal stand-in for gtsam.Pose2 (x, y, theta, matrix,
between, compose) used when gtsam is not importable.
"""
import time as time_pkg

import numpy as np

from . import pcl

try:  # pragma: no cover - gtsam is not installed in the build image
    from gtsam import Pose2
except Exception:  # noqa: BLE001
    class Pose2(object):
        def __init__(self, x=0.0, y=0.0, theta=0.0):
            self._x, self._y, self._t = float(x), float(y), float(theta)

        def x(self):
            return self._x

        def y(self):
            return self._y

        def theta(self):
            return self._t

        def matrix(self):
            c, s = np.cos(self._t), np.sin(self._t)
            return np.array([[c, -s, self._x], [s, c, self._y], [0.0, 0.0, 1.0]])

        def translation(self):
            return np.array([self._x, self._y])

        def inverse(self):
            c, s = np.cos(self._t), np.sin(self._t)
            return Pose2(-(c * self._x + s * self._y), -(-s * self._x + c * self._y), -self._t)

        def compose(self, other):
            c, s = np.cos(self._t), np.sin(self._t)
            return Pose2(self._x + c * other._x - s * other._y, self._y + s * other._x + c * other._y,
                         self._t + other._t)

        def between(self, other):
            return self.inverse().compose(other)

        def __repr__(self):
            return "Pose2({:.6f}, {:.6f}, {:.6f})".format(self._x, self._y, self._t)


def transform_points(points, pose):
    """Keyframe.transform_points (slam_objects.py:178-198)."""
    if len(points) == 0:
        return np.empty_like(points, np.float32)
    T = pose.matrix().astype(np.float32)
    return points.dot(T[:2, :2].T) + T[:2, 2]


class SLAM(object):
    """Only the scan-matching surface of the reference's SLAM object."""

    def __init__(self):
        self.point_resolution = 0.5  # slam.yaml point_resolution
        self.point_noise = 0.5       # slam.py:73
        self.icp_odom_sigmas = np.array([0.1, 0.1, 0.01])
        self.icp = pcl.ICP()
        self.icp_ssm = pcl.ICP()
        self.keyframes = []

    # ---- slam.py:294-323
    def compute_icp(self, source_points, target_points, guess=None):
        guess = Pose2() if guess is None else guess
        source_points = np.array(source_points, np.float32)
        target_points = np.array(target_points, np.float32)
        message, T = self.icp.compute(source_points, target_points, guess.matrix())
        x, y = T[:2, 2]
        theta = np.arctan2(T[1, 0], T[0, 0])
        return message, Pose2(x, y, theta)

    # ---- slam.py:325-387 (the up-to-30 ICP runs go to the GPU as ONE batch instead of a timed loop)
    def compute_icp_with_cov(self, source_points, target_points, guesses):
        from sklearn.covariance import MinCovDet
        source_points = np.array(source_points, np.float32)
        target_points = np.array(target_points, np.float32)
        res = self.icp.compute_batch(source_points, target_points, [g.matrix() for g in guesses])
        sample_transforms = []
        for message, T in zip(res["messages"], res["T"]):
            if message == "success":
                x, y = T[:2, 2]
                sample_transforms.append((x, y, np.arctan2(T[1, 0], T[0, 0])))
        sample_transforms = np.array(sample_transforms)
        if len(sample_transforms) < 5:
            return "Too few samples for covariance computation", None, None, None
        try:
            fcov = MinCovDet(store_precision=False, support_fraction=0.8).fit(sample_transforms)
        except ValueError:
            return "Failed to calculate covariance", None, None, None
        m = Pose2(*fcov.location_)
        cov = fcov.covariance_
        R = m.matrix()[:2, :2]
        cov[:2, :] = R.T.dot(cov[:2, :])
        cov[:, :2] = cov[:, :2].dot(R)
        default_cov = np.diag(self.icp_odom_sigmas) ** 2
        if np.linalg.det(cov) < np.linalg.det(default_cov):
            cov = default_cov
        return "success", m, cov, sample_transforms

    # ---- slam.py:389-424
    def get_overlap(self, source_points, target_points, source_pose=None, target_pose=None, return_indices=False):
        if source_pose:
            source_points = transform_points(source_points, source_pose)
        if target_pose:
            target_points = transform_points(target_points, target_pose)
        indices, dists = pcl.match(target_points, source_points, 1, self.point_noise)
        if return_indices:
            return np.sum(indices != -1), indices
        return np.sum(indices != -1)

    # ---- slam.py:229-292 for (points, pose) keyframe tuples
    def get_points(self, frames=None, ref_frame=None, return_keys=False):
        """self.keyframes: list of objects with .points ([N,2] float32) and .pose (Pose2)."""
        if frames is None:
            frames = range(len(self.keyframes))
        ref_pose = None
        if ref_frame is not None:
            ref_pose = ref_frame if isinstance(ref_frame, Pose2) else self.keyframes[ref_frame].pose
        all_points = [np.zeros((0, 3 if return_keys else 2), np.float32)]
        for key in frames:
            kf = self.keyframes[key]
            transf = ref_pose.between(kf.pose) if ref_pose is not None else kf.pose
            pts = transform_points(kf.points, transf)
            if return_keys:
                pts = np.c_[pts, key * np.ones((len(pts), 1))]
            all_points.append(pts)
        all_points = np.concatenate(all_points)
        if return_keys:
            return pcl.downsample(all_points[:, :2], all_points[:, (2,)], self.point_resolution)
        return pcl.downsample(all_points, self.point_resolution)

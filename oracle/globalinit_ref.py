"""TEST INFRASTRUCTURE -- CPU oracle of the global-initialisation cost function.

Restates SLAM.get_matching_cost_subroutine1 (bruce_slam/src/bruce_slam/slam.py:461-570) with the same
libraries the reference uses (numpy for the arithmetic, the real cv2.getStructuringElement / cv2.dilate for
the dilation), without ROS and gtsam.  Pinned: tests/golden/globalinit.npz was produced by tools/make_golden.py
by running the reference's own function body (imported unmodified from /root/reference) on the same inputs;
tests/test_oracle_globalinit.py checks this restatement against it.  gtsam.Pose2 itself is absent here: the
pose algebra (compose / between, slam.py:551-553) is the textbook SE(2) one in float64 -- "parity unpinned"
for gtsam's internal rounding, which is below float32 resolution of the transform handed to the points.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np


class Pose2(object):
    """Minimal SE(2) stand-in for gtsam.Pose2 (x, y, theta; compose, between, matrix)."""

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

    def inverse(self):
        c, s = np.cos(self._t), np.sin(self._t)
        return Pose2(-(c * self._x + s * self._y), -(-s * self._x + c * self._y), -self._t)

    def compose(self, other):
        c, s = np.cos(self._t), np.sin(self._t)
        return Pose2(self._x + c * other._x - s * other._y, self._y + s * other._x + c * other._y,
                     self._t + other._t)

    def between(self, other):
        return self.inverse().compose(other)


def transform_points(points, pose):
    """Keyframe.transform_points (slam_objects.py:178-198)."""
    T = pose.matrix().astype(np.float32)
    return points.dot(T[:2, :2].T) + T[:2, 2]


def transform_points_explicit(points, pose):
    """The same transform with the float32 rounding spelled out, independent of which BLAS kernel numpy's float32
    `dot` dispatches to on the host: x' = fl32(fma(y, r01, fl32(x * r00))) + tx (and likewise y') -- what numpy's
    float32 [N,2] @ [2,2] gives on FMA hardware and what the device kernel evaluates (sonarfe.h, sfe_costmap_*).
    The fused multiply-add is formed in float64 (the product of two float32 values is exact there)."""
    T = pose.matrix().astype(np.float32)
    p = np.asarray(points, np.float32)
    x, y = p[:, 0], p[:, 1]

    def row(r0, r1, t):
        first = (x * r0).astype(np.float32)
        acc = (y.astype(np.float64) * np.float64(r1) + first.astype(np.float64)).astype(np.float32)
        return (acc + t).astype(np.float32)
    return np.stack([row(T[0, 0], T[0, 1], T[0, 2]), row(T[1, 0], T[1, 1], T[1, 2])], 1)


def target_grid(target_points, point_noise):
    """slam.py:506-530 -> (grid uint8 [rows, cols] 0/255, xmin, ymin, resolution, dilate_hs)."""
    import cv2
    xmin, ymin = np.min(target_points, axis=0) - 2 * point_noise
    xmax, ymax = np.max(target_points, axis=0) + 2 * point_noise
    resolution = point_noise / 10.0
    xs = np.arange(xmin, xmax, resolution)
    ys = np.arange(ymin, ymax, resolution)
    grid = np.zeros((len(ys), len(xs)), np.uint8)
    r = np.int32(np.round((target_points[:, 1] - ymin) / resolution))
    c = np.int32(np.round((target_points[:, 0] - xmin) / resolution))
    r = np.clip(r, 0, grid.shape[0] - 1)
    c = np.clip(c, 0, grid.shape[1] - 1)
    grid[r, c] = 255
    dilate_hs = int(np.ceil(point_noise / resolution))
    k = 2 * dilate_hs + 1
    kernel = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k), (dilate_hs, dilate_hs))
    return cv2.dilate(grid, kernel), xmin, ymin, resolution, dilate_hs


def cost_of_transform(grid, xmin, ymin, resolution, source_points, sample_transform, explicit=False):
    """slam.py:553-567 for one sample_transform (a Pose2).  explicit=True: the float32 dot product with its
    rounding spelled out (transform_points_explicit) instead of numpy's BLAS-dependent `dot`."""
    points = (transform_points_explicit if explicit else transform_points)(source_points, sample_transform)
    r = np.int32(np.round((points[:, 1] - ymin) / resolution))
    c = np.int32(np.round((points[:, 0] - xmin) / resolution))
    inside = (0 <= r) & (r < grid.shape[0]) & (0 <= c) & (c < grid.shape[1])
    return -int(np.sum(grid[r[inside], c[inside]] > 0))


def boundary_points(xmin, ymin, resolution, source_points, sample_transform, ulps=4):
    """Number of transformed points whose cell coordinate lies within `ulps` float32 ulps of a rounding boundary
    (x.5): the only points whose cell can depend on how the float32 dot product was rounded (BLAS kernel choice)."""
    T = sample_transform.matrix().astype(np.float32).astype(np.float64)
    p = source_points.astype(np.float64).dot(T[:2, :2].T) + T[:2, 2]
    n = 0
    flag = np.zeros(len(p), bool)
    for v, vmin in ((p[:, 1], float(ymin)), (p[:, 0], float(xmin))):
        q = (v - vmin) / float(np.float32(resolution))
        frac = np.abs(q - np.floor(q) - 0.5)
        tol = ulps * np.spacing(np.abs(q).astype(np.float32)).astype(np.float64) \
            + ulps * np.spacing(np.abs(v).astype(np.float32)).astype(np.float64) / resolution
        flag |= frac <= tol
    return int(flag.sum())


def matching_cost_subroutine1(source_points, source_pose, target_points, target_pose, point_noise=0.5):
    """(subroutine, pose_samples, grid) with the reference's closure semantics (slam.py:541-570)."""
    grid, xmin, ymin, resolution, _ = target_grid(target_points, point_noise)
    pose_samples = []

    def subroutine(x):
        delta = Pose2(*x)
        sample_source_pose = source_pose.compose(delta)
        sample_transform = target_pose.between(sample_source_pose)
        cost = cost_of_transform(grid, xmin, ymin, resolution, source_points, sample_transform)
        pose_samples.append(np.r_[sample_source_pose.x(), sample_source_pose.y(), sample_source_pose.theta(), cost])
        return cost

    return subroutine, pose_samples, grid

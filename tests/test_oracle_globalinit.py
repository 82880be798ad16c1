"""CPU: the global-initialisation oracle (oracle/globalinit_ref.py) against the fixture produced by running the
reference's SLAM.get_matching_cost_subroutine1 (slam.py:461-570; tools/make_golden.py), and the host-side
structuring-element formula against cv2."""
import os

import numpy as np
import pytest

from oracle import globalinit_ref as gref
from sonar_slam_b200 import _lib

GOLD = os.path.join(os.path.dirname(__file__), "golden", "globalinit.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_grid_matches_reference(gold):
    grid, xmin, ymin, res, hs = gref.target_grid(gold["target"], 0.5)
    assert grid.shape == tuple(gold["grid_shape"])
    assert hs == 10 and res == 0.05
    want = np.unpackbits(gold["grid_packed"])[: grid.size].reshape(grid.shape)
    assert np.array_equal(grid > 0, want.astype(bool))
    assert set(np.unique(grid)) <= {0, 255}


def test_oracle_costs_match_reference(gold):
    sp, tp = gref.Pose2(*gold["source_pose"]), gref.Pose2(*gold["target_pose"])
    sub, samples, _ = gref.matching_cost_subroutine1(gold["source"], sp, gold["target"], tp, 0.5)
    costs = np.array([sub(x) for x in gold["xs"]])
    assert np.array_equal(costs, gold["costs"])
    assert np.allclose(np.array(samples), gold["pose_samples"], rtol=0, atol=1e-12)
    assert costs.min() < -500 and costs.max() > -400  # the candidates discriminate


@pytest.mark.parametrize("hs", list(range(0, 33)) + [40, 64])
def test_ellipse_spans_match_cv2(hs):
    cv2 = pytest.importorskip("cv2")
    k = 2 * hs + 1
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k), (hs, hs))
    lo, hi = _lib.ellipse_spans(hs)
    m = np.zeros((k, k), np.uint8)
    for i in range(k):
        m[i, lo[i]:hi[i]] = 1
    assert np.array_equal(m, se)


def test_boundary_point_counter():
    pts = np.array([[0.025, 0.0], [0.0249, 0.0], [1.0, 1.0]], np.float32)  # first: exactly on x.5 of the 0.05 grid
    n = gref.boundary_points(np.float32(0.0), np.float32(0.0), 0.05, pts, gref.Pose2())
    assert n >= 1 and n <= 2


def test_explicit_float32_transform_equals_reference_costs(gold):
    """The float32 dot product with its rounding spelled out (what the device kernel evaluates and what the exact
    GPU test compares with) reproduces the costs of the reference-run fixture, and equals numpy's own `dot`
    point for point on random clouds -- so "exact against the explicit formula" is "exact against the reference"."""
    sp, tp = gref.Pose2(*gold["source_pose"]), gref.Pose2(*gold["target_pose"])
    grid, xmin, ymin, res, _ = gref.target_grid(gold["target"], 0.5)
    for x, want in zip(gold["xs"], gold["costs"]):
        tr = tp.between(sp.compose(gref.Pose2(*x)))
        assert gref.cost_of_transform(grid, xmin, ymin, res, gold["source"], tr, explicit=True) == want
    rng = np.random.default_rng(3)
    pts = rng.uniform(-40, 40, (20000, 2)).astype(np.float32)
    diff = 0
    for _ in range(20):
        pose = gref.Pose2(*(rng.uniform(-1, 1, 3) * [5.0, 5.0, 3.0]))
        diff += int((gref.transform_points(pts, pose) != gref.transform_points_explicit(pts, pose)).sum())
    assert diff == 0, diff

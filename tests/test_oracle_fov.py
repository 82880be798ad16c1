"""CPU: the field-of-view pre-filter restatement (oracle/fov_ref.py) against the fixture produced by exec'ing the
reference's own lines slam.py:876-899 (tools/make_golden.py --only-fov)."""
import os

import numpy as np

from oracle import fov_ref, globalinit_ref as gref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fov_select.npz")


def bounds_and_transforms(g):
    """the per-keyframe host expressions of slam.py:883-888,891"""
    inv_T, rb, bb = [], [], []
    for pose, cov in zip(g["poses"], g["covs"]):
        translation_std = np.sqrt(np.max(np.linalg.eigvals(cov[:2, :2])))
        rotation_std = np.sqrt(cov[2, 2])
        rb.append(translation_std * 5.0 + float(g["max_range"]))
        bb.append(rotation_std * 5.0 + float(g["horizontal_aperture"]) * 0.5)
        inv_T.append(gref.Pose2(*pose).inverse().matrix().astype(np.float32))
    return inv_T, rb, bb


def test_restatement_equals_reference_lines():
    g = np.load(GOLD)
    inv_T, rb, bb = bounds_and_transforms(g)
    sel = fov_ref.fov_select(g["target_points"], inv_T, rb, bb)
    assert np.array_equal(sel, g["sel"])
    assert np.array_equal(g["target_points"][sel], g["kept_points"])
    assert np.array_equal(g["target_keys"][sel], g["kept_keys"])
    assert 0.2 < sel.mean() < 0.8                                # the fixture discriminates
    one = fov_ref.fov_select(g["target_points"], inv_T[:1], rb[:1], bb[:1])
    assert one.sum() < sel.sum() and not (one & ~sel).any()      # union over the source keyframes

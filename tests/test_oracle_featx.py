"""CPU-only: pin oracle/featx_ref.py against fixtures produced by the reference's own
FeatureExtraction.callback (tools/make_golden.py), and the integer remap model against cv2."""
import hashlib

import numpy as np
import pytest

from oracle import featx_ref, oracle as orc
from sonar_slam_b200 import synth

TAU_SOCA = 2.749063720096473


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("tag", ["uniform", "oculus"])
def test_geometry_and_points_equal_reference_callback(tag, golden_dir):
    g = np.load(f"{golden_dir}/featx_config1.npz")
    bearings = synth.bearings_uniform(512) if tag == "uniform" else synth.bearings_oculus(512)
    geo = featx_ref.Geometry(30.0 / 512, 512, bearings)
    assert [geo.rows, geo.cols] == list(g[tag + "_rows_cols"])
    assert np.array_equal(np.array([geo.width, geo.height, geo.res]), g[tag + "_width_height_res"])
    assert _sha(geo.map_x) == str(g[tag + "_map_x_sha256"])
    assert _sha(geo.map_y) == str(g[tag + "_map_y_sha256"])
    img = synth.make_frame(1)
    mask = orc.cfar_u8("SOCA", img, 20, 5, 0, TAU_SOCA, 65)
    locs, pts = featx_ref.cart_points(mask, geo)
    assert np.array_equal(locs, g[tag + "_locs"])
    assert np.array_equal(pts, g[tag + "_points"])          # float64, bit-exact


@pytest.mark.parametrize("tag", ["uniform", "oculus"])
def test_integer_remap_model_equals_cv2(tag):
    import cv2
    bearings = synth.bearings_uniform(512) if tag == "uniform" else synth.bearings_oculus(512)
    geo = featx_ref.Geometry(30.0 / 512, 512, bearings)
    rng = np.random.default_rng(0)
    for density in (0.005, 0.2, 0.9):
        mask = (rng.random((512, 512)) < density).astype(np.uint8)
        want = cv2.remap(mask, geo.map_x, geo.map_y, cv2.INTER_LINEAR)
        assert np.array_equal(featx_ref.remap_mask_model(mask, geo.map_x, geo.map_y), want)
    img = synth.make_frame(4)
    mask = orc.cfar_u8("SOCA", img, 20, 5, 0, TAU_SOCA, 65)
    want = cv2.remap(mask, geo.map_x, geo.map_y, cv2.INTER_LINEAR)
    assert np.array_equal(featx_ref.remap_mask_model(mask, geo.map_x, geo.map_y), want)


def test_small_odd_geometry_model_equals_cv2():
    import cv2
    geo = featx_ref.Geometry(0.1, 200, np.round(np.linspace(-3000, 3500, 96)).astype(np.int16))
    rng = np.random.default_rng(1)
    mask = (rng.random((200, 96)) < 0.3).astype(np.uint8)
    want = cv2.remap(mask, geo.map_x, geo.map_y, cv2.INTER_LINEAR)
    assert np.array_equal(featx_ref.remap_mask_model(mask, geo.map_x, geo.map_y), want)

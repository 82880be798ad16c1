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


# ---------------------------------------------------------------------------------------------------------------
# The detection-driven kernel only tests the pixels listed under a detection's polar cell.  The lists are pruned
# (featx.cu: build_inverse_lists): a pixel appears under the heaviest of its taps until the remaining taps together
# weigh < 512.  Host-only checks that this loses nothing: the lists come from the library itself
# (sfe_maps_inverse_lists_host, no GPU), the firing rule from the integer model already pinned to cv2.remap above.
def _walk_lists(mask, off, idx, geo):
    """numpy emulation of cart_scatter_kernel: candidates = lists of the lit cells, each tested with the full rule."""
    lit = np.flatnonzero(mask.ravel())
    starts, ends = off[lit], off[lit + 1]
    n = int((ends - starts).sum())
    if n == 0:
        return np.zeros(0, np.int64), 0
    pos = np.repeat(starts - np.r_[0, np.cumsum(ends - starts)[:-1]], ends - starts) + np.arange(n)
    cand = np.unique(idx[pos])
    fire = featx_ref.remap_mask_model(mask, geo.map_x, geo.map_y).ravel() != 0
    return cand[fire[cand]], n


def test_pruned_list_rule_by_enumeration():
    """Every (fx, fy) fraction, every in-image pattern of the four taps, every lit pattern: a firing pixel has a lit
    tap among the selected ones."""
    import itertools

    def select(w):
        w, rest, sel = list(w), sum(w), []
        while rest >= 512:
            best = max(range(4), key=lambda t: (w[t], -t))
            if w[best] == 0:
                break
            sel.append(best)
            rest -= w[best]
            w[best] = 0
        return sel

    sizes = []
    for fx in range(32):
        for fy in range(32):
            base = [(32 - fx) * (32 - fy), fx * (32 - fy), (32 - fx) * fy, fx * fy]
            for x0, x1, y0, y1 in itertools.product([0, 1], repeat=4):
                w = [base[0] * (x0 & y0), base[1] * (x1 & y0), base[2] * (x0 & y1), base[3] * (x1 & y1)]
                if sum(w) < 512:
                    continue  # unreachable pixel: not listed at all
                sel = select(w)
                if x0 & x1 & y0 & y1:
                    sizes.append(len(sel))
                for lit in itertools.product([0, 1], repeat=4):
                    if sum(w[t] for t in range(4) if lit[t]) >= 512:
                        assert any(lit[t] for t in sel), (fx, fy, w, lit, sel)
    assert np.mean(sizes) < 1.5      # vs 3.9 non-zero taps per pixel


@pytest.mark.parametrize("tag", ["uniform", "oculus"])
def test_pruned_inverse_lists_lose_no_firing_pixel(tag):
    from sonar_slam_b200 import _lib
    bearings = synth.bearings_uniform(512) if tag == "uniform" else synth.bearings_oculus(512)
    geo = featx_ref.Geometry(30.0 / 512, 512, bearings)
    off, idx = _lib.inverse_lists(geo.map_x, geo.map_y, 512, 512)
    assert off[0] == 0 and off[-1] == len(idx) and np.all(np.diff(off) >= 0)
    assert idx.min() >= 0 and idx.max() < geo.rows * geo.cols
    reachable = int((featx_ref.remap_mask_model(np.ones((512, 512), np.uint8), geo.map_x, geo.map_y) != 0).sum())
    assert reachable <= len(idx) < 1.6 * reachable          # ~1.4 entries per reachable pixel (3.9 unpruned)
    rng = np.random.default_rng(7)
    masks = [orc.cfar_u8("SOCA", synth.make_frame(s), 20, 5, 0, TAU_SOCA, 65) for s in (1, 4)]
    masks += [(rng.random((512, 512)) < dens).astype(np.uint8) for dens in (0.002, 0.05, 0.5, 0.97)]
    one = np.zeros((512, 512), np.uint8)
    one[300, 200] = 1
    masks += [one, np.zeros((512, 512), np.uint8), np.ones((512, 512), np.uint8)]
    for m in masks:
        got, n_cand = _walk_lists(m, off, idx, geo)
        locs, _ = featx_ref.cart_points(m, geo)            # the real cv2.remap + np.nonzero
        want = locs[:, 0].astype(np.int64) * geo.cols + locs[:, 1]
        assert np.array_equal(got, want), (tag, int(m.sum()), len(got), len(want))


def test_pruned_inverse_lists_small_odd_geometry():
    from sonar_slam_b200 import _lib
    bearings = np.round(np.linspace(-3000, 3000, 37)).astype(np.int16)
    geo = featx_ref.Geometry(0.11, 53, bearings)
    off, idx = _lib.inverse_lists(geo.map_x, geo.map_y, 53, 37)
    rng = np.random.default_rng(3)
    for dens in (0.03, 0.3, 0.8, 1.0):
        m = (rng.random((53, 37)) < dens).astype(np.uint8)
        got, _ = _walk_lists(m, off, idx, geo)
        locs, _ = featx_ref.cart_points(m, geo)
        assert np.array_equal(got, locs[:, 0].astype(np.int64) * geo.cols + locs[:, 1])

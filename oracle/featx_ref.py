"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the polar -> Cartesian feature cloud.

Restates the numeric body of FeatureExtraction in the reference,
bruce_slam/src/bruce_slam/feature_extraction.py:
    generate_map_xy   :134-173   (polar->Cartesian sampling maps; scipy interp1d)
    callback          :223-238   (CFAR mask, `&= img > threshold`, cv2.remap of the MASK
                                  with INTER_LINEAR, np.nonzero, pixel -> metres)
using the REAL cv2.remap and scipy.interpolate.interp1d (both pip packages present on the
dev container and on the GPU box), so the only restated parts are the array expressions.
Pinned by tests/golden/featx_config1.npz, which tools/make_golden.py produced by running
the reference's own FeatureExtraction.callback.

`remap_mask_model` is an independent integer model of what cv2.remap does to a 0/1 uint8
image (fixed-point bilinear, OpenCV imgwarp.cpp); the CUDA kernel implements that model and
tests check model == cv2 == GPU.
"""
import cv2
import numpy as np
from scipy.interpolate import interp1d


class Geometry:
    """What generate_map_xy leaves on the FeatureExtraction object (:142-173)."""

    def __init__(self, range_resolution, num_ranges, bearings_cdeg):
        to_rad = lambda bearing: bearing * np.pi / 18000
        self.res = range_resolution
        self.height = num_ranges * self.res
        self.rows = num_ranges
        self.width = np.sin(to_rad(bearings_cdeg[-1] - bearings_cdeg[0]) / 2) * self.height * 2
        self.cols = int(np.ceil(self.width / self.res))
        bearings = to_rad(np.asarray(bearings_cdeg, dtype=np.float32))
        f_bearings = interp1d(bearings, range(len(bearings)), kind="linear", bounds_error=False, fill_value=-1,
                              assume_sorted=True)
        XX, YY = np.meshgrid(range(self.cols), range(self.rows))
        x = self.res * (self.rows - YY)
        y = self.res * (-self.cols / 2.0 + XX + 0.5)
        b = np.arctan2(y, x) * 1  # REVERSE_Z = 1
        r = np.sqrt(np.square(x) + np.square(y))
        self.map_y = np.asarray(r / self.res, dtype=np.float32)
        self.map_x = np.asarray(f_bearings(b), dtype=np.float32)


def cart_points(mask, geo):
    """mask: uint8 0/1 polar image.  Returns (locs int64 [K,2] (row, col), points float64 [K,2])."""
    cart = cv2.remap(np.ascontiguousarray(mask), geo.map_x, geo.map_y, cv2.INTER_LINEAR)
    locs = np.c_[np.nonzero(cart)]
    x = locs[:, 1] - geo.cols / 2.
    x = (-1 * ((x / float(geo.cols / 2.)) * (geo.width / 2.)))
    y = (-1 * (locs[:, 0] / float(geo.rows)) * geo.height) + geo.height
    return locs, np.column_stack((y, x))


def remap_mask_model(mask, map_x, map_y):
    """Integer model of cv2.remap(mask01, map_x, map_y, INTER_LINEAR) (constant-0 border).

    Coordinates are quantised to 1/32 px with round-half-even (cvRound), the integer part is
    saturated to int16, bilinear weights are (32-fx)(32-fy), fx(32-fy), (32-fx)fy, fx*fy (sum
    1024, i.e. OpenCV's 15-bit table / 32) and the uint8 result is (32*sum_w + 16384) >> 15.
    """
    R, B = mask.shape
    sx = np.rint(map_x.astype(np.float32) * np.float32(32)).astype(np.int64)
    sy = np.rint(map_y.astype(np.float32) * np.float32(32)).astype(np.int64)
    ix = np.clip(sx >> 5, -32768, 32767)
    iy = np.clip(sy >> 5, -32768, 32767)
    fx, fy = sx & 31, sy & 31
    acc = np.zeros(map_x.shape, np.int64)
    for dy, dx, w in ((0, 0, (32 - fx) * (32 - fy)), (0, 1, fx * (32 - fy)), (1, 0, (32 - fx) * fy), (1, 1, fx * fy)):
        yy, xx = iy + dy, ix + dx
        ok = (yy >= 0) & (yy < R) & (xx >= 0) & (xx < B)
        v = np.zeros(map_x.shape, np.int64)
        v[ok] = mask[yy[ok], xx[ok]]
        acc += w * v
    return ((32 * acc + 16384) >> 15).astype(np.uint8)

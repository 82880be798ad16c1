"""`bruce_slam.cfar` -- same eight functions as the pybind11 module built from
bruce_slam/src/bruce_slam/cpp/cfar.cpp:194-204, running on the GPU.

    ca / soca / goca (img, train_hs, guard_hs, tau)          -> uint8[R, B] (Fortran order)
    os               (img, train_hs, guard_hs, k, tau)        -> uint8[R, B]
    ca2 / soca2 / goca2 / os2 (...)                           -> (uint8[R, B], float32[R, B])

Argument conversion follows pybind11's for `const Eigen::MatrixXf &`: any real 2-D
array is accepted and converted to float32 (uint8 images are passed to the device
as uint8, which gives the same result); anything else raises TypeError.  Returned
arrays are freshly allocated and Fortran-ordered like Eigen's column-major results.
"""
import numbers

import numpy as np

from .. import _lib

_CA, _SOCA, _GOCA, _OS = 0, 1, 2, 3


def _image(img):
    a = np.asarray(img)
    if a.ndim != 2 or a.dtype.kind not in "buif":
        raise TypeError("cfar: incompatible function arguments (expected a 2-D real array for `img`)")
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a), 0
    return np.ascontiguousarray(a, dtype=np.float32), 1


def _int(v, name):
    if isinstance(v, (bool, np.bool_)) or not isinstance(v, (numbers.Integral, np.integer)):
        raise TypeError(f"cfar: incompatible function arguments ({name} must be an int, got {type(v).__name__})")
    return int(v)


def _run(alg, img, train_hs, guard_hs, k, tau, want_thr):
    a, dt = _image(img)
    train_hs, guard_hs, k = _int(train_hs, "train_hs"), _int(guard_hs, "guard_hs"), _int(k, "k")
    if not isinstance(tau, (numbers.Real, np.floating, np.integer)):
        raise TypeError("cfar: incompatible function arguments (tau must be a float)")
    R, B = a.shape
    mask = np.empty((R, B), np.uint8)
    thr = np.empty((R, B), np.float32) if want_thr else None
    ctx = _lib.default_context()
    _lib.check(ctx.lib.sfe_cfar_host(ctx.handle, _lib.ptr(a), dt, 1, R, B, alg, train_hs, guard_hs, k, float(tau),
                                     0, 0.0, _lib.ptr(mask), _lib.ptr(thr)), "cfar")
    if want_thr:
        return np.asfortranarray(mask), np.asfortranarray(thr)
    return np.asfortranarray(mask)


def ca(img, train_hs, guard_hs, tau):
    return _run(_CA, img, train_hs, guard_hs, 0, tau, False)


def soca(img, train_hs, guard_hs, tau):
    return _run(_SOCA, img, train_hs, guard_hs, 0, tau, False)


def goca(img, train_hs, guard_hs, tau):
    return _run(_GOCA, img, train_hs, guard_hs, 0, tau, False)


def os(img, train_hs, guard_hs, k, tau):
    return _run(_OS, img, train_hs, guard_hs, k, tau, False)


def ca2(img, train_hs, guard_hs, tau):
    return _run(_CA, img, train_hs, guard_hs, 0, tau, True)


def soca2(img, train_hs, guard_hs, tau):
    return _run(_SOCA, img, train_hs, guard_hs, 0, tau, True)


def goca2(img, train_hs, guard_hs, tau):
    return _run(_GOCA, img, train_hs, guard_hs, 0, tau, True)


def os2(img, train_hs, guard_hs, k, tau):
    return _run(_OS, img, train_hs, guard_hs, k, tau, True)

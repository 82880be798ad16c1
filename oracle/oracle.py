"""TEST INFRASTRUCTURE ONLY -- ctypes doorways onto the CPU oracle libraries.

  oracle/_build/liboracle.so   our C restatement (cfar_ref.c, cloud_ref.c, icp_ref.c)
  oracle/_ref/libcfar_ref.so   the unmodified reference cfar.cpp (see oracle/Makefile)

Return conventions follow the reference's pybind modules (cfar.cpp:194-204): masks
are uint8 0/1 arrays of the image's shape in Fortran order, the "2" variants also
return the float32 threshold image.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ALG = {"CA": 0, "SOCA": 1, "GOCA": 2, "OS": 3}

_port = None
_ref = None


def build(quiet=True):
    """(Re)build the oracle libraries with oracle/Makefile."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _load_port():
    global _port
    if _port is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        _port = ctypes.CDLL(path)
    return _port


def have_reference():
    return os.path.exists(os.path.join(_HERE, "_ref", "libcfar_ref.so"))


def _load_ref():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(os.path.join(_HERE, "_ref", "libcfar_ref.so"))
    return _ref


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _alg(alg):
    return ALG[alg] if isinstance(alg, str) else int(alg)


def cfar(alg, img, train_hs, guard_hs, k, tau, want_thr=False):
    """Our restatement (oracle/cfar_ref.c).  img: [R,B] any real dtype."""
    lib = _load_port()
    img = np.ascontiguousarray(img, np.float32)
    R, B = img.shape
    mask = np.empty((R, B), np.uint8)
    thr = np.empty((R, B), np.float32) if want_thr else None
    rc = lib.orc_cfar(ctypes.c_int(_alg(alg)), _p(img, ctypes.c_float), ctypes.c_long(B), ctypes.c_long(1),
                      R, B, int(train_hs), int(guard_hs), int(k), ctypes.c_double(tau),
                      _p(mask, ctypes.c_uint8), _p(thr, ctypes.c_float) if want_thr else None)
    if rc:
        raise ValueError("orc_cfar: bad arguments")
    return np.asfortranarray(mask), (np.asfortranarray(thr) if want_thr else None)


def cfar_u8(alg, img_u8, train_hs, guard_hs, k, tau, threshold=-1):
    """uint8 frame -> CFAR -> `&= img > threshold` (feature_extraction.py:223-224)."""
    lib = _load_port()
    img_u8 = np.ascontiguousarray(img_u8, np.uint8)
    R, B = img_u8.shape
    mask = np.empty((R, B), np.uint8)
    rc = lib.orc_cfar_u8(_alg(alg), _p(img_u8, ctypes.c_uint8), R, B, int(train_hs), int(guard_hs), int(k),
                         ctypes.c_double(tau), int(threshold), _p(mask, ctypes.c_uint8))
    if rc:
        raise ValueError("orc_cfar_u8: bad arguments")
    return mask


def cfar_reference(alg, img, train_hs, guard_hs, k, tau, want_thr=False):
    """The unmodified reference cfar.cpp (oracle/_ref)."""
    lib = _load_ref()
    img = np.ascontiguousarray(img, np.float32)
    R, B = img.shape
    mask = np.empty((R, B), np.uint8)
    thr = np.empty((R, B), np.float32) if want_thr else None
    rc = lib.ref_cfar(_alg(alg), _p(img, ctypes.c_float), R, B, int(train_hs), int(guard_hs), int(k),
                      ctypes.c_double(tau), _p(mask, ctypes.c_uint8),
                      _p(thr, ctypes.c_float) if want_thr else None)
    if rc:
        raise ValueError("ref_cfar: bad arguments")
    return np.asfortranarray(mask), (np.asfortranarray(thr) if want_thr else None)

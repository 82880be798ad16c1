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


# ------------------------------------------------------------------ point-cloud helpers / ICP
class IcpParams(ctypes.Structure):
    """Mirror of orc_icp_params (oracle/icp_ref.c); defaults = bruce_slam/config/icp.yaml."""
    _fields_ = [("matcher_max_dist", ctypes.c_float), ("outlier_max_dist", ctypes.c_float),
                ("trim_ratio", ctypes.c_float), ("max_iterations", ctypes.c_int),
                ("min_diff_rot", ctypes.c_float), ("min_diff_trans", ctypes.c_float),
                ("smooth_length", ctypes.c_int), ("flags", ctypes.c_int), ("minimizer", ctypes.c_int),
                ("normals_knn", ctypes.c_int)]

    def __init__(self, matcher_max_dist=10.0, outlier_max_dist=3.0, trim_ratio=0.8, max_iterations=40,
                 min_diff_rot=0.01, min_diff_trans=0.1, smooth_length=4, flags=0, minimizer=0, normals_knn=5):
        super().__init__(matcher_max_dist, outlier_max_dist, trim_ratio, max_iterations, min_diff_rot,
                         min_diff_trans, smooth_length, flags, minimizer, normals_knn)


ICP_MESSAGES = {0: "success", 1: "no outlier to filter", 2: "ErrorMnimizer: no point to minimize",
                3: "abs rotation norm not a number", 4: "abs translation norm not a number",
                5: "RigidTransformation: Error, rotation matrix is not orthogonal.",
                6: "reference cloud is empty"}


def _f32(a, cols=2):
    a = np.ascontiguousarray(a, np.float32)
    return a.reshape(-1, cols) if a.size else a.reshape(0, cols)


def match(ref, pts, max_dist, brute=False):
    """pcl.match(ref, pts, 1, max_dist) -> (ids int32 [1,N], squared dists float32 [1,N])."""
    lib = _load_port()
    ref, pts = _f32(ref), _f32(pts)
    ids = np.empty(len(pts), np.int32)
    d = np.empty(len(pts), np.float32)
    fn = lib.orc_match_brute if brute else lib.orc_match
    fn(_p(ref, ctypes.c_float), len(ref), _p(pts, ctypes.c_float), len(pts), ctypes.c_float(max_dist),
       _p(ids, ctypes.c_int32), _p(d, ctypes.c_float))
    return ids[None, :], d[None, :]


def remove_outlier(pts, radius, min_points):
    lib = _load_port()
    pts = np.ascontiguousarray(pts, np.float32)
    n, dim = pts.shape
    keep = np.empty(n, np.uint8)
    lib.orc_remove_outlier(_p(pts, ctypes.c_float), n, dim, ctypes.c_double(radius), int(min_points),
                           _p(keep, ctypes.c_uint8))
    return pts[keep.astype(bool)], keep.astype(bool)


def downsample(pts, resolution):
    """pcl.downsample(pts, resolution) -> (kept points float32 in output order, their indices)."""
    lib = _load_port()
    pts = _f32(pts)
    idx = np.empty(len(pts), np.int32)
    lib.orc_downsample.restype = ctypes.c_int
    m = lib.orc_downsample(_p(pts, ctypes.c_float), len(pts), ctypes.c_float(resolution), _p(idx, ctypes.c_int32))
    idx = idx[:m].copy()
    return pts[idx], idx


def surface_normals(pts, knn=5):
    """SurfaceNormalDataPointsFilter(knn) on a 2-D cloud -> unit normals float32 [N,2] (orc_surface_normals)."""
    lib = _load_port()
    pts = _f32(pts)
    out = np.zeros((len(pts), 2), np.float32)
    lib.orc_surface_normals.restype = None
    lib.orc_surface_normals(_p(pts, ctypes.c_float), len(pts), int(knn), _p(out, ctypes.c_float))
    return out


def icp(src, tgt, guess=None, params=None):
    """pcl.ICP().compute(src, tgt, guess) -> dict(message, T float32 3x3, iterations, inliers, status)."""
    lib = _load_port()
    src, tgt = _f32(src), _f32(tgt)
    g = np.eye(3, dtype=np.float32) if guess is None else np.ascontiguousarray(guess, np.float32)
    prm = params or IcpParams()
    T = np.empty((3, 3), np.float32)
    it, inl = ctypes.c_int(0), ctypes.c_int(0)
    st = lib.orc_icp(_p(src, ctypes.c_float), len(src), _p(tgt, ctypes.c_float), len(tgt), _p(g, ctypes.c_float),
                     ctypes.byref(prm), _p(T, ctypes.c_float), ctypes.byref(it), ctypes.byref(inl))
    return dict(message=ICP_MESSAGES[st], status=st, T=T if st == 0 else g.copy(), iterations=it.value,
                inliers=inl.value)

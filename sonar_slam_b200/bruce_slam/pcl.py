"""`bruce_slam.pcl` -- same names as the pybind11 module built from
bruce_slam/src/bruce_slam/cpp/pcl.cpp:176-214, running on the GPU through libsonarfe.

    remove_outlier(points[N,2|3], radius, min_points)          -> float32[M,2|3]      (pcl.cpp:54)
    downsample(points[N,2], resolution)                         -> float32[M,2]        (pcl.cpp:128)
    downsample(points[N,2], desc[N,d], resolution)              -> (points, desc)      (pcl.cpp:143)
    match(ref[N_t,2], in[N_s,2], knn, max_dist)                 -> (int32[knn,N_s], float32[knn,N_s])  (:161)
    ICP().loadFromYaml(path) / .compute(source, target, guess)  -> (message, float32[3,3])   (:185-212)

Like pybind11 (`const PM::Matrix &` = Eigen float matrix) every array argument is converted to
float32; results are fresh float32 arrays.  Not provided: density_filter (dead code in the
reference: its only call site is commented out, feature_extraction.py:246) and
ICP.getCovariance (never called from Python) -- both raise NotImplementedError.
"""
import os

import numpy as np

from .. import _lib


def _points(a, name, cols=(2, 3)):
    a = np.asarray(a)
    if a.ndim != 2 or a.dtype.kind not in "buif" or (a.shape[1] not in cols and a.shape[0] != 0):
        raise TypeError(f"pcl: incompatible function arguments ({name} must be a real [N,{'|'.join(map(str, cols))}] array)")
    return np.ascontiguousarray(a, dtype=np.float32)


def remove_outlier(mat_in, radius, min_points):
    pts = _points(mat_in, "points")
    n, dim = pts.shape
    if n == 0:
        return pts.copy()
    ctx = _lib.default_context()
    out = np.empty_like(pts)
    idx = np.empty(n, np.int32)
    m = np.zeros(1, np.int32)
    _lib.check(ctx.lib.sfe_remove_outlier_host(ctx.handle, _lib.ptr(pts), n, dim, float(radius), int(min_points),
                                               _lib.ptr(out), _lib.ptr(idx), _lib.ptr(m)), "remove_outlier")
    return out[:int(m[0])].copy()


def downsample(mat_in, *args):
    """downsample(points, resolution) or downsample(points, desc, resolution)."""
    if len(args) == 1:
        desc, resolution = None, args[0]
    elif len(args) == 2:
        desc, resolution = args
    else:
        raise TypeError("pcl.downsample: incompatible function arguments")
    pts = _points(mat_in, "points", cols=(2,))
    if desc is not None:
        desc = np.asarray(desc)
        if desc.ndim != 2 or len(desc) != len(pts):
            raise TypeError("pcl.downsample: descriptors must be [N,d]")
        desc = np.ascontiguousarray(desc, dtype=np.float32)
    n = len(pts)
    if n == 0:  # pcl.cpp:130,145: empty input comes straight back
        return pts.copy() if desc is None else (pts.copy(), desc.copy())
    ctx = _lib.default_context()
    out = np.empty_like(pts)
    idx = np.empty(n, np.int32)
    m = np.zeros(1, np.int32)
    _lib.check(ctx.lib.sfe_downsample_host(ctx.handle, _lib.ptr(pts), n, 2, float(resolution), _lib.ptr(out),
                                           _lib.ptr(idx), _lib.ptr(m)), "downsample")
    k = int(m[0])
    if desc is None:
        return out[:k].copy()
    return out[:k].copy(), desc[idx[:k]].copy()


def density_filter(*args, **kwargs):
    raise NotImplementedError("pcl.density_filter is dead code in the reference (its only call site is commented "
                              "out, feature_extraction.py:246) and is not provided")


def match(mat_ref, mat_in, knn, max_dist):
    if int(knn) != 1:
        raise NotImplementedError("pcl.match: only knn = 1 is implemented (the only value the reference uses, "
                                  "slam.py:418)")
    ref = _points(mat_ref, "ref", cols=(2,))
    pts = _points(mat_in, "in", cols=(2,))
    ids = np.empty((1, len(pts)), np.int32)
    dists = np.empty((1, len(pts)), np.float32)
    if len(pts):
        ctx = _lib.default_context()
        _lib.check(ctx.lib.sfe_match_host(ctx.handle, _lib.ptr(ref), len(ref), _lib.ptr(pts), len(pts),
                                          float(max_dist), _lib.ptr(ids), _lib.ptr(dists)), "match")
    return ids, dists


# ------------------------------------------------------------------------------------------ ICP
_SUPPORTED = ("only the module chain of the shipped bruce_slam/config/icp.yaml is implemented, plus its commented-out "
              "PointToPlaneErrorMinimizer with a SurfaceNormalDataPointsFilter on the reference")


def parse_icp_yaml(text):
    """libpointmatcher ICP YAML (the subset icp.yaml uses) -> _lib.IcpParams."""
    import yaml
    cfg = yaml.safe_load(text) or {}

    def one(node):  # "Name" | {"Name": {params}} -> (name, params)
        if isinstance(node, str):
            return node, {}
        if isinstance(node, dict) and len(node) == 1:
            (name, prm), = node.items()
            return name, (prm or {})
        raise ValueError(f"ICP yaml: cannot parse module {node!r}")

    prm = _lib.IcpParams(outlier_max_dist=-1.0, trim_ratio=-1.0, max_iterations=40, smooth_length=0)
    if cfg.get("readingDataPointsFilters"):
        raise NotImplementedError(f"ICP yaml: readingDataPointsFilters are not supported ({_SUPPORTED})")
    have_normals = False
    for node in cfg.get("referenceDataPointsFilters") or []:
        name, p = one(node)
        # the one reference filter point-to-plane needs: it attaches the "normals" descriptor
        if name != "SurfaceNormalDataPointsFilter" or float(p.get("epsilon", 0)) != 0:
            raise NotImplementedError(f"ICP yaml: reference filter {name} {p} ({_SUPPORTED})")
        if not int(p.get("keepNormals", 1)):
            continue
        prm.normals_knn = int(p.get("knn", 5))
        if not 3 <= prm.normals_knn <= 16:
            raise NotImplementedError(f"ICP yaml: SurfaceNormalDataPointsFilter knn {prm.normals_knn} (3..16)")
        have_normals = True
    name, p = one(cfg.get("matcher", "KDTreeMatcher"))
    if name != "KDTreeMatcher" or int(p.get("knn", 1)) != 1 or float(p.get("epsilon", 0)) != 0:
        raise NotImplementedError(f"ICP yaml: matcher {name} {p} ({_SUPPORTED})")
    prm.matcher_max_dist = float(p.get("maxDist", float("inf")))
    for node in cfg.get("outlierFilters") or []:
        name, p = one(node)
        if name == "MaxDistOutlierFilter":
            prm.outlier_max_dist = float(p.get("maxDist", 1.0))
        elif name == "TrimmedDistOutlierFilter":
            prm.trim_ratio = float(p.get("ratio", 0.85))
        else:
            raise NotImplementedError(f"ICP yaml: outlier filter {name} ({_SUPPORTED})")
    name, p = one(cfg.get("errorMinimizer", "PointToPointErrorMinimizer"))
    if name == "PointToPlaneErrorMinimizer":  # icp.yaml:18-19 (commented out upstream); 2-D clouds: force2D is moot
        if not have_normals:
            # libpointmatcher: DataPoints::getDescriptorViewByName throws InvalidField at the first iteration
            raise ValueError("ICP yaml: PointToPlaneErrorMinimizer needs the reference's normals "
                             "(InvalidField: Cannot find descriptor normals): add a SurfaceNormalDataPointsFilter "
                             "to referenceDataPointsFilters")
        prm.minimizer = 1
    elif name != "PointToPointErrorMinimizer":
        raise NotImplementedError(f"ICP yaml: error minimizer {name} ({_SUPPORTED})")
    prm.max_iterations = 40
    for node in cfg.get("transformationCheckers") or []:
        name, p = one(node)
        if name == "CounterTransformationChecker":
            prm.max_iterations = int(p.get("maxIterationCount", 40))
        elif name == "DifferentialTransformationChecker":
            prm.min_diff_rot = float(p.get("minDiffRotErr", 0.001))
            prm.min_diff_trans = float(p.get("minDiffTransErr", 0.001))
            prm.smooth_length = int(p.get("smoothLength", 3))
        else:
            raise NotImplementedError(f"ICP yaml: transformation checker {name} ({_SUPPORTED})")
    return prm


class ICP(object):
    """libpointmatcher PM::ICP as exposed by pcl.cpp:185-213."""

    def __init__(self):
        self.params = None  # no modules until loadFromYaml, like a default-constructed PM::ICP

    def loadFromYaml(self, filename):
        if not os.path.isfile(filename):
            # pcl.cpp:190-194 falls back to libpointmatcher's setDefault() (point-to-plane with sampled
            # surface normals), which is not implemented; the shipped icp.yaml values are used instead.
            print("Failed to load " + str(filename) + ". Use default configuration.")
            self.params = _lib.IcpParams()
            return
        with open(filename) as f:
            self.params = parse_icp_yaml(f.read())

    def compute(self, source, target, guess):
        res = self.compute_batch(source, target, [guess])
        return res["messages"][0], res["T"][0]

    def compute_batch(self, source, target, guesses):
        """All `guesses` (iterable of 3x3) on the same cloud pair in one launch (N1: the loop of
        SLAM.compute_icp_with_cov, slam.py:346-358).  Returns dict(messages, T[n,3,3], iterations,
        inliers, status)."""
        if self.params is None:
            raise RuntimeError("You must setup a matcher before running ICP")
        src = _points(source, "source", cols=(2,))
        tgt = _points(target, "target", cols=(2,))
        g = np.ascontiguousarray(np.asarray(guesses, dtype=np.float32).reshape(-1, 3, 3))
        n = len(g)
        T = np.empty((n, 3, 3), np.float32)
        iters, inl, st = (np.zeros(n, np.int32) for _ in range(3))
        ctx = _lib.default_context()
        import ctypes
        _lib.check(ctx.lib.sfe_icp_host(ctx.handle, ctypes.byref(self.params), _lib.ptr(src), len(src), _lib.ptr(tgt),
                                        len(tgt), _lib.ptr(g), n, _lib.ptr(T), _lib.ptr(iters), _lib.ptr(inl),
                                        _lib.ptr(st)), "ICP.compute")
        msgs = [ctx.lib.sfe_icp_status_message(int(s)).decode() for s in st]
        for s, m in zip(st, msgs):
            if int(s) in (5, 6):  # not ConvergenceErrors in libpointmatcher: they propagate as exceptions
                raise RuntimeError(m)
        return dict(messages=msgs, T=T, iterations=iters, inliers=inl, status=st)

    def getCovariance(self):
        raise NotImplementedError("ICP.getCovariance is never called by the reference's Python and is not provided")

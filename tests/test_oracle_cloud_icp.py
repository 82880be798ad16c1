"""CPU-only: cross-check the third-party restatements (oracle/cloud_ref.c, oracle/icp_ref.c) by
independent means -- scipy's cKDTree, brute force, numpy SVD.  These parts of the reference live in
libpointmatcher / libnabo / PCL (not vendored, not installed): parity is UNPINNED, the checks here
only establish that the restatement does what its specification says."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import oracle as orc
from sonar_slam_b200 import synth


def test_grid_nn_equals_brute_force_and_kdtree():
    rng = np.random.default_rng(0)
    for n_ref, n_in, spread in [(20000, 2000, 60.0), (300, 500, 5.0), (1, 10, 1.0), (5000, 100, 0.01)]:
        ref = rng.uniform(0, spread, (n_ref, 2)).astype(np.float32)
        pts = rng.uniform(-0.2 * spread, 1.2 * spread, (n_in, 2)).astype(np.float32)
        for md in (0.5, 10.0):
            i0, d0 = orc.match(ref, pts, md, brute=True)
            i1, d1 = orc.match(ref, pts, md)
            assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
            dd, ii = cKDTree(ref.astype(np.float64)).query(pts.astype(np.float64))
            ok = i1[0] >= 0
            assert np.array_equal(ok, d1[0] <= np.float32(md) ** 2)
            close = np.abs(dd[ok] ** 2 - d1[0][ok]) <= 1e-5 * (1 + dd[ok] ** 2)
            assert close.all()
    i, d = orc.match(np.zeros((0, 2), np.float32), pts, 1.0)
    assert (i == -1).all() and np.isinf(d).all()


def test_match_ties_go_to_lowest_index():
    ref = np.array([[1, 0], [0, 1], [-1, 0], [1, 0]], np.float32)
    ids, d = orc.match(ref, np.array([[0, 0], [1, 0]], np.float32), 5.0)
    assert ids.tolist() == [[0, 0]] and d.tolist() == [[1.0, 0.0]]


def test_remove_outlier_counts_neighbours():
    rng = np.random.default_rng(1)
    pts = np.concatenate([rng.normal(0, 0.4, (300, 2)), rng.uniform(-30, 30, (200, 2))]).astype(np.float32)
    kept, keep = orc.remove_outlier(pts, 1.0, 5)
    d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2)
    d2 = (d2[..., 0] + d2[..., 1]).astype(np.float32)          # same float32 op order
    want = (d2 <= np.float32(1.0)).sum(1) >= 6                  # itself + 5 others
    assert np.array_equal(keep, want)
    assert np.array_equal(kept, pts[want])                      # input order preserved
    p3 = np.c_[pts, rng.normal(0, 0.5, len(pts))].astype(np.float32)
    _, keep3 = orc.remove_outlier(p3, 1.0, 5)
    d3 = ((p3[:, None, :] - p3[None, :, :]) ** 2)
    d3 = ((d3[..., 0] + d3[..., 1]) + d3[..., 2]).astype(np.float32)
    assert np.array_equal(keep3, (d3 <= 1.0).sum(1) >= 6)


def test_downsample_quadtree_medoid_properties():
    rng = np.random.default_rng(2)
    pts = np.concatenate([rng.uniform(0, 20, (3000, 2)), [[0, 0], [20, 20]]]).astype(np.float32)
    out, idx = orc.downsample(pts, 0.5)
    assert len(out) == len(idx) and np.array_equal(out, pts[idx]) and len(set(idx.tolist())) == len(idx)
    # leaf size: bounding square 20 m halves to 0.3125 m (<= 0.5): every leaf is a 64 x 64 grid cell
    cell = 20.0 / 64
    key = lambda p: (np.minimum((p[:, 0] / cell).astype(int), 63), np.minimum((p[:, 1] / cell).astype(int), 63))
    kx, ky = key(pts)
    ox, oy = key(out)
    assert len(set(zip(kx, ky))) == len(out)                   # one point per occupied leaf
    assert len(set(zip(ox, oy))) == len(out)
    # the kept point is the medoid of its leaf
    for j in rng.choice(len(out), 40, replace=False):
        members = pts[(kx == ox[j]) & (ky == oy[j])]
        cost = np.sqrt(((members[:, None] - members[None]) ** 2).sum(-1)).sum(1)
        assert np.allclose(np.sqrt(((members - out[j]) ** 2).sum(-1)).sum(), cost.min(), rtol=1e-5)
    # single point / duplicates / resolution larger than the cloud
    assert np.array_equal(orc.downsample(pts[:1], 0.5)[0], pts[:1])
    assert len(orc.downsample(np.repeat(pts[:1], 7, 0), 0.5)[0]) == 1
    assert len(orc.downsample(pts, 100.0)[0]) == 1


def _numpy_icp(src, tgt, iters):
    """float64 numpy/scipy restatement of the same pipeline (fixed iteration count)."""
    tree = cKDTree(tgt)
    T = np.eye(3)
    for _ in range(iters):
        p = src @ T[:2, :2].T + T[:2, 2]
        d, j = tree.query(p, distance_upper_bound=10.0)
        fin = np.isfinite(d)
        d2 = d ** 2
        lim = np.sort(d2[fin])[int(np.float32(fin.sum()) * np.float32(0.8))]
        keep = fin & (d2 <= 9.0) & (d2 <= lim)
        a, b = p[keep], tgt[j[keep]]
        ma, mb = a.mean(0), b.mean(0)
        U, _, Vt = np.linalg.svd((b - mb).T @ (a - ma))
        R = U @ Vt
        if np.linalg.det(R) < 0:
            Vt[-1] *= -1
            R = U @ Vt
        dT = np.eye(3)
        dT[:2, :2], dT[:2, 2] = R, mb - R @ ma
        T = dT @ T
    return T, int(keep.sum())


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_icp_restatement_tracks_float64_numpy_version_and_ground_truth(seed):
    src, tgt, Tgt = synth.make_icp_pair(seed)
    r = orc.icp(src, tgt, None, orc.IcpParams(smooth_length=0, max_iterations=20))
    assert r["message"] == "success" and r["iterations"] == 20
    Tn, inl = _numpy_icp(src.astype(np.float64), tgt.astype(np.float64), 20)
    T = r["T"].astype(np.float64)
    assert np.abs(T[:2, 2] - Tn[:2, 2]).max() < 2e-3
    assert abs(np.arctan2(T[1, 0], T[0, 0]) - np.arctan2(Tn[1, 0], Tn[0, 0])) < 5e-4
    assert abs(r["inliers"] - inl) <= 2
    assert np.abs(T[:2, 2] - Tgt[:2, 2]).max() < 0.15          # close to the ground truth after 20 iterations
    # shipped checkers: 4 <= iterations <= 40, stops early on this kind of pair
    r2 = orc.icp(src, tgt)
    assert r2["message"] == "success" and 4 <= r2["iterations"] <= 40


def test_icp_failure_messages_and_guess_passthrough():
    src, tgt, _ = synth.make_icp_pair(5, n_source=200, n_target=500)
    g = synth.se2(100.0, 100.0, 0.3).astype(np.float32)         # far away: nothing within 10 m
    r = orc.icp(src, tgt, g)
    assert r["message"] == "no outlier to filter" and np.array_equal(r["T"], g)
    r = orc.icp(src, tgt, g, orc.IcpParams(trim_ratio=-1.0))
    assert r["message"] == "ErrorMnimizer: no point to minimize" and np.array_equal(r["T"], g)
    bad = np.eye(3, dtype=np.float32)
    bad[0, 0] = 1.2
    assert orc.icp(src, tgt, bad)["status"] == 5
    assert orc.icp(np.zeros((0, 2), np.float32), tgt)["message"] == "no outlier to filter"


# ------------------------------------------------------------------ downsample: a second, literal implementation
def _py_quadtree_medoid(pts, resolution):
    """Independent restatement of libpointmatcher's OctreeGridDataPointsFilter (samplingMethod 3) as the SURVEY
    describes it, written recursively over Python lists with numpy float32 scalars: bounding square, split while
    2*radius > resolution and more than one point, children by (x > cx) | (y > cy) << 1 visited in index order,
    medoid = first minimum of the sequential float32 sums of Euclidean distances.  Shares no code with
    oracle/cloud_ref.c (which partitions index ranges in place)."""
    f = np.float32
    pts = np.asarray(pts, np.float32)
    if len(pts) == 0:
        return []
    mn, mx = pts.min(0), pts.max(0)
    rad = mx - mn                                   # float32
    cx, cy = mn[0] + rad[0] * f(0.5), mn[1] + rad[1] * f(0.5)
    radius = (rad[0] if rad[0] >= rad[1] else rad[1]) * f(0.5)
    out = []

    def leaf(members):
        p = pts[members]
        dx = p[:, None, 0] - p[None, :, 0]
        dy = p[:, None, 1] - p[None, :, 1]
        d = np.sqrt(dx * dx + dy * dy)              # float32, one rounding per operation
        acc = np.zeros(len(members), np.float32)
        for b in range(len(members)):               # sequential accumulation in member order
            acc = acc + d[:, b]
        out.append(members[int(np.argmin(acc))])    # argmin: first minimum

    def build(members, cx, cy, radius):
        if not members:
            return
        if float(radius) * 2.0 <= float(f(resolution)) or len(members) <= 1:
            leaf(members)
            return
        kids = [[], [], [], []]
        for i in members:
            kids[int(pts[i, 0] > cx) | (int(pts[i, 1] > cy) << 1)].append(i)
        hr = radius * f(0.5)
        for k in range(4):
            build(kids[k], cx + (hr if k & 1 else -hr), cy + (hr if k & 2 else -hr), hr)

    build(list(range(len(pts))), cx, cy, radius)
    return out


def _sonar_like_cloud(rng, n):
    """Wall samples on a pixel lattice (many exact coordinate ties, like the Cartesian pixel centres of the feature
    node) plus scattered outliers."""
    t = rng.uniform(0, 1, n)
    a = np.c_[3 + 20 * t, -8 + 10 * t + rng.normal(0, 0.2, n)]
    b = np.c_[rng.uniform(0, 30, n // 4), rng.uniform(-27, 27, n // 4)]
    p = np.concatenate([a, b])
    return (np.round(p / 0.0586) * 0.0586).astype(np.float32)


@pytest.mark.parametrize("seed", range(6))
def test_downsample_equals_a_literal_recursive_quadtree(seed):
    rng = np.random.default_rng(100 + seed)
    clouds = [_sonar_like_cloud(rng, 400), rng.uniform(-5, 5, (300, 2)).astype(np.float32),
              np.c_[np.linspace(0, 9, 200), np.zeros(200)].astype(np.float32),            # a line: zero extent in y
              np.repeat(rng.uniform(0, 3, (20, 2)), 5, 0).astype(np.float32)]              # coincident points
    for pts in clouds:
        for res in (0.5, 0.1, 2.0):
            out, idx = orc.downsample(pts, res)
            want = _py_quadtree_medoid(pts, res)
            assert idx.tolist() == want, (seed, len(pts), res)
            assert np.array_equal(out, pts[want])

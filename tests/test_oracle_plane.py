"""CPU: the oracle's point-to-plane mode (X1: icp.yaml:18-19 `PointToPlaneErrorMinimizer force2D` + a
SurfaceNormalDataPointsFilter on the reference; oracle/icp_ref.c, minimizer 1).  libpointmatcher is not vendored
(parity unpinned), so the oracle is checked against analytic cases and against an independent float64 numpy arm
(scipy cKDTree neighbours, numpy eigh normals, numpy solve)."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import oracle as orc
from sonar_slam_b200 import synth


def f64_normals(pts, knn):
    pts = pts.astype(np.float64)
    _, idx = cKDTree(pts).query(pts, k=knn)
    out = np.zeros_like(pts)
    for i, nb in enumerate(idx):
        q = pts[nb] - pts[nb].mean(0)
        w, v = np.linalg.eigh(q.T @ q / knn)
        out[i] = v[:, 0]
    return out


def plane_arm(src, tgt, guess, iters, knn=5):
    src, tgt = src.astype(np.float64), tgt.astype(np.float64)
    nrm = f64_normals(tgt, knn)
    tree = cKDTree(tgt)
    # libpointmatcher linearises about the origin of the reference-centred frame: do the same
    mean = tgt.mean(0)
    T = np.array(guess, np.float64)
    for _ in range(iters):
        p = src @ T[:2, :2].T + T[:2, 2]
        d, j = tree.query(p, distance_upper_bound=10.0)
        fin = np.isfinite(d)
        d2 = d ** 2
        lim = np.sort(d2[fin])[int(np.float32(fin.sum()) * np.float32(0.8))]
        keep = fin & (d2 <= 9.0) & (d2 <= lim)
        q, r, n = p[keep] - mean, tgt[j[keep]] - mean, nrm[j[keep]]
        cr = q[:, 0] * n[:, 1] - q[:, 1] * n[:, 0]
        F = np.c_[cr, n]
        x = np.linalg.solve(F.T @ F, -F.T @ ((q - r) * n).sum(1))
        c, s = np.cos(x[0]), np.sin(x[0])
        dT = np.array([[c, -s, x[1]], [s, c, x[2]], [0, 0, 1]])
        Tm = np.array([[1, 0, mean[0]], [0, 1, mean[1]], [0, 0, 1.0]])
        T = Tm @ dT @ np.linalg.inv(Tm) @ T
    return T


def _pose(T):
    T = np.asarray(T, np.float64)
    return np.array([T[0, 2], T[1, 2], np.arctan2(T[1, 0], T[0, 0])])


def test_normals_of_lines_and_circle():
    rng = np.random.default_rng(1)
    t = np.sort(rng.random(300))
    ang = 0.7
    line = np.c_[t * 30 * np.cos(ang), t * 30 * np.sin(ang)] + [4.0, -7.0]
    n = orc.surface_normals(line, 5)
    assert np.allclose(np.linalg.norm(n, axis=1), 1, atol=1e-6)
    assert np.abs(n @ [np.cos(ang), np.sin(ang)]).max() < 1e-3          # perpendicular to the line
    phi = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    circ = 25.0 * np.c_[np.cos(phi), np.sin(phi)]
    n = orc.surface_normals(circ, 7)
    assert np.abs(np.abs((n * circ).sum(1)) / 25.0 - 1).max() < 1e-3    # radial
    # all neighbours coincide: C == 0, the normal stays zero (upstream's rank test)
    assert np.array_equal(orc.surface_normals(np.ones((6, 2)), 5), np.zeros((6, 2), np.float32))
    # fewer points than knn: realKnn = n
    n = orc.surface_normals(np.array([[0, 0], [1, 0], [2, 0.0]]), 5)
    assert np.allclose(np.abs(n), [[0, 1]] * 3, atol=1e-6)


def test_normals_against_float64_eigh():
    src, tgt, _ = synth.make_icp_pair(3, n_source=300, n_target=1500)
    got = orc.surface_normals(tgt, 5).astype(np.float64)
    want = f64_normals(tgt, 5)
    cross = np.abs(got[:, 0] * want[:, 1] - got[:, 1] * want[:, 0])     # sine of the angle between them
    # float32 covariance of points ~50 m from the origin: near-isotropic neighbourhoods are ill-conditioned
    assert np.percentile(cross, 95) < 2e-3, np.percentile(cross, [50, 90, 95, 99, 100])


@pytest.mark.parametrize("seed", range(8))
def test_plane_icp_recovers_the_pose_and_matches_the_float64_arm(seed):
    src, tgt, Tgt = synth.make_icp_pair(900 + seed, n_source=400, n_target=1200)
    rng = np.random.default_rng(seed)
    guess = (Tgt @ synth.se2(*rng.normal(0, [0.1, 0.1, 0.01]))).astype(np.float32)
    prm = orc.IcpParams(smooth_length=0, max_iterations=20, minimizer=1, normals_knn=5)
    r = orc.icp(src, tgt, guess, prm)
    assert r["status"] == 0 and r["iterations"] == 20
    arm = plane_arm(src, tgt, guess, 20)
    d = np.abs(_pose(r["T"]) - _pose(arm))
    assert d[:2].max() < 5e-3 and d[2] < 1e-3, (d, _pose(r["T"]), _pose(arm))
    # and both sit near the ground truth (walls sampled with 3 cm noise)
    assert np.abs(_pose(r["T"]) - _pose(Tgt))[:2].max() < 0.1


def test_plane_icp_on_a_single_wall_stays_finite():
    """All normals parallel: translation along the wall is unobservable (A is rank 2)."""
    rng = np.random.default_rng(5)
    tgt = np.c_[np.linspace(0, 40, 800), np.zeros(800)].astype(np.float32)
    src = (tgt[::4] + [0.3, 0.2]).astype(np.float32)
    r = orc.icp(src, tgt, None, orc.IcpParams(minimizer=1))
    assert r["status"] == 0 and np.isfinite(r["T"]).all()
    assert abs(r["T"][1, 2] + 0.2) < 1e-3          # the observable component is solved

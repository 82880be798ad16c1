"""Seeded synthetic inputs for the sonar front end.

The reference ships no data (its README points at a rosbag on Google Drive), so
every test / bench input is generated here:

  make_frame(seed)        one uint8 polar sonar image [num_ranges, num_beams]:
                          Rayleigh speckle plus a few bright range-arcs
  bearings_uniform / bearings_oculus
                          int16 centi-degree beam tables (the OculusPing.bearings
                          field read at feature_extraction.py:146,157)
  Ping                    ROS-free stand-in for the OculusPing fields the hot path
                          reads (feature_extraction.py:142-146,201,217)
  make_walls / make_icp_pair
                          2-D wall scene, (source, target, ground truth) scan pairs
  make_trajectory_frames  frames seen from a moving vehicle (pipeline replay)
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Ping:
    """Fields of sonar_oculus/OculusPing that the feature extractor reads."""
    ping_id: int
    image: np.ndarray            # uint8 [num_ranges, num_beams], row 0 = nearest range
    range_resolution: float      # metres per range bin
    num_ranges: int
    bearings: np.ndarray         # int16 centi-degrees, ascending, one per beam
    stamp: float = 0.0


def bearings_uniform(num_beams=512, half_aperture_cdeg=6500):
    return np.round(np.linspace(-half_aperture_cdeg, half_aperture_cdeg, num_beams)).astype(np.int16)


def bearings_oculus(num_beams=512, half_aperture_deg=65.0):
    """Non-uniform table: beams equally spaced in sin(bearing), like the real head."""
    s = np.sin(np.deg2rad(half_aperture_deg))
    b = np.rad2deg(np.arcsin(np.linspace(-s, s, num_beams))) * 100.0
    return np.round(b).astype(np.int16)


def make_frame(seed, num_ranges=512, num_beams=512, n_arcs=6, sigma=18.0):
    """Config-1 style frame: clip(rayleigh(sigma)) + `n_arcs` arcs, 3 bins thick."""
    rng = np.random.default_rng(seed)
    img = rng.rayleigh(sigma, size=(num_ranges, num_beams))
    for _ in range(n_arcs):
        r0 = int(rng.integers(30, num_ranges - 33))
        w = int(rng.integers(30, 111))
        b0 = int(rng.integers(0, max(1, num_beams - w)))
        amp = rng.uniform(90.0, 200.0)
        img[r0:r0 + 3, b0:b0 + w] += amp
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_frames(seeds, **kw):
    return np.stack([make_frame(int(s), **kw) for s in seeds])


# ------------------------------------------------------------------ 2-D scenes / ICP pairs
def make_walls(rng, n_segments=120, extent=60.0, min_len=8.0, max_len=40.0):
    """Random wall segments (x0, y0, x1, y1) inside [0, extent]^2."""
    p0 = rng.uniform(0.0, extent, size=(n_segments, 2))
    ang = rng.uniform(0.0, 2 * np.pi, size=n_segments)
    ln = rng.uniform(min_len, max_len, size=n_segments)
    p1 = np.clip(p0 + np.c_[np.cos(ang), np.sin(ang)] * ln[:, None], 0.0, extent)
    return np.c_[p0, p1]


def sample_walls(rng, walls, n, sigma):
    seg_len = np.hypot(walls[:, 2] - walls[:, 0], walls[:, 3] - walls[:, 1])
    which = rng.choice(len(walls), size=n, p=seg_len / seg_len.sum())
    t = rng.uniform(0.0, 1.0, size=n)
    pts = walls[which, :2] + (walls[which, 2:] - walls[which, :2]) * t[:, None]
    return pts + rng.normal(0.0, sigma, size=pts.shape)


def grid_thin(points, cell):
    """Keep the first point of every `cell`-sized square (order preserving)."""
    keys = np.floor(points / cell).astype(np.int64)
    keys = keys[:, 0] * 1000003 + keys[:, 1]
    _, first = np.unique(keys, return_index=True)
    return points[np.sort(first)]


def se2(x, y, theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s, x], [s, c, y], [0.0, 0.0, 1.0]])


def make_icp_pair(seed, n_source=2000, n_target=20000, extent=60.0, outlier_frac=0.2,
                  max_t=1.0, max_rot=0.1, sensor_range=30.0, half_aperture_deg=65.0):
    """Config-3 pair.  Returns (source[N_s,2] f32, target[N_t,2] f32, T_gt 3x3 f64).

    target: wall samples (sigma 0.03 m), thinned on a 0.1 m grid, in a frame centred on
            the sensor pose;
    source: wall samples inside the sonar wedge (sigma 0.05 m) with `outlier_frac`
            uniform outliers, expressed in a frame displaced by T_gt^-1, so that
            T_gt maps source onto target (what ICP should recover from identity).
    """
    rng = np.random.default_rng(seed)
    n_seg = max(6, int(48 * n_target / 20000))
    walls = make_walls(rng, n_segments=n_seg, extent=extent, min_len=15.0, max_len=50.0)
    tgt = grid_thin(sample_walls(rng, walls, 4 * n_target, 0.03), 0.1)
    rng.shuffle(tgt)
    tgt = tgt[:n_target]
    # sensor pose in the scene
    pose = se2(rng.uniform(0.3, 0.5) * extent, rng.uniform(0.3, 0.7) * extent, rng.uniform(-0.5, 0.5))
    inv = np.linalg.inv(pose)
    to_sensor = lambda p: p @ inv[:2, :2].T + inv[:2, 2]
    tgt_s = to_sensor(tgt)
    cand = to_sensor(sample_walls(rng, walls, 40 * n_source, 0.05))
    r = np.hypot(cand[:, 0], cand[:, 1])
    b = np.arctan2(cand[:, 1], cand[:, 0])
    cand = cand[(r < sensor_range) & (r > 0.5) & (np.abs(b) < np.deg2rad(half_aperture_deg))]
    n_in = int(round(n_source * (1.0 - outlier_frac)))
    if len(cand) < n_in:  # sparse view: pad with resampled structure points
        cand = np.concatenate([cand, cand[rng.integers(0, max(1, len(cand)), n_in - len(cand))]
                               + rng.normal(0, 0.05, (n_in - len(cand), 2))]) if len(cand) else \
            rng.uniform(-5, 5, size=(n_in, 2))
    src = cand[:n_in]
    rr = rng.uniform(0.5, sensor_range, n_source - n_in)
    bb = rng.uniform(-1, 1, n_source - n_in) * np.deg2rad(half_aperture_deg)
    src = np.concatenate([src, np.c_[rr * np.cos(bb), rr * np.sin(bb)]])
    rng.shuffle(src)
    T_gt = se2(rng.uniform(-max_t, max_t), rng.uniform(-max_t, max_t), rng.uniform(-max_rot, max_rot))
    Ti = np.linalg.inv(T_gt)
    src = src @ Ti[:2, :2].T + Ti[:2, 2]
    return src.astype(np.float32), tgt_s.astype(np.float32), T_gt

"""CPU-only: float32 emulations of index arithmetic the kernels rely on, checked exhaustively against brute force.
These do not run the kernels; they test the *claims* the kernels' shortcuts rest on, with numpy float32 standing
in for the device's round-to-nearest float32 operations (grid.cuh: grid_cell_coord, nn_query_seeded)."""
import numpy as np
import pytest

f32 = np.float32


def _cell(v, o, inv_cell, n):
    """grid.cuh: grid_cell_coord -- clamp(int(floorf((v - o) * inv_cell)))"""
    c = np.floor((v.astype(f32) - f32(o)).astype(f32) * f32(inv_cell)).astype(np.int64)
    return np.clip(c, 0, n - 1)


def _d2(qx, qy, tx, ty):
    """grid.cuh: dist2_rn -- every product and the sum rounded to float32"""
    dx, dy = (qx - tx).astype(f32), (qy - ty).astype(f32)
    return ((dx * dx).astype(f32) + (dy * dy).astype(f32)).astype(f32)


@pytest.mark.parametrize("extent,cell", [(30.0, 1.2), (60.0, 0.45), (5.0, 0.05), (2000.0, 8.0), (400.0, 1.6), (1.0, 0.05),
                                         (60000.0, 4.5), (100000.0, 900.0)])
def test_seeded_search_rectangle_contains_every_point_at_least_as_close_as_the_seed(extent, cell):
    """nn_query_seeded scans only the cells [xa, xb] x [ya, yb] that overlap the disc of radius sqrt(d0) * 1.0001 +
    1e-4 * cell around the query (d0 = float32 distance to the seed).  Claim: every target point whose float32
    distance is <= d0 is binned (by grid_build's own cell arithmetic) inside that rectangle -- so the exact nearest
    neighbour and all its ties are found.  The grid geometry is grid_geometry's: cell >= sqrt(area / n), origin at the
    bounding box minimum, coordinates relative to the cloud mean (|v| <= extent)."""
    rng = np.random.default_rng(int(extent * 7 + cell * 1000))
    n = 4000
    t = rng.uniform(-extent / 2, extent / 2, (n, 2)).astype(f32)
    t[: n // 4] = (np.round(t[: n // 4] / cell) * cell).astype(f32)            # points on cell boundaries
    ox, oy = f32(t[:, 0].min()), f32(t[:, 1].min())
    inv = f32(1.0) / f32(cell)
    nx = int((t[:, 0].max() - ox) * inv) + 1
    ny = int((t[:, 1].max() - oy) * inv) + 1
    tcx, tcy = _cell(t[:, 0], ox, inv, nx), _cell(t[:, 1], oy, inv, ny)
    bad = 0
    for trial in range(300):
        seed = rng.integers(0, n)
        # queries near a target point, near cell boundaries, and far outside the grid
        kind = trial % 3
        if kind == 0:
            q = t[rng.integers(0, n)] + rng.normal(0, 0.3 * cell, 2).astype(f32)
        elif kind == 1:
            q = (np.round(rng.uniform(-extent / 2, extent / 2, 2) / cell) * cell).astype(f32)
        else:
            q = rng.uniform(-extent, extent, 2).astype(f32)
        qx, qy = f32(q[0]), f32(q[1])
        d0 = _d2(qx, qy, t[seed, 0], t[seed, 1])
        rad = f32(f32(np.sqrt(d0, dtype=f32) * f32(1.0001)) + f32(f32(1e-4) * f32(cell)))
        xa, xb = _cell(np.array([qx - rad, qx + rad], f32), ox, inv, nx)
        ya, yb = _cell(np.array([qy - rad, qy + rad], f32), oy, inv, ny)
        d = _d2(qx, qy, t[:, 0], t[:, 1])
        need = d <= d0
        inside = (tcx >= xa) & (tcx <= xb) & (tcy >= ya) & (tcy <= yb)
        bad += int(np.sum(need & ~inside))
        assert inside[seed]
    assert bad == 0


def test_radix_select_digit_order_matches_float_order():
    """block_select_kth orders non-negative float32 distances by their bit patterns (4 x 8-bit digits)."""
    rng = np.random.default_rng(0)
    v = np.abs(rng.normal(0, 3, 5000)).astype(f32) ** 2
    v[:10] = 0.0
    bits = v.view(np.uint32)
    assert np.array_equal(np.argsort(bits, kind="stable"), np.argsort(v, kind="stable"))
    for kk in (0, 1, 2500, 3999, 4999):
        prefix = 0
        k = kk
        for shift in (24, 16, 8, 0):  # the kernel's passes: histogram of the digit among entries matching the prefix
            sel = (bits >> (shift + 8)) == (prefix >> (shift + 8)) if shift < 24 else np.ones(len(bits), bool)
            hist = np.bincount((bits[sel] >> shift) & 255, minlength=256)
            cum = np.cumsum(hist)
            b = int(np.searchsorted(cum, k, side="right"))
            k -= int(cum[b - 1]) if b else 0
            prefix |= b << shift
        assert np.uint32(prefix).view(f32) == np.sort(v)[kk]

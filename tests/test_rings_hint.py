"""cs_rings_hint (host-only): the number of rings the library launches for a scan must cover the longest ray the reference
would draw — a session batch launches exactly that many rings per session, so a hint that is too small would lose the far end
of long rays.  Checked against the oracle's own ray set-up and clipping (UpdateHoleMap / ClipRay) on random scans."""
import ctypes as C

import numpy as np
import pytest

import slam.net_b200 as sn
from oracle import oracle as orc


def _hint(size, phys, hw, pts):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    return int(sn.lib().cs_rings_hint(size, phys, hw, pts.ctypes.data_as(C.POINTER(C.c_float)), pts.shape[0]))


def _rings_needed(size, phys, hw, pts, pose):
    m = orc.HoleMap(size, phys)
    m.fill(32750)
    _, rays = orc.update_hole_map(m, pts, pose, hw, 50, rays=True)
    need = 0
    for x1, y1, x2, y2, xp, yp in rays.tolist():
        ok, x2c, y2c = orc.clip_ray(size, x2, y2, x1, y1)        # CoreSLAMProcessor.cs:365
        if not ok:
            continue
        ok, y2c, x2c = orc.clip_ray(size, y2c, x2c, y1, x1)      # :366
        if not ok:
            continue
        need = max(need, max(abs(x2c - x1), abs(y2c - y1)) + 1)  # the loop runs x = 0..dxc (:404)
    return need


@pytest.mark.parametrize("size,phys,hw", [(256, 8.0, 0.6), (512, 40.0, 0.6), (1600, 40.0, 0.6), (300, 16.0, 2.5), (2048, 40.96, 0.05)])
def test_hint_covers_the_longest_drawn_ray(size, phys, hw):
    rng = np.random.default_rng(size)
    for trial in range(12):
        n = int(rng.integers(1, 400))
        reach = phys * (0.05, 0.3, 1.0, 3.0)[trial % 4]          # short scans, room-sized, far beyond the map
        ang = rng.uniform(-np.pi, np.pi, n)
        rad = rng.uniform(1e-3, reach, n)
        pts = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1).astype(np.float32)
        pose = np.array([rng.uniform(0.02, 0.98) * phys, rng.uniform(0.02, 0.98) * phys, rng.uniform(-4, 4)], dtype=np.float32)
        need = _rings_needed(size, phys, hw, pts, pose)
        hint = _hint(size, phys, hw, pts)
        assert 1 <= hint <= size
        assert hint >= min(need, size), (trial, need, hint)
        # and it is a hint, not "always everything": a short scan gets a short grid
        longest = float(np.max(np.hypot(pts[:, 0], pts[:, 1]))) * size / phys + 0.5 * hw * size / phys
        assert hint <= min(size, int(longest * 1.001) + 8)


def test_hint_degenerate_inputs():
    pts = np.array([[np.nan, 1.0], [1.0, 1.0]], dtype=np.float32)
    assert _hint(512, 40.0, 0.6, pts) == 512                      # NaN: all rings
    pts = np.array([[1e30, 1.0]], dtype=np.float32)
    assert _hint(512, 40.0, 0.6, pts) == 512                      # r^2 overflows float: all rings
    pts = np.array([[np.inf, 0.0]], dtype=np.float32)
    assert _hint(512, 40.0, 0.6, pts) == 512
    pts = np.zeros((5, 2), dtype=np.float32)
    assert 1 <= _hint(512, 40.0, 0.6, pts) <= 16                  # nothing but the hole half-width
    assert _hint(512, 40.0, 0.6, pts[:0]) >= 1                    # empty scan

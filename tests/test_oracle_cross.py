"""coreslam_oracle.c vs the independent Python transliteration, bit for bit, on randomized inputs."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import transliteration as tr


def _pair(size, meters, rng, random_fill=True):
    a = orc.HoleMap(size, meters)
    b = tr.HoleMapT(size, meters)
    px = rng.integers(0, 65536, size * size, dtype=np.uint16) if random_fill else np.full(size * size, 32750, np.uint16)
    a.pixels[:] = px
    b.Pixels[:] = px
    assert np.float32(a.scale) == b.Scale
    return a, b


@pytest.mark.parametrize("seed", range(4))
def test_distance_random(seed):
    rng = np.random.default_rng(seed)
    a, b = _pair(48, 6.0, rng)
    for _ in range(40):
        n = int(rng.integers(1, 40))
        pts = (rng.normal(0, 2.5, (n, 2))).astype(np.float32)
        pose = np.array([rng.uniform(-1, 7), rng.uniform(-1, 7), rng.uniform(-10, 10)], dtype=np.float32)
        assert orc.distance(a, pts, pose) == tr.calculate_distance(b, pts, pose)


def test_distance_special_values():
    rng = np.random.default_rng(11)
    a, b = _pair(32, 4.0, rng)
    pts = np.array([[np.nan, 0.0], [1e30, 1e30], [-1e30, 3.0], [0.1, -0.1], [np.inf, 1.0], [-0.05, -0.05]], dtype=np.float32)
    for pose in ([2.0, 2.0, 0.0], [0.0, 0.0, 1.0], [np.nan, 0.0, 0.0], [1e38, -1e38, 5.0], [0.01, 0.01, np.inf]):
        assert orc.distance(a, pts, pose) == tr.calculate_distance(b, pts, np.array(pose, dtype=np.float32))


@pytest.mark.parametrize("seed", range(6))
def test_draw_random_rays(seed):
    rng = np.random.default_rng(100 + seed)
    size = 40
    a, b = _pair(size, 10.0, rng)
    for _ in range(150):
        x1, y1 = (int(v) for v in rng.integers(0, size, 2))
        x2, y2 = (int(v) for v in rng.integers(-60, size + 60, 2))
        # hit point somewhere between start and end (or beyond, or equal)
        t = rng.uniform(-0.2, 1.2)
        xp = int(round(x1 + t * (x2 - x1))) + int(rng.integers(-1, 2))
        yp = int(round(y1 + t * (y2 - y1))) + int(rng.integers(-1, 2))
        alpha = int(rng.integers(1, 256))
        na = orc.draw_ray(a, x1, y1, x2, y2, xp, yp, 0, alpha)
        nb = tr.draw_laser_ray(b, x1, y1, x2, y2, xp, yp, 0, alpha)
        assert na == nb
    assert np.array_equal(np.array(a.pixels), b.Pixels)


def test_clip_ray_random():
    rng = np.random.default_rng(5)
    for _ in range(3000):
        size = int(rng.integers(4, 100))
        xyc, yxc = (int(v) for v in rng.integers(-300, 400, 2))
        xy, yx = (int(v) for v in rng.integers(0, size, 2))
        assert orc.clip_ray(size, xyc, yxc, xy, yx) == tr.clip_ray(size, xyc, yxc, xy, yx)


@pytest.mark.parametrize("seed", range(3))
def test_update_hole_map_random(seed):
    rng = np.random.default_rng(200 + seed)
    a, b = _pair(64, 8.0, rng, random_fill=False)
    for k in range(3):
        n = 90
        ang = np.linspace(0, 2 * np.pi, n, endpoint=False) + rng.uniform(0, 0.1)
        rad = rng.uniform(0.05, 6.0, n)
        pts = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1).astype(np.float32)
        pose = np.array([rng.uniform(0.5, 7.5), rng.uniform(0.5, 7.5), rng.uniform(-4, 4)], dtype=np.float32)
        hw = float(rng.choice([0.3, 0.6, 2.0]))
        q = int(rng.choice([1, 50, 200, 255]))
        va = orc.update_hole_map(a, pts, pose, hw, q)
        vb = tr.update_hole_map(b, pts, pose, hw, q)
        assert va == vb
    assert np.array_equal(np.array(a.pixels), b.Pixels)


def test_parallel_search_and_tiebreak():
    rng = np.random.default_rng(31)
    a, b = _pair(64, 8.0, rng)
    n = 30
    pts = rng.normal(0, 1.5, (n, 2)).astype(np.float32)
    sp = np.array([4.0, 4.0, 0.2], dtype=np.float32)
    T, I = 3, 12
    off = rng.normal(0, 0.2, (T * I, 3)).astype(np.float32)
    best, bd, d, bi = orc.parallel_search(a, pts, sp, off, I, T)
    tb, tbd = tr.parallel_monte_carlo_search(b, pts, sp, off, I, T)
    assert bd == tbd
    assert np.array_equal(best, np.array(tb, dtype=np.float32))
    # flat order: index 0 = searchPose, then 1 + t*I + i; winner = first minimum
    assert d[0] == tr.calculate_distance(b, pts, sp)
    assert bi == int(np.argmin(d)) and bd == int(d.min())
    # ties: constant map -> every in-bounds candidate ties, searchPose (index 0) must win
    a.fill(777)
    best, bd, d, bi = orc.parallel_search(a, pts * 0.1, sp, off * 0.01, I, T)
    assert bi == 0 and np.array_equal(best, sp)
    # nothing in bounds anywhere -> int.MaxValue and searchPose
    far = np.array([1e6, 1e6, 0.0], dtype=np.float32)
    best, bd, d, bi = orc.parallel_search(a, pts, far, off, I, T)
    assert bd == 2147483647 and bi == 0 and np.array_equal(best, far)


def test_processor_replay_small():
    rng = np.random.default_rng(77)
    T, I = 2, 10
    p = orc.Processor(8.0, 64, [4.0, 4.0, 0.0], 0.1, 0.1, I, T)
    q = tr.ProcessorT(8.0, 64, [4.0, 4.0, 0.0], 0.1, 0.1, I, T)
    p.hole_width = 0.8
    q.HoleWidth = np.float32(0.8)
    for k in range(9):
        n = 60
        ang = np.linspace(-np.pi, np.pi, n, endpoint=False)
        rad = 2.5 + 0.3 * np.sin(3 * ang) + rng.uniform(-0.02, 0.02, n)
        pts = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1).astype(np.float32)
        odo = np.array([4.0 + 0.02 * k, 4.0 - 0.01 * k, 0.03 * k], dtype=np.float32)
        off = np.concatenate([rng.normal(0, 0.1, (T * I, 2)), rng.normal(0, 0.1, (T * I, 1))], axis=1).astype(np.float32)
        p.update(pts, odo, off)
        q.Update(pts, odo, off)
        assert np.array_equal(p.pose, np.array(q.Pose, dtype=np.float32)), k
    assert p.scan_count == 5
    assert np.array_equal(np.array(p.map.pixels), q.HoleMap.Pixels)


def test_worker_pool_matches_serial():
    rng = np.random.default_rng(3)
    a, _ = _pair(128, 16.0, rng)
    pts = rng.normal(0, 3.0, (200, 2)).astype(np.float32)
    sp = np.array([8.0, 8.0, -0.4], dtype=np.float32)
    T, I = 4, 50
    off = rng.normal(0, 0.3, (T * I, 3)).astype(np.float32)
    w = orc.Worker(T)
    for _ in range(3):
        best, bd = w.parallel_search(a, pts, sp, off, I)
        sbest, sbd, _, _ = orc.parallel_search(a, pts, sp, off, I, T)
        assert bd == sbd and np.array_equal(best, sbest)
    w.close()

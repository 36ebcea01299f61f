"""ObstacleMap half of Update (CoreSLAMProcessor.cs:456-490, 540-593): hand-derived known answers (KAT-F..H),
the C oracle against the independent Python transliteration, and the closed form of the modified Bresenham
walk that the CUDA kernel evaluates per (ray, step)."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import transliteration as tr


def _fresh(size=8, meters=8.0, fill=-5):
    a = orc.ObstacleMap(size, meters)
    b = tr.ObstacleMapT(size, meters)
    a.fill(fill)
    b.Pixels[:] = fill
    return a, b


def test_kat_f_shallow_ray_hits():
    """(1,1)->(6,3): dx=5, dy=2, err=2; traced by hand from :456-490: cells (1,1)(2,1)(3,2)(4,2)(5,3) are
    marked no-hit, (6,3) takes the hit."""
    a, b = _fresh()
    nh = np.zeros((8, 8), dtype=bool)
    assert orc.draw_ray_obstacle(a, 1, 1, 6, 3) == 6
    assert tr.draw_laser_ray_on_obstacle_map(b, nh, 1, 1, 6, 3, 10) == 6
    want_nohit = {(1, 1), (2, 1), (3, 2), (4, 2), (5, 3)}
    assert {(int(x), int(y)) for y, x in zip(*np.nonzero(a.no_hit))} == want_nohit
    assert {(int(x), int(y)) for y, x in zip(*np.nonzero(nh))} == want_nohit
    exp = np.full((8, 8), -5, np.int8)
    exp[3, 6] = -4
    assert np.array_equal(a.pixels, exp) and np.array_equal(b.Pixels, exp)


def test_kat_g_steep_ray_leaves_the_map():
    """(2,5)->(0,-3): dx=2, dy=8, err=-4: (2,5)(2,4)(2,3)(1,2)(1,1)(1,0), then y=-1 leaves the map: no hit."""
    a, b = _fresh()
    nh = np.zeros((8, 8), dtype=bool)
    assert orc.draw_ray_obstacle(a, 2, 5, 0, -3) == 6
    assert tr.draw_laser_ray_on_obstacle_map(b, nh, 2, 5, 0, -3, 10) == 6
    want = {(2, 5), (2, 4), (2, 3), (1, 2), (1, 1), (1, 0)}
    assert {(int(x), int(y)) for y, x in zip(*np.nonzero(a.no_hit))} == want
    assert {(int(x), int(y)) for y, x in zip(*np.nonzero(nh))} == want
    assert (np.asarray(a.pixels) == -5).all()


def test_kat_h_update_saturation_and_decay():
    """Scale 1 (8 px over 8 m), pose (1,1,0) -> px=py=1.5, x1=y1=1.  Points (5,0) and (5,0) again and (3,0):
    rays end in (6,1), (6,1), (4,1).  Cell (6,1) starts at 9: two hits saturate at MaxObstacleHits=10 (:474).
    Cell (4,1) starts at 10: its own hit leaves 10, but the two longer rays mark it no-hit, so the sweep (:576-592)
    takes it to 9.  Cells (1..3,1),(5,1) are no-hit: -5 -> -4, a 0 stays 0, a +3 goes to +2."""
    a, b = _fresh()
    for m in (a.pixels, b.Pixels):
        m[1, 6] = 9
        m[1, 4] = 10
        m[1, 2] = 0
        m[1, 3] = 3
    pts = np.array([[5, 0], [5, 0], [3, 0]], dtype=np.float32)
    pose = np.array([1, 1, 0], dtype=np.float32)
    ta = orc.update_obstacle_map(a, pts, pose, 10)
    tb = tr.update_obstacle_map(b, pts, pose, 10)
    assert ta == tb == 6 + 6 + 4
    exp = np.full((8, 8), -5, np.int8)
    exp[1, 1:7] = [-4, 0, 2, 9, -4, 10]
    assert np.array_equal(a.pixels, exp), a.pixels[1]
    assert np.array_equal(b.Pixels, exp)


def test_robot_off_map_changes_nothing():
    a, b = _fresh()
    pts = np.array([[1, 0], [0, 1]], dtype=np.float32)
    for pose in ([-1.6, 2.0, 0.0], [2.0, 8.2, 1.0], [np.nan, 1.0, 0.0]):
        assert orc.update_obstacle_map(a, pts, np.array(pose, np.float32), 10) == 0
        assert tr.update_obstacle_map(b, pts, np.array(pose, np.float32), 10) == 0
    assert (np.asarray(a.pixels) == -5).all() and (b.Pixels == -5).all()


@pytest.mark.parametrize("seed", range(5))
def test_draw_random_rays_c_vs_transliteration(seed):
    rng = np.random.default_rng(700 + seed)
    size = 24
    a, b = _fresh(size, 6.0)
    init = rng.integers(-6, 12, (size, size)).astype(np.int8)
    a.pixels[:] = init
    b.Pixels[:] = init
    nh = np.zeros((size, size), dtype=bool)
    for _ in range(200):
        x1, y1 = (int(v) for v in rng.integers(0, size, 2))
        x2, y2 = (int(v) for v in rng.integers(-40, size + 40, 2))
        if rng.random() < 0.1:
            x2, y2 = x1 + int(rng.integers(-1, 2)), y1 + int(rng.integers(-1, 2))
        mh = int(rng.integers(1, 12))
        assert orc.draw_ray_obstacle(a, x1, y1, x2, y2, mh) == tr.draw_laser_ray_on_obstacle_map(b, nh, x1, y1, x2, y2, mh)
    assert np.array_equal(a.pixels, b.Pixels)
    assert np.array_equal(np.asarray(a.no_hit).astype(bool), nh)


@pytest.mark.parametrize("seed", range(4))
def test_update_random_c_vs_transliteration(seed):
    rng = np.random.default_rng(900 + seed)
    size, meters = 32, 8.0
    a, b = _fresh(size, meters)
    for _ in range(6):
        n = int(rng.integers(1, 50))
        pts = rng.normal(0, 3.0, (n, 2)).astype(np.float32)
        pose = np.array([rng.uniform(0, 8), rng.uniform(0, 8), rng.uniform(-7, 7)], dtype=np.float32)
        mh = int(rng.integers(1, 15))
        assert orc.update_obstacle_map(a, pts, pose, mh) == tr.update_obstacle_map(b, pts, pose, mh)
        assert np.array_equal(a.pixels, b.Pixels)


def test_processor_update_with_obstacle_map_c_vs_transliteration():
    rng = np.random.default_rng(77)
    a = orc.Processor(8.0, 32, [4, 4, 0], 0.1, 0.1, 6, 2, obstacle_map_size=16)
    b = tr.ProcessorT(8.0, 32, [4, 4, 0], 0.1, 0.1, 6, 2, obstacle_map_size=16)
    assert (np.asarray(a.obstacle_map.pixels) == -5).all()
    for k in range(8):
        pts = rng.normal(0, 2.0, (20, 2)).astype(np.float32)
        odo = np.array([4 + 0.05 * k, 4, 0.02 * k], dtype=np.float32)
        off = rng.normal(0, 0.1, (12, 3)).astype(np.float32)
        a.update(pts, odo, off)
        b.Update(pts, odo, off)
        assert np.array_equal(a.pose, np.array(b.Pose, dtype=np.float32))
    assert np.array_equal(np.array(a.map.pixels), b.HoleMap.Pixels)
    assert np.array_equal(a.obstacle_map.pixels, b.ObstacleMap.Pixels)
    assert (np.asarray(a.obstacle_map.pixels) != -5).any()


# ---- the closed form the CUDA kernel uses -----------------------------------------------------------------------
def closed_form_cells(size, x1, y1, x2, y2):
    """Cells of DrawLaserRayOnObstacleMap without the loop-carried error term: with h = major/2 the minor
    coordinate after k steps is max(0, ceil((k*minor - h)/major)); the major one is k.  The walk
    ends at the first cell outside the map or at k = major (the hit)."""
    ddx, ddy = x2 - x1, y2 - y1
    dx, dy = abs(ddx), abs(ddy)
    sx, sy = (ddx > 0) - (ddx < 0), (ddy > 0) - (ddy < 0)
    major, minor = (dx, dy) if dx > dy else (dy, dx)
    h = major // 2
    cells, hit = [], None
    for k in range(major + 1):
        num = k * minor - h
        m = 0 if num <= 0 else (num + major - 1) // major
        x, y = (x1 + sx * k, y1 + sy * m) if dx > dy else (x1 + sx * m, y1 + sy * k)
        if not (0 <= x < size and 0 <= y < size):
            break
        if k == major:
            hit = (x, y)
        else:
            cells.append((x, y))
    return cells, hit


@pytest.mark.parametrize("seed", range(4))
def test_closed_form_matches_the_loop(seed):
    rng = np.random.default_rng(1300 + seed)
    size = 40
    for _ in range(400):
        m = orc.ObstacleMap(size, 10.0)
        m.fill(0)
        x1, y1 = (int(v) for v in rng.integers(0, size, 2))
        x2, y2 = (int(v) for v in rng.integers(-70, size + 70, 2))
        if rng.random() < 0.15:
            x2, y2 = x1 + int(rng.integers(-2, 3)), y1 + int(rng.integers(-2, 3))
        if rng.random() < 0.1:
            d = int(rng.integers(-30, 30))
            x2, y2 = x1 + d, y1 + (d if rng.random() < 0.5 else -d)  # exact diagonals: dx == dy
        n = orc.draw_ray_obstacle(m, x1, y1, x2, y2, 10)
        cells, hit = closed_form_cells(size, x1, y1, x2, y2)
        assert n == len(cells) + (hit is not None)
        assert {(int(x), int(y)) for y, x in zip(*np.nonzero(m.no_hit))} == set(cells)
        hits = {(int(x), int(y)) for y, x in zip(*np.nonzero(m.pixels))}
        assert hits == ({hit} if hit is not None else set())

"""ObstacleMap half of Update on the device vs the CPU oracle (CoreSLAMProcessor.cs:456-490, 540-593) through the
C ABI: bit-exact sbyte maps.  Needs a B200 (pytest -m gpu); the oracle is the checker only."""
import numpy as np
import pytest

import slam.net_b200 as sn
from slam.net_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _pair(phys, hole, obst, start=(0, 0, 0), iters=4, threads=2, max_points=0):
    p = sn.Processor(phys, hole, start, 0.1, 0.17, iters, threads, max_points=max_points, obstacle_map_size=obst, seed=3)
    m = orc.ObstacleMap(obst, phys)
    m.fill(-5)
    assert p.obstacle_size == obst and np.float32(p.obstacle_scale) == np.float32(m.scale)
    return p, m


def test_reset_fills_unmapped_value_and_kats():
    p, m = _pair(8.0, 32, 8)
    assert (p.obstacle_map_download() == -5).all()
    # KAT-H of tests/test_oracle_obstacle.py
    init = np.full((8, 8), -5, np.int8)
    init[1, 6], init[1, 4], init[1, 2], init[1, 3] = 9, 10, 0, 3
    p.obstacle_map_upload(init)
    assert np.array_equal(p.obstacle_map_download(), init)
    pts = np.array([[5, 0], [5, 0], [3, 0]], dtype=np.float32)
    p.integrate(pts, [1, 1, 0])
    exp = np.full((8, 8), -5, np.int8)
    exp[1, 1:7] = [-4, 0, 2, 9, -4, 10]
    assert np.array_equal(p.obstacle_map_download(), exp)
    assert p.obstacle_visits() == 16
    # UnmappedObstacleHits takes effect at Reset (:96-98, :170)
    p.set_unmapped_obstacle_hits(-9)
    assert np.array_equal(p.obstacle_map_download(), exp)
    p.reset()
    assert (p.obstacle_map_download() == -9).all() and p.obstacle_visits() == 0
    p.close()


@pytest.mark.parametrize("obst,n_points", [(8, 5), (30, 77), (64, 360), (250, 1024), (1000, 3000), (2048, 1024)])
def test_update_obstacle_map_bit_exact(obst, n_points):
    rng = np.random.default_rng(obst * 31 + n_points)
    phys = 20.0
    p, m = _pair(phys, 64, obst, max_points=n_points)
    init = rng.integers(-7, 12, (obst, obst)).astype(np.int8)
    p.obstacle_map_upload(init)
    m.pixels[:] = init
    total = 0
    for trial in range(6):
        pose = np.array([rng.uniform(0.5, phys - 0.5), rng.uniform(0.5, phys - 0.5), rng.uniform(-7, 7)], dtype=np.float32)
        scale = [3.0, 8.0, 30.0][trial % 3]  # short rays, typical rays, rays that leave the map
        pts = rng.normal(0, scale, (n_points, 2)).astype(np.float32)
        if trial == 3:
            pts[: n_points // 3] = pts[0]        # many rays ending in one cell: saturation
            pts[n_points // 3: n_points // 2] *= 1e-3  # rays that hit in the start cell
        mh = [10, 3, 127, 1, 10, 10][trial]
        p.set_max_obstacle_hits(mh)
        total += orc.update_obstacle_map(m, pts, pose, mh)
        p.integrate(pts, pose)
        got = p.obstacle_map_download()
        assert np.array_equal(got, m.pixels), "trial %d: %d cells differ" % (trial, int((got != m.pixels).sum()))
    assert p.obstacle_visits() == total
    p.close()


def test_obstacle_map_special_points_and_off_map_robot():
    phys, obst = 10.0, 50
    p, m = _pair(phys, 64, obst)
    rng = np.random.default_rng(5)
    init = rng.integers(-5, 11, (obst, obst)).astype(np.int8)
    p.obstacle_map_upload(init)
    m.pixels[:] = init
    # far, tiny, axis-aligned and exactly diagonal rays; non-finite points are excluded (Math.Abs(int.MinValue) throws
    # in the reference for the ray they produce from x1 = 0 only; elsewhere they are ordinary far rays)
    pts = np.array([[1e6, 1e6], [-1e6, 3.0], [0.0, 0.0], [1e-30, -1e-30], [2.0, 0.0], [0.0, -2.0], [1.0, 1.0], [-1.5, 1.5],
                    [3e9, 0.0], [0.0, -3e9], [1e20, 1e20]], dtype=np.float32)
    for pose in ([5, 5, 0], [5, 5, 0.7853982], [0.02, 9.97, 3.0], [9.99, 0.0, -2.0]):
        pose = np.array(pose, dtype=np.float32)
        orc.update_obstacle_map(m, pts, pose, 10)
        p.integrate(pts, pose)
        assert np.array_equal(p.obstacle_map_download(), m.pixels)
    before = p.obstacle_map_download()
    for pose in ([-3.0, 5.0, 0.0], [5.0, 10.6, 1.0]):  # robot off the map: :557-560 returns before any ray
        p.integrate(pts, np.array(pose, dtype=np.float32))
        assert np.array_equal(p.obstacle_map_download(), before)
    p.close()


@pytest.mark.parametrize("mode", ["offsets", "philox"])
def test_update_replay_with_obstacle_map_bit_exact(mode):
    """Whole Update (search + HoleMap + ObstacleMap) against the oracle processor, scan by scan, both maps."""
    n_scans, n_points, size, obst, phys, iters, threads = 14, 360, 512, 128, 40.0, 64, 4
    rp = synth.make_replay(n_scans, n_points, phys)
    p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, max_points=n_points, obstacle_map_size=obst, seed=11)
    o = orc.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, obstacle_map_size=obst)
    total = 0
    for k in range(n_scans):
        if mode == "offsets":
            off = synth.candidate_offsets(2, k, iters * threads, 0.1, 0.17)
            r = p.update(rp.points[k], rp.odometry[k], off)
        else:
            off = sn.philox_offsets(p.seed, k, iters * threads, 0.1, 0.17)
            r = p.update(rp.points[k], rp.odometry[k], None)
        o.update(rp.points[k], rp.odometry[k], off)
        total += o.obstacle_visits
        assert np.array_equal(r.pose, o.pose), k
    assert np.array_equal(p.map_download(), np.array(o.map.pixels))
    got = p.obstacle_map_download()
    assert np.array_equal(got, o.obstacle_map.pixels)
    assert (got > 0).any() and (got == 0).any() and (got == -5).any()  # walls, free space, unmapped
    assert p.obstacle_visits() == total
    p.close()


def test_replay_log_with_obstacle_map_matches_updates():
    n_scans, n_points, size, obst, phys, iters, threads = 12, 256, 256, 64, 40.0, 32, 2
    rp = synth.make_replay(n_scans, n_points, phys)
    a = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, max_points=n_points, obstacle_map_size=obst, seed=4)
    b = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, max_points=n_points, obstacle_map_size=obst, seed=4)
    log = sn.ScanLog(n_scans, n_points, n_offsets=0)
    for k in range(n_scans):
        log.set(k, rp.points[k], rp.odometry[k])
        a.update(rp.points[k], rp.odometry[k], None)
    log.upload()
    b.replay(log, 0, n_scans, want_results=False)
    b.sync()
    assert np.array_equal(a.get_pose(), b.get_pose())
    assert np.array_equal(a.obstacle_map_download(), b.obstacle_map_download())
    assert np.array_equal(a.map_download(), b.map_download())
    a.close(); b.close(); log.close()


def test_obstacle_map_absent_is_an_error_not_a_fallback():
    p = sn.Processor(10.0, 64, [5, 5, 0], 0.1, 0.1, 4, 1)
    with pytest.raises(sn.CoreSlamError):
        p.obstacle_map_fill(0)
    p.close()


def test_reference_shaped_processor_exposes_obstacle_map():
    from slam.net_b200.coreslam import CoreSLAMProcessor, ScanSegment
    n_points, phys = 180, 40.0
    rp = synth.make_replay(3, n_points, phys)
    with CoreSLAMProcessor(phys, 256, 64, rp.odometry[0], 0.1, 0.17, 16, 2) as slam:
        assert slam.ObstacleMap.Size == 64 and slam.MaxObstacleHits == 10 and slam.UnmappedObstacleHits == -5
        o = orc.Processor(phys, 256, rp.odometry[0], 0.1, 0.17, 16, 2, obstacle_map_size=64)
        for k in range(3):
            ang = np.arctan2(rp.points[k][:, 1], rp.points[k][:, 0]).astype(np.float32)
            rad = np.hypot(rp.points[k][:, 0], rp.points[k][:, 1]).astype(np.float32)
            seg = ScanSegment(Rays=np.stack([ang, rad], axis=1), Pose=rp.odometry[k], IsLast=True)
            slam.Update([seg])
            pts = orc.segment_to_cloud(np.stack([ang, rad], axis=1), rp.odometry[k], rp.odometry[k])
            o.update(pts, rp.odometry[k], None)
        assert np.array_equal(slam.ObstacleMap.Pixels, o.obstacle_map.pixels)

"""ScanSegmentsToCloud on the device: cs_segments_to_cloud / cs_update_segments vs the CPU oracle's
restatement of CoreSLAMProcessor.cs:187-207 — bit-exact points, and whole multi-segment Updates (motion-compensated scans:
every segment has its own pose, the last one is the odometry pose, :719) bit-exact in pose, HoleMap and ObstacleMap.
"""
import numpy as np
import pytest

import slam.net_b200 as sn
from slam.net_b200 import _native as N
from slam.net_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _oracle_cloud(segments, odo):
    parts = [orc.segment_to_cloud(seg.rays_array(), np.asarray(seg.Pose, dtype=np.float32), odo) for seg in segments
             if seg.rays_array().shape[0]]
    return np.concatenate(parts, axis=0)


def _split(angles, radii, poses, cuts):
    """One scan cut into segments at ray indices `cuts`, segment s at poses[s]."""
    segs, lo = [], 0
    for s, hi in enumerate(list(cuts) + [len(angles)]):
        segs.append(sn.ScanSegment(Rays=np.stack([angles[lo:hi], radii[lo:hi]], axis=1).astype(np.float32), Pose=poses[s],
                                   IsLast=(hi == len(angles))))
        lo = hi
    return segs


def test_segments_to_cloud_bit_exact():
    rng = np.random.default_rng(5)
    p = sn.Processor(40.0, 256, (0, 0, 0), 0.1, 0.17, 4, 1, max_points=4096)
    for n, cuts in ((1, []), (37, [5, 5, 20]), (1000, [1, 333, 334, 900]), (4096, [1024, 2048, 3072])):
        ang = rng.uniform(-7.0, 7.0, n).astype(np.float32)
        ang[::7] *= 50.0  # a few angles far outside [-pi, pi]: the large-argument reduction of cosf/sinf
        rad = rng.uniform(0.05, 30.0, n).astype(np.float32)
        poses = rng.normal(0, [3.0, 3.0, 2.0], (len(cuts) + 1, 3)).astype(np.float32)
        segs = _split(ang, rad, poses, cuts)
        odo = poses[-1] + np.array([0.01, -0.02, 0.003], dtype=np.float32)
        got = p.segments_to_cloud(segs, odo)
        want = _oracle_cloud(segs, odo)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        # and the host twin of the mirror (same arithmetic through the library's host build of cosf/sinf)
        host = sn.coreslam.scan_segments_to_cloud(segs, odo).Points
        assert np.array_equal(host.view(np.uint32), want.view(np.uint32))
    p.close()


def test_segments_argument_errors():
    p = sn.Processor(10.0, 64, (5, 5, 0), 0.1, 0.1, 8, 1, max_points=16)
    with pytest.raises(sn.CoreSlamError):
        p.update_segments([])  # segments.Last() on an empty list (:719)
    rays = np.array([[0.0, 1.0], [0.1, 1.0]], dtype=np.float32)
    poses = np.array([[5, 5, 0]], dtype=np.float32)
    r = N.Result()
    bad_first = np.array([0, 1], dtype=np.int32)  # does not end at n_rays
    assert N.lib().cs_update_segments(p._h, rays.ctypes.data_as(N._fp), bad_first.ctypes.data_as(N._ip), poses.ctypes.data_as(N._fp),
                                      2, 1, None, N.C.byref(r)) == 1
    with pytest.raises(sn.CoreSlamError):  # more rays than max_points
        p.update_segments([sn.ScanSegment(Rays=np.ones((17, 2), dtype=np.float32), Pose=(5, 5, 0))])
    p.update_segments([sn.ScanSegment(Rays=rays, Pose=(5, 5, 0), IsLast=True)])  # the handle is still usable
    p.close()


@pytest.mark.parametrize("search", [N.FLAG_SEARCH_WARP, N.FLAG_SEARCH_SLAB])
def test_multi_segment_update_replay_bit_exact(search):
    """A moving robot: each 360-degree scan arrives as 4 segments taken at 4 poses along the way."""
    n_scans, P, size, phys, iters, threads = 30, 360, 400, 40.0, 100, 4
    rp = synth.make_replay(n_scans + 1, P, phys, seed=31)
    p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, max_points=P, flags=search, obstacle_map_size=200)
    o = orc.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, obstacle_map_size=200)
    ang = (np.arange(P) * (2 * np.pi / P)).astype(np.float32)
    for k in range(n_scans):
        rad = np.hypot(rp.points[k][:, 0], rp.points[k][:, 1]).astype(np.float32)
        # poses of the 4 segments: from the previous odometry pose to this one (the last segment carries the odometry pose)
        a, b = rp.odometry[max(k - 1, 0)], rp.odometry[k]
        poses = np.stack([a + (b - a) * t for t in (0.25, 0.5, 0.75, 1.0)]).astype(np.float32)
        segs = _split(ang, rad, poses, [90, 180, 270])
        off = synth.candidate_offsets(17, k, iters * threads, 0.1, 0.17)
        res = p.update_segments(segs, off)
        odo = poses[-1]
        o.update(_oracle_cloud(segs, odo), odo, off)
        assert np.array_equal(res.pose, o.pose), k
        if res.searched:
            assert (res.distance, res.index) == (o.last_distance, o.last_index)
    assert np.array_equal(p.map_download(), np.array(o.map.pixels))
    assert np.array_equal(p.obstacle_map_download(), o.obstacle_map.pixels)
    p.close()


def test_update_with_segments_that_carry_no_rays_advances_the_state_like_the_reference():
    """Update(List<ScanSegment>) with non-empty segments whose Rays lists are empty still runs in the
    reference (:719-747): lastOdometryPose / scanCount advance, every distance is int.MaxValue, searchPose wins.  The drop-in
    must do the same (and the next scan's searchPose must agree with the oracle's)."""
    n_scans, P, size, phys, iters, threads = 12, 90, 256, 40.0, 30, 2
    rp = synth.make_replay(n_scans, P, phys, seed=77)
    p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, max_points=P)
    o = orc.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads)
    ang = (np.arange(P) * (2 * np.pi / P)).astype(np.float32)
    empty = np.zeros((0, 2), dtype=np.float32)
    for k in range(n_scans):
        off = synth.candidate_offsets(78, k, iters * threads, 0.1, 0.17)
        odo = rp.odometry[k]
        if k in (2, 7, 8):  # one map-only scan and two searched scans arrive without a single ray
            segs = [sn.ScanSegment(Rays=empty, Pose=odo), sn.ScanSegment(Rays=empty, Pose=odo, IsLast=True)]
            cloud = empty
        else:
            rad = np.hypot(rp.points[k][:, 0], rp.points[k][:, 1]).astype(np.float32)
            n = rad.shape[0]
            segs = [sn.ScanSegment(Rays=np.stack([ang[:n], rad], axis=1), Pose=odo, IsLast=True)]
            cloud = orc.segment_to_cloud(segs[0].rays_array(), np.asarray(odo, dtype=np.float32), np.asarray(odo, dtype=np.float32))
        r = p.update_segments(segs, off)
        o.update(cloud, odo, off)
        assert np.array_equal(r.pose, o.pose), k
        if k >= 5:
            assert (r.distance, r.index) == (o.last_distance, o.last_index), k
            if k in (7, 8):
                assert (r.distance, r.index) == (2147483647, 0)
    assert np.array_equal(p.map_download(), np.array(o.map.pixels))
    p.close()

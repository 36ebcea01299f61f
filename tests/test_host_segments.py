"""Host side of the segments path (no GPU): pack_segments lays List<ScanSegment> out the way cs_update_segments takes it,
and the host twin of ScanSegmentsToCloud (CoreSLAMProcessor.cs:187-207) equals the oracle's restatement bit for bit."""
import numpy as np

import slam.net_b200 as sn
from slam.net_b200 import coreslam as cs
from oracle import oracle as orc


def _segments(rng, sizes):
    segs = []
    for k, n in enumerate(sizes):
        rays = np.stack([rng.uniform(-7, 7, n), rng.uniform(0.05, 30, n)], axis=1).astype(np.float32)
        pose = rng.normal(0, [3, 3, 2]).astype(np.float32)
        segs.append(sn.ScanSegment(Rays=rays if k % 2 else [sn.Ray(float(a), float(r)) for a, r in rays], Pose=pose,
                                   IsLast=(k == len(sizes) - 1)))
    return segs


def test_pack_segments_layout():
    segs = _segments(np.random.default_rng(1), [3, 0, 5, 1])
    rays, first, poses = cs.pack_segments(segs)
    assert rays.dtype == np.float32 and rays.shape == (9, 2)
    assert first.dtype == np.int32 and first.tolist() == [0, 3, 3, 8, 9]
    assert poses.shape == (4, 3) and np.array_equal(poses[2], np.asarray(segs[2].Pose, dtype=np.float32))
    assert np.array_equal(rays[3:8], segs[2].rays_array())
    r0, f0, p0 = cs.pack_segments([])
    assert r0.shape == (0, 2) and f0.tolist() == [0] and p0.shape == (0, 3)


def test_host_twin_matches_oracle_cloud():
    rng = np.random.default_rng(2)
    segs = _segments(rng, [40, 1, 0, 300])
    odo = np.asarray(segs[-1].Pose, dtype=np.float32)  # :719 the last segment's pose
    got = cs.scan_segments_to_cloud(segs, odo).Points
    want = np.concatenate([orc.segment_to_cloud(s.rays_array(), np.asarray(s.Pose, dtype=np.float32), odo) for s in segs
                           if s.rays_array().shape[0]])
    assert got.shape == want.shape == (341, 2)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # the last segment's own points are pure polar -> cartesian (pose - odo = 0)
    last = segs[-1].rays_array()
    c, s = orc.libm_sincos(last[:, 0])
    assert np.array_equal(got[-300:, 0], (np.float32(0) + last[:, 1] * c).astype(np.float32))

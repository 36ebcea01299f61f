"""ScanSegmentsToCloud (CoreSLAMProcessor.cs:187-207): a hand-derived known-answer vector and the cross check of the
two restatements (C oracle vs the statement-by-statement Python transliteration), plus Update(List<ScanSegment>) as a whole."""
import numpy as np

from oracle import oracle as orc
from oracle import transliteration as tr


def test_kat_i_segments_to_cloud_exact_angles():
    """KAT-I, traced by hand from :194-201.  Two segments, odometry pose = the second one's pose (1, 2, 0):
    segment A at (1.5, 2.25, 0): pose - odo = (0.5, 0.25, 0); rays (0, 2) -> (0.5 + 2*cos 0, 0.25 + 2*sin 0) = (2.5, 0.25)
    and (0, 0.5) -> (1.0, 0.25).  Segment B at (1, 2, 0): pose - odo = 0; ray (0, 4) -> (4, 0).  cosf(0) = 1 and sinf(0) = 0
    exactly in every libm, all sums exact in binary32."""
    segs = [((1.5, 2.25, 0.0), [(0.0, 2.0), (0.0, 0.5)]), ((1.0, 2.0, 0.0), [(0.0, 4.0)])]
    want = np.array([[2.5, 0.25], [1.0, 0.25], [4.0, 0.0]], dtype=np.float32)
    got_t = tr.scan_segments_to_cloud(segs, (1.0, 2.0, 0.0))
    assert np.array_equal(got_t, want)
    got_c = np.concatenate([orc.segment_to_cloud(np.array(r, dtype=np.float32), np.array(p, dtype=np.float32),
                                                 np.array([1.0, 2.0, 0.0], dtype=np.float32)) for p, r in segs])
    assert np.array_equal(got_c, want)
    # a quarter turn carried by the SEGMENT pose (angle + pose.Z, :200): (r cos(pi/2), r sin(pi/2)) with libm's float values
    hp = np.float32(np.pi / 2)
    q = tr.scan_segments_to_cloud([((0.0, 0.0, float(hp)), [(0.0, 3.0)])], (0.0, 0.0, 0.0))
    assert q[0, 1] == np.float32(3.0) and abs(float(q[0, 0])) < 1e-6 and q[0, 0] == np.float32(3.0) * tr.cosf(hp)


def test_cross_segments_to_cloud_random():
    rng = np.random.default_rng(12)
    for trial in range(20):
        segs = []
        for _ in range(rng.integers(1, 5)):
            n = int(rng.integers(0, 30))
            rays = np.stack([rng.uniform(-400, 400, n) if trial % 5 == 0 else rng.uniform(-7, 7, n), rng.uniform(0.01, 40, n)], axis=1).astype(np.float32)
            segs.append((rng.normal(0, [4, 4, 3]).astype(np.float32), rays))
        odo = segs[-1][0] + rng.normal(0, 0.01, 3).astype(np.float32)
        t = tr.scan_segments_to_cloud([(p, [tuple(r) for r in rays]) for p, rays in segs], odo)
        parts = [orc.segment_to_cloud(rays, p, odo) for p, rays in segs if rays.shape[0]]
        c = np.concatenate(parts) if parts else np.zeros((0, 2), dtype=np.float32)
        assert t.shape == c.shape and np.array_equal(t.view(np.uint32), c.view(np.uint32))


def test_cross_update_with_segments_small():
    """Whole Update(List<ScanSegment>) through the transliteration vs the C oracle fed the C oracle's cloud."""
    rng = np.random.default_rng(3)
    size, phys, iters, threads = 48, 12.0, 6, 2
    t = tr.ProcessorT(phys, size, (6.0, 6.0, 0.1), 0.1, 0.17, iters, threads)
    o = orc.Processor(phys, size, (6.0, 6.0, 0.1), 0.1, 0.17, iters, threads)
    pose = np.array([6.0, 6.0, 0.1], dtype=np.float32)
    for k in range(8):
        pose = pose + np.array([0.02, 0.01, 0.01], dtype=np.float32)
        segs = []
        for s in range(2):
            rays = np.stack([rng.uniform(-3.1, 3.1, 10), rng.uniform(0.5, 4.0, 10)], axis=1).astype(np.float32)
            segs.append(((pose - np.float32(0.01) * (1 - s)).astype(np.float32), rays))
        off = rng.normal(0, [0.1, 0.1, 0.17], (iters * threads, 3)).astype(np.float32)
        t.UpdateSegments([(p, [tuple(r) for r in rays]) for p, rays in segs], off)
        odo = segs[-1][0]
        o.update(np.concatenate([orc.segment_to_cloud(rays, p, odo) for p, rays in segs]), odo, off)
        assert np.array_equal(np.array(t.Pose, dtype=np.float32), o.pose), k
    assert np.array_equal(np.asarray(t.HoleMap.Pixels).reshape(-1), np.array(o.map.pixels))

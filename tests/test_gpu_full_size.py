"""BASELINE configs[3] and [4] at their full sizes: the 65 536-candidate search on the 8192 x 8192 map
with every distance compared (this is also the only path through the two-kernel candidate sort at its real size), whole
Updates of that configuration in both candidate modes, and a 256-session batch whose sampled sessions must equal the CPU
oracle in pose and map checksum.  The oracle only checks."""
import numpy as np
import pytest

import slam.net_b200 as sn
from slam.net_b200 import _native as N
from slam.net_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_cfg4_full_size_65536_candidates_distances_and_argmin():
    size, phys, iters, threads, P = 8192, 81.92, 1024, 64, 1024
    n_cand = iters * threads
    rp = synth.make_replay(4, P, phys)
    p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, max_points=P, flags=N.FLAG_KEEP_DISTANCES)
    assert p.search_plan(P)["slab"]
    m = orc.HoleMap(size, phys)
    m.fill(32750)
    for k in range(3):
        pose = rp.truth[k].astype(np.float32)
        orc.update_hole_map(m, rp.points[k], pose, 0.6, 50)
        p.integrate(rp.points[k], pose)
    assert p.map_checksum() == sn.host_map_checksum(np.array(m.pixels), size)
    sp = rp.truth[3].astype(np.float32)
    # uploaded table (verification mode) ...
    off = synth.candidate_offsets(31, 0, n_cand, 0.1, 0.17)
    best, bd, d, bi = orc.parallel_search(m, rp.points[3], sp, off, iters, threads)
    res, dist = p.search(rp.points[3], sp, sp[None, :] + off)
    assert dist.shape == (n_cand + 1,) and np.array_equal(dist, d)
    assert (res.distance, res.index) == (bd, bi)
    # ... and a table with ties and out-of-map candidates at full size: lowest flat index wins
    off2 = off.copy()
    off2[1000:3000] = off2[7]           # 2000 duplicates of candidate 8
    off2[50000:50100, 0] = 1.0e6        # far off the map: int.MaxValue
    best, bd, d, bi = orc.parallel_search(m, rp.points[3], sp, off2, iters, threads)
    res, dist = p.search(rp.points[3], sp, sp[None, :] + off2)
    assert np.array_equal(dist, d) and (res.distance, res.index) == (bd, bi)
    assert np.all(dist[50001:50101] == 2147483647)
    p.close()


@pytest.mark.parametrize("philox", [False, True])
def test_cfg4_full_size_update_replay(philox):
    """Whole Updates of configs[3] — 65 537 poses x 1024 points on the 128 MB map — poses, winners and the map."""
    size, phys, iters, threads, P, n_scans = 8192, 81.92, 1024, 64, 1024, 8
    n_cand = iters * threads
    seed = 0xC0FFEE
    rp = synth.make_replay(n_scans, P, phys, seed=44)
    p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, max_points=P, seed=seed)
    o = orc.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads)
    w = orc.Worker(threads)  # one oracle thread per search thread of the reference (ParallelWorker)
    for k in range(n_scans):
        off = sn.philox_offsets(seed, k, n_cand, 0.1, 0.17) if philox else synth.candidate_offsets(45, k, n_cand, 0.1, 0.17)
        r = p.update(rp.points[k], rp.odometry[k], None if philox else off)
        o.update(rp.points[k], rp.odometry[k], off, worker=w)
        assert np.array_equal(r.pose, o.pose), k
        if k >= 5:  # (the threaded oracle keeps the winning distance and pose, not the flat index)
            assert r.distance == o.last_distance and 0 <= r.index <= n_cand, k
    w.close()
    assert p.map_checksum() == sn.host_map_checksum(np.array(o.map.pixels), size)
    p.close()


def test_cfg5_256_sessions_sampled_against_the_oracle():
    """A quarter of configs[4] on one GPU: 256 sessions of the cfg1 geometry (360 points, 1000 iterations, 1600 x 1600 map),
    production mode, parameter grid as in bench.py; 8 sampled sessions must equal the oracle (pose and map checksum), and
    the sweep must really have produced different sessions."""
    n_sess, P, size, phys, iters, threads, n_scans = 256, 360, 1600, 40.0, 1000, 1, 9
    n_cand = iters * threads
    rp = synth.make_replay(n_scans, P, phys, seed=77)

    def params(s):
        return (np.float32(0.05 + 0.025 * (s % 5)), np.float32(0.0873 + 0.0436 * ((s // 5) % 4)), 30 + 20 * ((s // 20) % 6),
                np.float32(0.4 + 0.2 * ((s // 120) % 4)), 9000 + s)

    prm = [params(s) for s in range(n_sess)]
    b = sn.Batch(n_sess, phys, size, rp.odometry[0], [q[0] for q in prm], [q[1] for q in prm], iters, threads, max_points=P,
                 seeds=[q[4] for q in prm])
    for j, q in enumerate(prm):
        b.set_params(j, q[2], float(q[3]))
    log = sn.ScanLog(n_scans, P, n_offsets=0)
    for k in range(n_scans):
        log.set(k, rp.points[k], rp.odometry[k])
    log.upload()
    res = b.replay(log, 0, n_scans)
    poses, sums = b.poses(), b.map_checksums()
    for s in (0, 37, 64, 101, 128, 199, 230, 255):
        sxy, sth, q, hw, seed = prm[s]
        o = orc.Processor(phys, size, rp.odometry[0], sxy, sth, iters, threads)
        o.quality, o.hole_width = q, hw
        for k in range(n_scans):
            o.update(rp.points[k], rp.odometry[k], sn.philox_offsets(seed, k, n_cand, sxy, sth) if k >= 5 else None)
        assert np.array_equal(poses[s], o.pose), s
        assert np.array_equal(res[s].pose, o.pose) and (res[s].distance, res[s].index) == (o.last_distance, o.last_index), s
        assert int(sums[s]) == sn.host_map_checksum(np.array(o.map.pixels), size), s
    assert len(set(int(x) for x in sums)) > n_sess // 2
    log.close()
    b.close()

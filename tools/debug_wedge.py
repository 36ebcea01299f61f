"""Diagnostics: cs_integrate vs the oracle on awkward scans; prints where the maps differ (ring, position)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import slam.net_b200 as sn
from slam.net_b200 import synth
from oracle import oracle as orc

size, phys = 512, 40.0
bad = 0
for trial, (P, keep, seed) in enumerate([(2600, 700, 152), (2600, 1500, 151), (2600, 2600, 150), (360, 360, 1), (1024, 300, 2), (5000, 1200, 3)]):
    rp = synth.make_replay(4, P, phys, seed=seed)
    p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.1, 4, 1, max_points=P)
    m = orc.HoleMap(size, phys)
    m.fill(32750)
    for k in range(4):
        pts = rp.points[k][:keep]
        pose = rp.odometry[k]
        p.integrate(pts, pose)
        orc.update_hole_map(m, pts, pose, 0.6, 50)
        got = p.map_download().reshape(size, size).astype(int)
        want = np.array(m.pixels).reshape(size, size).astype(int)
        d = np.argwhere(got != want)
        if len(d):
            bad += 1
            x1 = int(np.float32(pose[0]) * np.float32(size / phys) + np.float32(0.5)); y1 = int(np.float32(pose[1]) * np.float32(size / phys) + np.float32(0.5))
            ring = np.maximum(np.abs(d[:, 1] - x1), np.abs(d[:, 0] - y1))
            print("trial %d scan %d: %d cells differ; rings %s" % (trial, k, len(d), sorted(set(ring.tolist()))[:40]))
            for (y, x) in d[:12]:
                print("   cell (%d,%d) rel (%d,%d) got %d want %d" % (x, y, x - x1, y - y1, got[y, x], want[y, x]))
            p.map_upload(want.astype(np.uint16).reshape(-1))
    p.close()
print("mismatching scans:", bad)

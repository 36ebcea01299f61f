"""Diagnostics: per-ring cycle counts of the integrate kernel on the cfg2 workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import slam.net_b200 as sn
from slam.net_b200 import synth
P, size = 1024, 2048
rp = synth.make_replay(12, P, 40.0)
p = sn.Processor(40.0, size, rp.odometry[0], 0.1, 0.17, 1024, 4, max_points=P, flags=2)
p.ring_cycles()
for k in range(12):
    off = synth.candidate_offsets(1, k, 4096, 0.1, 0.17)
    r = p.update(rp.points[k], rp.odometry[k], off)
    p.sync()
    t = p.timing()
    rc = p.ring_cycles()
    tot = rc[:, 0]
    nz = np.nonzero(tot)[0]
    print("scan %d search %.1fus rings-kernel %.1fus visits %d rings %d max %d at ring %d median %d" % (
        k, t.search_ms * 1e3, t.integrate_ms * 1e3, p.visits(), len(nz), tot.max(), tot.argmax(), np.median(tot[nz])))
for ring in (0, 1, 2, 5, 10, 30, 100, 200, 300, 500, 800):
    print("ring %4d: total %7d  sess-loads %d clear %d ray-loads %d compute %d collect %d table %d apply %d" % (ring, rc[ring, 0], rc[ring, 4], rc[ring, 5], rc[ring, 6], rc[ring, 7], rc[ring, 1], rc[ring, 2], rc[ring, 3]))

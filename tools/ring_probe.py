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
    nz = np.nonzero(rc)[0]
    print("scan %d search %.1fus fin %.1fus integ %.1fus visits %d rings %d max_cycles %d at ring %d; ring0 %d ring1 %d ring10 %d ring100 %d ring500 %d median %d" % (
        k, t.search_ms*1e3, t.finalize_ms*1e3, t.integrate_ms*1e3, r.visits, len(nz), rc.max(), rc.argmax(), rc[0], rc[1], rc[10], rc[100], rc[500], np.median(rc[nz])))
top = np.argsort(-rc)[:12]
print("slowest rings:", [(int(i), int(rc[i])) for i in top])

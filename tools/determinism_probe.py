"""Diagnostics: is a long cfg2 replay reproducible run to run, and equal to the oracle?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import slam.net_b200 as sn
from slam.net_b200 import synth, _native as N
from oracle import oracle as orc
P, size = 1024, 2048
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rp = synth.make_replay(n, P, 40.0)
offs = [synth.candidate_offsets(1, k, 4096, 0.1, 0.17) for k in range(n)]
log = sn.ScanLog(n, P, n_offsets=4096)
for k in range(n):
    log.set(k, rp.points[k], rp.odometry[k], offs[k])
log.upload()
runs = []
for rep in range(3):
    p = sn.Processor(40.0, size, rp.odometry[0], 0.1, 0.17, 1024, 4, max_points=P)
    res = p.replay(log, 0, n)
    runs.append((np.array([r.pose for r in res]), np.array([r.visits for r in res]), p.map_checksum(), p.map_download()))
    p.close()
for rep in (1, 2):
    same = np.array_equal(runs[0][0], runs[rep][0])
    first = -1 if same else int(np.nonzero((runs[0][0] != runs[rep][0]).any(axis=1))[0][0])
    print("run %d vs run 0: poses equal %s (first diff at scan %d), checksum equal %s, cells differing %d" % (
        rep, same, first, runs[0][2] == runs[rep][2], int(np.count_nonzero(runs[0][3] != runs[rep][3]))))
o = orc.Processor(40.0, size, rp.odometry[0], 0.1, 0.17, 1024, 4)
w = orc.Worker(4)
bad = -1
for k in range(n):
    o.update(rp.points[k], rp.odometry[k], offs[k], worker=w)
    if bad < 0 and not np.array_equal(o.pose, runs[0][0][k]):
        bad = k
        m = np.array(o.map.pixels)
print("oracle vs run 0: first pose diff at scan %d; final map cells differing %d" % (bad, int(np.count_nonzero(np.array(o.map.pixels) != runs[0][3]))))
w.close()

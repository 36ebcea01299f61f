"""Diagnostics: where a cfg4 Update (65 537 candidates x 1024 points, 8192 x 8192 map, production mode) spends its time on
ONE GPU, per kernel stage (CS_FLAG_TIMING: events around the search stage and the draw kernel, kernels serialised), for the
whole candidate set and for the slices a rank of a 2 / 4 / 8-GPU group evaluates (cs_update_begin with a slice)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import slam.net_b200 as sn
from slam.net_b200 import _native as N, synth

P, size, phys, iters, threads = 1024, 8192, 81.92, 1024, 64
n_scans = 30
rp = synth.make_replay(n_scans, P, phys)
p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17453292, iters, threads, max_points=P, seed=7)
log = sn.ScanLog(n_scans, P, n_offsets=0)
for k in range(n_scans):
    log.set(k, rp.points[k], rp.odometry[k])
log.upload()
p.replay(log, 0, 10, want_results=False)
p.sync()
p.set_flags(N.FLAG_TIMING)
s, i, t = [], [], []
for k in range(10, 30):
    p.replay(log, k, 1, want_results=True)
    tm = p.timing()
    s.append(tm.search_ms); i.append(tm.integrate_ms); t.append(tm.total_device_ms)
print("whole candidate set: search stage %.1f us, draw kernel %.1f us, total (serialised) %.1f us" % (np.mean(s) * 1e3, np.mean(i) * 1e3, np.mean(t) * 1e3))
p.set_flags(0)
import time
for world in (1, 2, 4, 8):
    n_flat = iters * threads + 1
    cnt = n_flat // world
    # the search of one rank's slice alone (cs_update_begin launches only the search kernels), 20 scans, wall clock over sync
    p.sync()
    t0 = time.perf_counter()
    for k in range(10, 30):
        p.update_begin(rp.points[k], rp.odometry[k], None, 0, cnt)
        p.update_finish()
    p.sync()
    dt = (time.perf_counter() - t0) / 20
    print("world %d: slice of %6d candidates, begin+finish through the host (no exchange): %.1f us per Update" % (world, cnt, dt * 1e6))
p.close()

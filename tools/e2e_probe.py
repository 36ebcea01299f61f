"""Diagnostics: where a cs_update call spends its time on the cfg2 workload (host wall vs device events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import slam.net_b200 as sn
from slam.net_b200 import synth, _native as N
P, size, n = 1024, 2048, 120
rp = synth.make_replay(n, P, 40.0)
offs = [synth.candidate_offsets(1, k, 4096, 0.1, 0.17) for k in range(n)]
for flags in (0, N.FLAG_TIMING, N.FLAG_NO_HOST_SPIN):
    p = sn.Processor(40.0, size, rp.odometry[0], 0.1, 0.17, 1024, 4, max_points=P, flags=flags)
    lat, rows = [], []
    t0 = time.perf_counter()
    for k in range(n):
        ta = time.perf_counter()
        r = p.update(rp.points[k], rp.odometry[k], offs[k])
        lat.append(time.perf_counter() - ta)
        if flags & N.FLAG_TIMING:
            t = p.timing()
            rows.append((t.h2d_ms, t.search_ms, t.integrate_ms, t.total_device_ms, t.host_wait_ms))
    p.sync()
    dt = time.perf_counter() - t0
    lat = np.array(lat[20:]) * 1e6
    print("flags=%d: %.1f us/step wall; latency p50 %.1f p90 %.1f p99 %.1f max %.1f us" % (flags, dt / n * 1e6, np.percentile(lat, 50), np.percentile(lat, 90), np.percentile(lat, 99), lat.max()))
    if rows:
        a = np.array(rows[20:]) * 1e3
        print("   device us: h2d %.1f search %.1f integrate %.1f total %.1f host_wait %.1f" % tuple(a.mean(axis=0)))
    p.close()

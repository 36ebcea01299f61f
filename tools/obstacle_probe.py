"""Device time of UpdateObstacleMap next to the rest of an Update (cfg2 geometry), for several ObstacleMap sizes."""
import sys, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import slam.net_b200 as sn
from slam.net_b200 import _native as N, synth

P, size, phys, iters, threads = 1024, 2048, 40.0, 1024, 4
n = 60
rp = synth.make_replay(n, P, phys)
for obst in (0, 64, 256, 2048, 8192):
    p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17453292, iters, threads, max_points=P, obstacle_map_size=obst, seed=1,
                     flags=N.FLAG_TIMING)
    log = sn.ScanLog(n, P, n_offsets=0)
    for k in range(n):
        log.set(k, rp.points[k], rp.odometry[k])
    log.upload()
    p.replay(log, 0, 20, want_results=False)
    s, i, o, t = [], [], [], []
    for k in range(20, n):
        p.replay(log, k, 1, want_results=True)
        tm = p.timing()
        s.append(tm.search_ms); i.append(tm.integrate_ms); o.append(tm.obstacle_ms); t.append(tm.total_device_ms)
    p.set_flags(0)
    print(json.dumps({"obstacle_map_size": obst, "search_us": 1e3 * float(np.mean(s)), "rings_us": 1e3 * float(np.mean(i)),
                      "obstacle_us": 1e3 * float(np.mean(o)), "total_us": 1e3 * float(np.mean(t)),
                      "touched_per_scan": (p.obstacle_visits() / n) if obst else 0}))
    p.close(); log.close()

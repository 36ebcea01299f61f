"""Diagnostics: block-level timeline (%globaltimer stamps) of one cfg2 Update — where the 45 us go."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import slam.net_b200 as sn
from slam.net_b200 import synth

P, size, n_scans = 1024, 2048, 14
rp = synth.make_replay(n_scans, P, 40.0)
p = sn.Processor(40.0, size, rp.odometry[0], 0.1, 0.17, 1024, 4, max_points=P)
log = sn.ScanLog(n_scans, P, n_offsets=4096)
for k in range(n_scans):
    log.set(k, rp.points[k], rp.odometry[k], synth.candidate_offsets(1, k, 4096, 0.1, 0.17))
log.upload()
p.ring_cycles()
p.replay(log, 0, n_scans - 3, want_results=False)
p.sync()


def pct(a, q):
    return float(np.percentile(a, q))


for rep in range(2):
    # two steps back to back: the second one is the steady state (its search overlaps the first one's tail)
    p.replay(log, n_scans - 3 + rep, 1, want_results=False)
    p.sync()
    rc = p.ring_cycles(size + 8192)
    rg = rc[:size]
    se = rc[size:]
    se = se[:8190]
    se = se[se[:, 1] > 0]
    rg = rg[rg[:, 2] > 0]
    t0 = se[:, 1].min()
    print("---- step %d: %d search blocks on %d SMs, %d rings blocks on %d SMs" % (
        rep, len(se), len(np.unique(se[:, 0])), len(rg), len(np.unique(rg[:, 1]))))
    f = lambda col: "min %6.2f p50 %6.2f p90 %6.2f max %6.2f" % tuple((np.array([col.min(), pct(col, 50), pct(col, 90), col.max()]) - t0) / 1e3)
    print("search start        ", f(se[:, 1]))
    print("search wait over    ", f(se[:, 2]))
    print("search staged       ", f(se[:, 3]))
    print("search warp0 done   ", f(se[:, 4]))
    print("search block done   ", f(se[:, 5]))
    last = se[se[:, 7] == 1]
    if len(last):
        print("publish done         %6.2f (its block done at %6.2f)" % ((last[0, 6] - t0) / 1e3, (last[0, 5] - t0) / 1e3))
    d = rc[size + 8191]
    print("publisher: entry %.2f pose known %.2f flag A %.2f host record %.2f" % tuple((d[:4] - t0) / 1e3))
    d = rc[size + 8190]
    print("prep block 0: flag A seen %.2f pose loaded %.2f rays computed %.2f fence done %.2f counted %.2f" % tuple((d[:5] - t0) / 1e3))
    per_sm = np.bincount(se[:, 0].astype(int))
    per_sm = per_sm[per_sm > 0]
    print("search blocks per SM: min %d max %d; loop time per block us: p50 %.2f max %.2f" % (
        per_sm.min(), per_sm.max(), pct(se[:, 5] - se[:, 3], 50) / 1e3, (se[:, 5] - se[:, 3]).max() / 1e3))
    print("rings start         ", f(rg[:, 2]))
    print("rings wait over     ", f(rg[:, 3]))
    print("rings prep seen     ", f(rg[:, 4][rg[:, 4] > 0]))
    print("rings rays loaded   ", f(rg[:, 5][rg[:, 5] > 0]))
    print("rings end           ", f(rg[:, 6]))
    body = (rg[:, 6] - rg[:, 5])[rg[:, 5] > 0] / 1e3
    print("rings body per block us: p10 %.2f p50 %.2f p90 %.2f max %.2f" % (pct(body, 10), pct(body, 50), pct(body, 90), body.max()))
    order = np.argsort(rg[:, 2])
    print("rings blocks by start time (idx: start, end):", ", ".join("%d: %.1f-%.1f" % (i, (rg[i, 2] - t0) / 1e3, (rg[i, 6] - t0) / 1e3) for i in order[::max(1, len(order) // 16)]))
    worst = np.argsort(-body)[:8]
    print("slowest rings blocks:", ", ".join("%d: %.1f us" % (np.nonzero(rg[:, 5] > 0)[0][i], body[i]) for i in worst))

"""Diagnostics: per-task timeline (%globaltimer stamps) of the wedge kernel for one Update.
usage: python tools/wedge_probe.py [cfg2|cfg3|cfg1|cfg4]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import slam.net_b200 as sn
from slam.net_b200 import synth

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
if wl == "custom":  # custom P size phys iters threads
    P, size, phys, iters, threads = int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
else:
  P, size, phys, iters, threads = {"cfg2": (1024, 2048, 40.0, 1024, 4), "cfg3": (8192, 4096, 40.96, 1, 1), "cfg1": (360, 1600, 40.0, 1000, 1),
                                   "cfg4": (1024, 8192, 81.92, 1024, 64)}[wl]
n_scans = 12
rp = synth.make_replay(n_scans, P, phys)
p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, max_points=P)
if wl == "cfg3":
    p.set_position_search_beginning(2 ** 31 - 1)
log = sn.ScanLog(n_scans, P, n_offsets=0)
for k in range(n_scans):
    log.set(k, rp.points[k], rp.odometry[k])
log.upload()
p.ring_cycles()
p.replay(log, 0, n_scans - 2, want_results=False)
p.sync()
FLUSH = os.environ.get("PROBE_FLUSH") == "1"  # write 256 MB before each probed step, like bench.py's default line
if FLUSH:
    import torch
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rep in range(2):
    p.ring_cycles()  # (clears nothing; records are overwritten by the next step)
    if FLUSH:
        flush_buf.fill_(rep + 1)
        torch.cuda.synchronize()
    p.replay(log, n_scans - 2 + rep, 1, want_results=False)
    p.sync()
    full = p.ring_cycles(size + 8192)
    pub = full[size + 8191]
    rc = full[:size]
    rc = rc[rc[:, 0] > 0]
    t0 = rc[rc[:, 4] > 0, 4].min()
    us = lambda a: (a - t0) / 1e3
    se = full[size:size + 8190]
    se = se[se[:, 1] > t0 - 200000]  # this step's search blocks
    if len(se):
        print("search blocks: %d; start min %.2f p50 %.2f max %.2f | wait over p50 %.2f max %.2f | lookups done p50 %.2f max %.2f | atomic back p50 %.2f max %.2f" % (
            len(se), us(se[:, 1]).min(), np.percentile(us(se[:, 1]), 50), us(se[:, 1]).max(), np.percentile(us(se[:, 2]), 50), us(se[:, 2]).max(),
            np.percentile(us(se[:, 4]), 50), us(se[:, 4]).max(), np.percentile(us(se[:, 5]), 50), us(se[:, 5]).max()))
    if pub[2] > 0:
        print("pose: publisher entry %.2f, pose known %.2f, pose words out %.2f, record out %.2f" % tuple((pub[:4] - t0) / 1e3),
              "(glue from the table)" if pub[4] == 1 else "(glue computed by the publisher)")
    print("---- step %d: %d tasks recorded on %d SMs" % (rep, len(rc), len(np.unique(rc[:, 7]))))
    q = lambda a: "min %6.2f p50 %6.2f p90 %6.2f max %6.2f" % (a.min(), np.percentile(a, 50), np.percentile(a, 90), a.max())
    print("block start      ", q(us(rc[:, 4])))
    print("rays ready       ", q(us(rc[:, 5])))
    print("table ready      ", q(us(rc[:, 6])))
    print("task start       ", q(us(rc[:, 0])))
    print("task filtered    ", q(us(rc[:, 1])))
    print("task end         ", q(us(rc[:, 2])))
    blk = (rc[:, 7] >> 16).astype(int)
    ub, first = np.unique(blk, return_index=True)
    sel = first[:: max(1, len(first) // 24)]
    print("blocks (id: start / rays ready / table ready):", ", ".join("%d: %.1f/%.1f/%.1f" % (blk[i], us(rc[i, 4]), us(rc[i, 5]), us(rc[i, 6])) for i in sel))
    k0 = (rc[:, 3] >> 32).astype(int)
    cand = (rc[:, 3] & 0xffff).astype(int)
    mode = ((rc[:, 3] >> 16) & 0xf).astype(int)
    pieces = ((rc[:, 3] >> 20) & 0xfff).astype(int)
    for k in sorted(set(k0)):
        m = k0 == k
        print("k0 %5d: %4d tasks  cand p50 %3d max %3d  general %3d  pieces max %d | filter us p50 %5.2f max %5.2f | body us p50 %6.2f max %6.2f | end max %6.2f" % (
            k, m.sum(), np.percentile(cand[m], 50), cand[m].max(), mode[m].sum(), pieces[m].max(),
            np.percentile((rc[m, 1] - rc[m, 0]) / 1e3, 50), ((rc[m, 1] - rc[m, 0]) / 1e3).max(),
            np.percentile((rc[m, 2] - rc[m, 1]) / 1e3, 50), ((rc[m, 2] - rc[m, 1]) / 1e3).max(), us(rc[m, 2]).max()))

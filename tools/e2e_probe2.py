import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import slam.net_b200 as sn
from slam.net_b200 import synth, _native as N
P, size, n = 1024, 2048, 200
rp = synth.make_replay(n, P, 40.0)
offs = [synth.candidate_offsets(1, k, 4096, 0.1, 0.17) for k in range(n)]
def e2e(p, k0, cnt, tag):
    L = sn.lib(); fp = C.POINTER(C.c_float); res = N.Result()
    lat = []
    for k in range(k0, k0 + cnt):
        pts = rp.points[k].ctypes.data_as(fp); odo = np.ascontiguousarray(rp.odometry[k]); offp = offs[k].ctypes.data_as(fp)
        ta = time.perf_counter()
        st = L.cs_update(p._h, pts, P, odo.ctypes.data_as(fp), offp, C.byref(res))
        lat.append(time.perf_counter() - ta)
        assert st == 0
    p.sync()
    lat = np.array(lat) * 1e6
    print("%s: p50 %.1f p90 %.1f max %.1f" % (tag, np.percentile(lat, 50), np.percentile(lat, 90), lat.max()))
for use_torch_stream in (False, True):
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        p = sn.Processor(40.0, size, rp.odometry[0], 0.1, 0.17, 1024, 4, max_points=P, stream=stream.cuda_stream if use_torch_stream else 0)
        e2e(p, 0, 40, "torch_stream=%s fresh" % use_torch_stream)
        log = sn.ScanLog(n, P, n_offsets=4096)
        for k in range(n):
            log.set(k, rp.points[k], rp.odometry[k], offs[k])
        log.upload()
        p.replay(log, 40, 20, want_results=False)
        e2e(p, 60, 30, "  after replay")
        p.set_flags(N.FLAG_TIMING)
        for i in range(10):
            p.replay(log, 90 + i, 1, want_results=True)
        p.set_flags(0)
        e2e(p, 100, 30, "  after timing pass + set_flags(0)")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        flush.fill_(1)
        torch.cuda.synchronize()
        e2e(p, 130, 30, "  after flush alloc")
        p.close(); log.close()

"""Asynchronous map export: cs_map_export_begin / cs_map_export_wait deliver the map AS OF the call
(after the last integration, before the next Update) in the reference's viewer formats, while further Updates run."""
import ctypes as C

import numpy as np
import pytest

import slam.net_b200 as sn
from slam.net_b200 import _native as N
from slam.net_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _pinned(nbytes, dtype):
    ptr = C.c_void_p()
    N.check(N.lib().cs_pinned_alloc(C.byref(ptr), nbytes))
    buf = (C.c_uint8 * nbytes).from_address(ptr.value)
    return ptr, np.frombuffer(buf, dtype=dtype)


@pytest.mark.parametrize("layout", [0, N.FLAG_ROW_MAJOR_MAP])
def test_export_snapshots_while_updates_continue(layout):
    n_scans, P, size, obst, phys, iters, threads = 16, 300, 320, 160, 40.0, 40, 2
    rp = synth.make_replay(n_scans, P, phys, seed=8)
    p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, max_points=P, flags=layout, obstacle_map_size=obst)
    o = orc.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, obstacle_map_size=obst)
    ptr_g, gray = _pinned(size * size * 2, np.uint16)
    ptr_k, packed = _pinned(size * size // 2, np.uint8)
    ptr_o, obs = _pinned(obst * obst, np.int8)
    snaps = {}
    for k in range(n_scans):
        off = synth.candidate_offsets(3, k, iters * threads, 0.1, 0.17)
        p.update(rp.points[k], rp.odometry[k], off)
        o.update(rp.points[k], rp.odometry[k], off)
        if k == 6:
            p.map_export_begin(N.EXPORT_GRAY16, gray)
            snaps["gray"] = np.array(o.map.pixels).copy()
        if k == 9:
            p.map_export_wait()
            assert np.array_equal(gray, snaps["gray"])  # the map of scan 6, although scans 7..9 have been integrated since
            p.map_export_begin(N.EXPORT_PACKED4, packed)
            snaps["packed"] = o.map.packed().copy()
        if k == 12:
            p.map_export_begin(N.EXPORT_OBSTACLE_I8, obs)  # queues behind the packed export, no wait in between
            snaps["obs"] = np.array(o.obstacle_map.pixels).reshape(-1).copy()
    p.map_export_wait()
    assert np.array_equal(obs, snaps["obs"])
    assert np.array_equal(packed, snaps["packed"])  # complete before the obstacle export could start (stream order)
    assert np.array_equal(p.map_download(), np.array(o.map.pixels))  # and the exports disturbed nothing
    p.close()
    for ptr in (ptr_g, ptr_k, ptr_o):
        N.lib().cs_pinned_free(ptr)


def test_export_packed_matches_sync_accessor_and_errors():
    size, phys = 128, 8.0
    p = sn.Processor(phys, size, (4, 4, 0), 0.1, 0.17, 4, 1)
    px = synth.random_map(size, 5)
    p.map_upload(px)
    out = np.zeros(size * size // 2, dtype=np.uint8)
    p.map_export_begin(N.EXPORT_PACKED4, out)
    p.map_export_wait()
    assert np.array_equal(out, p.map_packed())
    want = ((px[0::2] >> 12) << 4 | (px[1::2] >> 12)).astype(np.uint8)  # HoleMap.cs:51
    assert np.array_equal(out, want)
    with pytest.raises(sn.CoreSlamError):
        p.map_export_begin(N.EXPORT_OBSTACLE_I8, out)  # no ObstacleMap on this handle
    with pytest.raises(sn.CoreSlamError):
        p.map_export_begin(7, out)
    p.map_export_wait()  # nothing in flight: returns at once
    p.close()

"""CSLG scan-log files: the numpy writer/reader and the library's host-only reader agree on the format
(no GPU needed); on a GPU, save -> load -> replay gives the same poses and map as the log it was saved from."""
import ctypes as C

import numpy as np
import pytest

import slam.net_b200 as sn
from slam.net_b200 import _native as N
from slam.net_b200 import scanlog_file as sf
from slam.net_b200 import synth


def _drive(n_scans, P, n_off, ragged=True):
    rp = synth.make_replay(n_scans, P, 40.0, seed=3)
    pts = [rp.points[k][: P - (k % 7 if ragged else 0)] for k in range(n_scans)]
    offs = [synth.candidate_offsets(2, k, n_off, 0.1, 0.17) for k in range(n_scans)] if n_off else None
    return rp, pts, offs


@pytest.mark.parametrize("n_off", [0, 24])
def test_numpy_writer_and_library_reader_agree(tmp_path, n_off):
    rp, pts, offs = _drive(9, 50, n_off)
    path = str(tmp_path / "drive.cslg")
    sf.write_scanlog(path, pts, rp.odometry[:9], offs)
    p2, odo2, offs2, mp = sf.read_scanlog(path)
    assert mp == 50 and np.array_equal(odo2, rp.odometry[:9].astype(np.float32))
    assert all(np.array_equal(a, b) for a, b in zip(p2, pts))
    assert (offs2 is None) == (n_off == 0) and (n_off == 0 or all(np.array_equal(a, b) for a, b in zip(offs2, offs)))
    # the C reader (host-only entry points of the library)
    L = N.lib()
    n, m, no = C.c_int32(), C.c_int32(), C.c_int32()
    assert L.cs_scanlog_file_info(path.encode(), C.byref(n), C.byref(m), C.byref(no)) == 0
    assert (n.value, m.value, no.value) == (9, 50, n_off)
    for k in (0, 4, 8):
        buf = np.zeros((50, 2), dtype=np.float32)
        ob = np.zeros((max(n_off, 1), 3), dtype=np.float32)
        odo = np.zeros(3, dtype=np.float32)
        cnt = C.c_int32()
        assert L.cs_scanlog_file_read(path.encode(), k, buf.ctypes.data_as(N._fp), C.byref(cnt), odo.ctypes.data_as(N._fp),
                                      ob.ctypes.data_as(N._fp) if n_off else None) == 0
        assert cnt.value == pts[k].shape[0] and np.array_equal(buf[: cnt.value], pts[k]) and np.array_equal(odo, rp.odometry[k])
        if n_off:
            assert np.array_equal(ob, offs[k])
    assert L.cs_scanlog_file_read(path.encode(), 9, None, C.byref(cnt), odo.ctypes.data_as(N._fp), None) == 1  # past the end


def test_bad_files_are_rejected(tmp_path):
    L = N.lib()
    bad = tmp_path / "bad.cslg"
    bad.write_bytes(b"NOPE" + bytes(28))
    assert L.cs_scanlog_file_info(str(bad).encode(), None, None, None) == 1
    assert L.cs_scanlog_file_info(str(tmp_path / "missing.cslg").encode(), None, None, None) == 1
    rp, pts, offs = _drive(3, 20, 0, ragged=False)
    path = tmp_path / "cut.cslg"
    sf.write_scanlog(str(path), pts, rp.odometry[:3])
    path.write_bytes(path.read_bytes()[:-40])  # truncated last record
    cnt, odo = C.c_int32(), np.zeros(3, dtype=np.float32)
    buf = np.zeros((20, 2), dtype=np.float32)
    assert L.cs_scanlog_file_read(str(path).encode(), 1, buf.ctypes.data_as(N._fp), C.byref(cnt), odo.ctypes.data_as(N._fp), None) == 0
    assert L.cs_scanlog_file_read(str(path).encode(), 2, buf.ctypes.data_as(N._fp), C.byref(cnt), odo.ctypes.data_as(N._fp), None) == 1
    with pytest.raises(ValueError):
        sf.read_scanlog(str(path))


@pytest.mark.gpu
def test_save_load_replay_roundtrip(tmp_path):
    n_scans, P, size, iters, threads = 20, 200, 256, 32, 2
    rp, pts, offs = _drive(n_scans, P, iters * threads)
    log = sn.ScanLog(n_scans, P, n_offsets=iters * threads)
    for k in range(n_scans):
        log.set(k, pts[k], rp.odometry[k], offs[k])
    log.upload()
    path = str(tmp_path / "drive.cslg")
    log.save(path)
    # the library's file equals the numpy writer's byte for byte
    sf.write_scanlog(str(tmp_path / "np.cslg"), pts, rp.odometry[:n_scans], offs, max_points=log.max_points)
    assert open(path, "rb").read() == open(str(tmp_path / "np.cslg"), "rb").read()
    log2 = sn.ScanLog.load(path)
    assert (log2.n_scans, log2.n_offsets) == (n_scans, iters * threads)
    res = []
    for lg in (log, log2):
        q = sn.Processor(40.0, size, rp.odometry[0], 0.1, 0.17, iters, threads, max_points=P)
        r = q.replay(lg)
        res.append((r, q.map_download()))
        q.close()
    assert all(np.array_equal(a.pose, b.pose) and a.index == b.index for a, b in zip(res[0][0], res[1][0]))
    assert np.array_equal(res[0][1], res[1][1])
    log.close()
    log2.close()

"""CSLG scan-log files in plain numpy (host-only twin of cs_scanlog_save / cs_scanlog_load; the format is specified in
include/coreslam_b200.h).  A recorded drive — per scan: the cloud (ScanCloud.Points), the odometry pose and, optionally,
the candidate deviates the reference would dequeue — so that the same log can be replayed through the reference, the CPU
oracle and the CUDA path.  The reference itself has no log format (Simulation/MainWindow.xaml.cs:380-407 generates scans
live)."""
import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np

MAGIC, VERSION = b"CSLG", 1


def write_scanlog(path, points: Sequence[np.ndarray], odometry: Sequence[Sequence[float]], offsets: Optional[Sequence[np.ndarray]] = None,
                  max_points: Optional[int] = None):
    """points[k]: (n_k, 2) float32, odometry[k]: 3 floats, offsets[k]: (n_offsets, 3) float32 or None for all scans."""
    n_scans = len(points)
    if n_scans == 0 or len(odometry) != n_scans or (offsets is not None and len(offsets) != n_scans):
        raise ValueError("need one odometry pose (and one offset table) per scan")
    pts = [np.ascontiguousarray(p, dtype="<f4").reshape(-1, 2) for p in points]
    n_off = 0 if offsets is None else int(np.asarray(offsets[0]).reshape(-1, 3).shape[0])
    mp = max(p.shape[0] for p in pts) if max_points is None else int(max_points)
    with open(path, "wb") as f:
        f.write(struct.pack("<4sIIII12x", MAGIC, VERSION, n_scans, mp, n_off))
        for k in range(n_scans):
            if not 0 < pts[k].shape[0] <= mp:
                raise ValueError("scan %d has %d points (max_points %d)" % (k, pts[k].shape[0], mp))
            f.write(struct.pack("<I3f", pts[k].shape[0], *[float(v) for v in odometry[k]]))
            f.write(pts[k].tobytes())
            if n_off:
                o = np.ascontiguousarray(offsets[k], dtype="<f4").reshape(-1, 3)
                if o.shape[0] != n_off:
                    raise ValueError("every scan needs %d offsets" % n_off)
                f.write(o.tobytes())


def read_scanlog(path) -> Tuple[List[np.ndarray], np.ndarray, Optional[List[np.ndarray]], int]:
    """-> (points per scan, odometry (n_scans, 3), offsets per scan or None, max_points)"""
    with open(path, "rb") as f:
        hd = f.read(32)
        if len(hd) != 32:
            raise ValueError("not a CSLG file")
        magic, version, n_scans, mp, n_off = struct.unpack("<4sIIII12x", hd)
        if magic != MAGIC or version != VERSION:
            raise ValueError("not a CSLG version 1 file")
        pts, odo, offs = [], np.zeros((n_scans, 3), dtype=np.float32), [] if n_off else None
        for k in range(n_scans):
            rec = f.read(16)
            if len(rec) != 16:
                raise ValueError("truncated record %d" % k)
            n = struct.unpack("<I", rec[:4])[0]
            if not 0 < n <= mp:
                raise ValueError("corrupt record %d" % k)
            odo[k] = np.frombuffer(rec[4:], dtype="<f4")
            raw = f.read(8 * n)
            if len(raw) != 8 * n:
                raise ValueError("truncated record %d" % k)
            pts.append(np.frombuffer(raw, dtype="<f4").reshape(n, 2).copy())
            if n_off:
                raw = f.read(12 * n_off)
                if len(raw) != 12 * n_off:
                    raise ValueError("truncated record %d" % k)
                offs.append(np.frombuffer(raw, dtype="<f4").reshape(n_off, 3).copy())
    return pts, odo, offs, mp

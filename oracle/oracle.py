"""ctypes binding of the CPU oracle (oracle/coreslam_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by slam.net_b200.  PARITY UNPINNED against an executed
reference (the reference is C#; no .NET runtime here) — see coreslam_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcoreslam_oracle.so")

INT32_MAX = 2147483647


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("coreslam_oracle.c", "coreslam_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _HoleMap(C.Structure):
    _fields_ = [("size", C.c_int), ("scale", C.c_float), ("pixels", C.POINTER(C.c_uint16))]


class _ObstacleMap(C.Structure):
    _fields_ = [("size", C.c_int), ("scale", C.c_float), ("pixels", C.POINTER(C.c_int8)), ("no_hit", C.POINTER(C.c_uint8))]


class _Processor(C.Structure):
    _fields_ = [
        ("map", C.POINTER(_HoleMap)),
        ("physical_map_size", C.c_float),
        ("start_pose", C.c_float * 3),
        ("sigma_xy", C.c_float),
        ("sigma_theta", C.c_float),
        ("iterations_per_thread", C.c_int),
        ("num_search_threads", C.c_int),
        ("quality", C.c_int),
        ("hole_width", C.c_float),
        ("position_search_beginning", C.c_int),
        ("pose", C.c_float * 3),
        ("last_odometry_pose", C.c_float * 3),
        ("scan_count", C.c_int),
        ("visits", C.c_int64),
        ("last_distance", C.c_int32),
        ("last_index", C.c_int32),
        ("omap", C.POINTER(_ObstacleMap)),
        ("unmapped_obstacle_hits", C.c_int),
        ("max_obstacle_hits", C.c_int),
        ("obstacle_visits", C.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int32)
    L.or_holemap_create.restype = C.POINTER(_HoleMap)
    L.or_holemap_create.argtypes = [C.c_int, C.c_float]
    L.or_holemap_destroy.argtypes = [C.POINTER(_HoleMap)]
    L.or_holemap_packed.argtypes = [C.POINTER(_HoleMap), C.c_void_p]
    L.or_cvt.restype = C.c_int32
    L.or_cvt.argtypes = [C.c_float]
    L.or_normalize_angle.restype = C.c_float
    L.or_normalize_angle.argtypes = [C.c_float]
    L.or_distance.restype = C.c_int32
    L.or_distance.argtypes = [C.POINTER(_HoleMap), fp, C.c_int, fp]
    L.or_clip_ray.restype = C.c_int
    L.or_clip_ray.argtypes = [C.c_int, ip, ip, C.c_int32, C.c_int32]
    L.or_draw_ray.restype = C.c_int64
    L.or_draw_ray.argtypes = [C.POINTER(_HoleMap)] + [C.c_int32] * 8 + [ip, C.c_int]
    L.or_update_hole_map.restype = C.c_int64
    L.or_update_hole_map.argtypes = [C.POINTER(_HoleMap), fp, C.c_int, fp, C.c_float, C.c_int, ip]
    L.or_monte_carlo_search.argtypes = [C.POINTER(_HoleMap), fp, C.c_int, fp, fp, C.c_int, fp, ip, ip]
    L.or_parallel_search.argtypes = [C.POINTER(_HoleMap), fp, C.c_int, fp, fp, C.c_int, C.c_int, fp, ip, ip, ip]
    L.or_segment_to_cloud.argtypes = [fp, C.c_int, fp, fp, fp]
    L.or_processor_create.restype = C.POINTER(_Processor)
    L.or_processor_create.argtypes = [C.c_float, C.c_int, fp, C.c_float, C.c_float, C.c_int, C.c_int]
    L.or_processor_create_full.restype = C.POINTER(_Processor)
    L.or_processor_create_full.argtypes = [C.c_float, C.c_int, C.c_int, fp, C.c_float, C.c_float, C.c_int, C.c_int]
    L.or_obstaclemap_create.restype = C.POINTER(_ObstacleMap)
    L.or_obstaclemap_create.argtypes = [C.c_int, C.c_float]
    L.or_obstaclemap_destroy.argtypes = [C.POINTER(_ObstacleMap)]
    L.or_draw_ray_obstacle.restype = C.c_int64
    L.or_draw_ray_obstacle.argtypes = [C.POINTER(_ObstacleMap)] + [C.c_int32] * 4 + [C.c_int]
    L.or_update_obstacle_map.restype = C.c_int64
    L.or_update_obstacle_map.argtypes = [C.POINTER(_ObstacleMap), fp, C.c_int, fp, C.c_int]
    L.or_processor_destroy.argtypes = [C.POINTER(_Processor)]
    L.or_processor_reset.argtypes = [C.POINTER(_Processor)]
    L.or_processor_update.argtypes = [C.POINTER(_Processor), fp, C.c_int, fp, fp]
    L.or_worker_create.restype = C.c_void_p
    L.or_worker_create.argtypes = [C.c_int]
    L.or_worker_destroy.argtypes = [C.c_void_p]
    L.or_parallel_search_mt.argtypes = [C.c_void_p, C.POINTER(_HoleMap), fp, C.c_int, fp, fp, C.c_int, fp, ip]
    L.or_processor_update_mt.argtypes = [C.POINTER(_Processor), C.c_void_p, fp, C.c_int, fp, fp]
    L.or_libm_sincos.argtypes = [fp, C.c_int64, fp, fp]
    L.or_normalize_angle_array.argtypes = [fp, C.c_int64, fp]
    L.or_crc32.restype = C.c_uint32
    L.or_crc32.argtypes = [C.c_void_p, C.c_uint64]
    _lib = L
    return L


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class HoleMap:
    """CoreSLAM/HoleMap.cs:17-37 — pixels is a numpy view of the C array (row-major y*size+x)."""

    def __init__(self, size_pixels: int, size_meters: float, _ptr=None):
        self._own = _ptr is None
        self._p = lib().or_holemap_create(size_pixels, float(size_meters)) if _ptr is None else _ptr
        self.size = self._p.contents.size
        self.scale = self._p.contents.scale
        self.pixels = np.ctypeslib.as_array(self._p.contents.pixels, shape=(self.size * self.size,))

    def fill(self, v: int):
        self.pixels[:] = v

    def crc32(self) -> int:
        return int(lib().or_crc32(self.pixels.ctypes.data, self.pixels.nbytes))

    def packed(self) -> np.ndarray:
        out = np.empty(self.size * self.size // 2, dtype=np.uint8)
        lib().or_holemap_packed(self._p, out.ctypes.data)
        return out

    def __del__(self):
        if getattr(self, "_own", False) and self._p:
            lib().or_holemap_destroy(self._p)
            self._p = None


class ObstacleMap:
    """CoreSLAM/ObstacleMap.cs:11-44 — pixels / no_hit are numpy views (size, size), first index Y."""

    def __init__(self, size_pixels: int, size_meters: float, _ptr=None):
        self._own = _ptr is None
        self._p = lib().or_obstaclemap_create(size_pixels, float(size_meters)) if _ptr is None else _ptr
        self.size = self._p.contents.size
        self.scale = self._p.contents.scale
        self.pixels = np.ctypeslib.as_array(self._p.contents.pixels, shape=(self.size, self.size))
        self.no_hit = np.ctypeslib.as_array(self._p.contents.no_hit, shape=(self.size, self.size))

    def fill(self, v: int):
        self.pixels[:] = v

    def __del__(self):
        if getattr(self, "_own", False) and self._p:
            lib().or_obstaclemap_destroy(self._p)
            self._p = None


def draw_ray_obstacle(m: ObstacleMap, x1, y1, x2, y2, max_obstacle_hits=10) -> int:
    return int(lib().or_draw_ray_obstacle(m._p, x1, y1, x2, y2, int(max_obstacle_hits)))


def update_obstacle_map(m: ObstacleMap, points, pose, max_obstacle_hits=10) -> int:
    pts, pp = _f(points)
    po, pop = _f(pose)
    return int(lib().or_update_obstacle_map(m._p, pp, pts.size // 2, pop, int(max_obstacle_hits)))


def cvt(f: float) -> int:
    return int(lib().or_cvt(float(np.float32(f))))


def normalize_angle(a: float) -> float:
    return float(lib().or_normalize_angle(float(np.float32(a))))


def libm_sincos(angles):
    a, ap = _f(angles)
    c = np.empty_like(a)
    s = np.empty_like(a)
    lib().or_libm_sincos(ap, a.size, c.ctypes.data_as(C.POINTER(C.c_float)), s.ctypes.data_as(C.POINTER(C.c_float)))
    return c, s


def normalize_angle_array(angles):
    a, ap = _f(angles)
    out = np.empty_like(a)
    lib().or_normalize_angle_array(ap, a.size, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def distance(m: HoleMap, points, pose) -> int:
    pts, pp = _f(points)
    po, pop = _f(pose)
    return int(lib().or_distance(m._p, pp, pts.size // 2, pop))


def clip_ray(size, xyc, yxc, xy, yx):
    a, b = C.c_int32(xyc), C.c_int32(yxc)
    ok = lib().or_clip_ray(size, C.byref(a), C.byref(b), xy, yx)
    return bool(ok), a.value, b.value


def draw_ray(m: HoleMap, x1, y1, x2, y2, xp, yp, value, alpha, trace=False):
    cap = 4 * m.size + 8 if trace else 0
    tr = np.zeros((max(cap, 1), 3), dtype=np.int32)
    n = int(lib().or_draw_ray(m._p, x1, y1, x2, y2, xp, yp, value, alpha, _i(tr) if trace else None, cap))
    return (n, tr[:n]) if trace else n


def update_hole_map(m: HoleMap, points, pose, hole_width, quality, rays=False):
    pts, pp = _f(points)
    po, pop = _f(pose)
    n = pts.size // 2
    r = np.zeros((max(n, 1), 6), dtype=np.int32)
    v = int(lib().or_update_hole_map(m._p, pp, n, pop, float(hole_width), int(quality), _i(r) if rays else None))
    return (v, r[:n]) if rays else v


def parallel_search(m: HoleMap, points, search_pose, offsets, iterations, threads):
    """Returns (best_pose[3], best_distance, distances[1+T*I], best_flat_index)."""
    pts, pp = _f(points)
    sp, spp = _f(search_pose)
    off, offp = _f(offsets)
    assert off.size == 3 * iterations * threads
    best = np.zeros(3, dtype=np.float32)
    bd = C.c_int32(0)
    bi = C.c_int32(0)
    d = np.zeros(1 + iterations * threads, dtype=np.int32)
    lib().or_parallel_search(m._p, pp, pts.size // 2, spp, offp, iterations, threads,
                             best.ctypes.data_as(C.POINTER(C.c_float)), C.byref(bd), _i(d), C.byref(bi))
    return best, bd.value, d, bi.value


def segment_to_cloud(rays, segment_pose, odometry_pose):
    r, rp = _f(rays)
    s, sp = _f(segment_pose)
    o, op = _f(odometry_pose)
    out = np.zeros(r.size, dtype=np.float32)
    lib().or_segment_to_cloud(rp, r.size // 2, sp, op, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out.reshape(-1, 2)


class Worker:
    """BaseSLAM/ParallelWorker.cs — persistent thread pool for the CPU baseline."""

    def __init__(self, num_threads: int):
        self.num_threads = num_threads
        self._w = lib().or_worker_create(num_threads)

    def parallel_search(self, m: HoleMap, points, search_pose, offsets, iterations):
        pts, pp = _f(points)
        sp, spp = _f(search_pose)
        off, offp = _f(offsets)
        best = np.zeros(3, dtype=np.float32)
        bd = C.c_int32(0)
        lib().or_parallel_search_mt(self._w, m._p, pp, pts.size // 2, spp, offp, iterations,
                                    best.ctypes.data_as(C.POINTER(C.c_float)), C.byref(bd))
        return best, bd.value

    def close(self):
        if self._w:
            lib().or_worker_destroy(self._w)
            self._w = None

    def __del__(self):
        self.close()


class Processor:
    """CoreSLAMProcessor state machine (ctor :119-162, Reset :167-175, Update :717-752)."""

    def __init__(self, physical_map_size, hole_map_size, start_pose, sigma_xy, sigma_theta,
                 iterations_per_thread, num_search_threads, obstacle_map_size=0):
        sp, spp = _f(start_pose)
        self._p = lib().or_processor_create_full(float(physical_map_size), int(hole_map_size), int(obstacle_map_size), spp,
                                                 float(sigma_xy), float(sigma_theta), int(iterations_per_thread),
                                                 int(num_search_threads))
        self.map = HoleMap(0, 0, _ptr=self._p.contents.map)
        self.obstacle_map = ObstacleMap(0, 0, _ptr=self._p.contents.omap) if obstacle_map_size > 0 else None

    unmapped_obstacle_hits = property(lambda s: s._p.contents.unmapped_obstacle_hits,
                                      lambda s, v: setattr(s._p.contents, "unmapped_obstacle_hits", int(v)))
    max_obstacle_hits = property(lambda s: s._p.contents.max_obstacle_hits,
                                 lambda s, v: setattr(s._p.contents, "max_obstacle_hits", int(v)))

    @property
    def obstacle_visits(self):
        return int(self._p.contents.obstacle_visits)

    quality = property(lambda s: s._p.contents.quality, lambda s, v: setattr(s._p.contents, "quality", int(v)))
    hole_width = property(lambda s: s._p.contents.hole_width,
                          lambda s, v: setattr(s._p.contents, "hole_width", float(v)))
    position_search_beginning = property(lambda s: s._p.contents.position_search_beginning,
                                         lambda s, v: setattr(s._p.contents, "position_search_beginning", int(v)))

    @property
    def pose(self):
        return np.array(list(self._p.contents.pose), dtype=np.float32)

    @property
    def visits(self):
        return int(self._p.contents.visits)

    @property
    def last_distance(self):
        return int(self._p.contents.last_distance)

    @property
    def last_index(self):
        return int(self._p.contents.last_index)

    @property
    def scan_count(self):
        return int(self._p.contents.scan_count)

    def reset(self):
        lib().or_processor_reset(self._p)

    def update(self, points, odometry_pose, offsets=None, worker: Worker | None = None):
        pts, pp = _f(points)
        od, odp = _f(odometry_pose)
        if offsets is None:
            offp = None
        else:
            off, offp = _f(offsets)
        if worker is None:
            lib().or_processor_update(self._p, pp, pts.size // 2, odp, offp)
        else:
            lib().or_processor_update_mt(self._p, worker._w, pp, pts.size // 2, odp, offp)
        return self.pose

    def __del__(self):
        if getattr(self, "_p", None):
            self.map = None
            self.obstacle_map = None
            lib().or_processor_destroy(self._p)
            self._p = None

"""Second, independent restatement of the CoreSLAM hot path — pure Python, statement by statement.

TEST INFRASTRUCTURE ONLY.  Its job is to catch transcription mistakes in coreslam_oracle.c: the two
were written separately from the C# and must agree bit-for-bit (tests/test_oracle_cross.py).
Floats are numpy float32 scalars (one rounding per operation), ints are wrapped to int32 by hand,
cosf/sinf/fmodf/sqrtf come from the host libm through ctypes — the same functions .NET 6 calls on
Linux x64.  Small inputs only (Python loops).

Citations: /root/reference/CoreSLAM/CoreSLAMProcessor.cs unless another file is named.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("cosf", "sinf", "sqrtf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_libm.fmodf.restype = ctypes.c_float
_libm.fmodf.argtypes = [ctypes.c_float, ctypes.c_float]

F = np.float32
TS_NO_OBSTACLE = 65500  # :21
TS_OBSTACLE = 0  # :22
INT_MAX = 2147483647
INT_MIN = -2147483648


def cosf(x):
    return F(_libm.cosf(float(x)))


def sinf(x):
    return F(_libm.sinf(float(x)))


def sqrtf(x):
    return F(_libm.sqrtf(float(x)))


def i32(v: int) -> int:
    """wrap a Python int to C# unchecked int"""
    v &= 0xFFFFFFFF
    return v - 0x100000000 if v & 0x80000000 else v


def to_int(f) -> int:
    """(int)f on .NET 6 x64: cvttss2si"""
    f = float(f)
    if math.isnan(f) or f >= 2147483648.0 or f < -2147483648.0:
        return INT_MIN
    return int(f)  # truncates toward zero


def cdiv(a: int, b: int) -> int:
    """C# integer division: truncation toward zero"""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def sign(v: int) -> int:
    return (v > 0) - (v < 0)


class HoleMapT:
    """CoreSLAM/HoleMap.cs:17-37"""

    def __init__(self, size_pixels: int, size_meters: float):
        self.Size = size_pixels
        self.Scale = F(size_pixels) / F(size_meters)  # :20 int -> float divide
        self.Pixels = np.zeros(size_pixels * size_pixels, dtype=np.uint16)


def normalize_angle(angle):
    """BaseSLAM/MathEx.cs:116-138"""
    with np.errstate(all="ignore"):
        pi = F(math.pi)  # MathF.PI
        pi2 = pi * F(2.0)
        a = F(_libm.fmodf(float(angle), float(pi2))) + pi2
        a = F(_libm.fmodf(float(a), float(pi2)))
        if a > pi:
            a = a - F(2.0) * pi
        return F(a)


def calculate_distance(hm: HoleMapT, points, pose) -> int:
    """:226-259"""
    with np.errstate(all="ignore"):
        nb_points = 0
        total = 0
        px = F(pose[0]) * hm.Scale + F(0.5)
        py = F(pose[1]) * hm.Scale + F(0.5)
        c = cosf(F(pose[2])) * hm.Scale
        s = sinf(F(pose[2])) * hm.Scale
        for X, Y in points:
            X = F(X)
            Y = F(Y)
            x = to_int(px + c * X - s * Y)
            y = to_int(py + s * X + c * Y)
            if 0 <= x < hm.Size and 0 <= y < hm.Size:
                total += int(hm.Pixels[y * hm.Size + x])
                nb_points += 1
        if nb_points > 0:
            return i32(cdiv(total * 1024, len(points)))
        return INT_MAX


def clip_ray(size: int, xyc: int, yxc: int, xy: int, yx: int):
    """:320-345 — returns (ok, xyc, yxc)"""
    if xyc < 0:
        if xyc == xy:
            return False, xyc, yxc
        yxc = i32(yxc + cdiv(i32(i32(yxc - yx) * i32(-xyc)), i32(xyc - xy)))
        xyc = 0
    if xyc >= size:
        if xyc == xy:
            return False, xyc, yxc
        yxc = i32(yxc + cdiv(i32(i32(yxc - yx) * i32(size - 1 - xyc)), i32(xyc - xy)))
        xyc = size - 1
    return True, xyc, yxc


def draw_laser_ray(hm: HoleMapT, x1, y1, x2, y2, xp, yp, value, alpha) -> int:
    """:359-443 — returns number of cells written"""
    x2c, y2c = x2, y2
    ok, x2c, y2c = clip_ray(hm.Size, x2c, y2c, x1, y1)
    if not ok:
        return 0
    ok, y2c, x2c = clip_ray(hm.Size, y2c, x2c, y1, x1)
    if not ok:
        return 0
    dx = abs(i32(x2 - x1))
    dy = abs(i32(y2 - y1))
    dxc = abs(i32(x2c - x1))
    dyc = abs(i32(y2c - y1))
    incptrx = sign(i32(x2 - x1))
    incptry = sign(i32(y2 - y1)) * hm.Size
    sincv = sign(value - TS_NO_OBSTACLE)
    if dx > dy:
        derrorv = abs(i32(xp - x2))
    else:
        dx = dy
        dxc, dyc = dyc, dxc
        incptrx, incptry = incptry, incptrx
        derrorv = abs(i32(yp - y2))
    if derrorv == 0:
        return 0
    error = i32(2 * dyc - dxc)
    horiz = i32(2 * dyc)
    diago = i32(2 * (dyc - dxc))
    errorv = cdiv(derrorv, 2)
    incv = cdiv(value - TS_NO_OBSTACLE, derrorv)
    incerrorv = i32(value - TS_NO_OBSTACLE - derrorv * incv)
    ptr = y1 * hm.Size + x1
    pixval = TS_NO_OBSTACLE
    visits = 0
    x = 0
    while x <= dxc:
        if x > i32(dx - 2 * derrorv):
            if x <= i32(dx - derrorv):
                pixval += incv
                errorv += incerrorv
                if errorv > derrorv:
                    pixval += sincv
                    errorv -= derrorv
            else:
                pixval -= incv
                errorv -= incerrorv
                if errorv < 0:
                    pixval -= sincv
                    errorv += derrorv
        hm.Pixels[ptr] = (i32((256 - alpha) * int(hm.Pixels[ptr]) + alpha * pixval) >> 8) & 0xFFFF
        visits += 1
        if error > 0:
            ptr += incptry
            error += diago
        else:
            error += horiz
        x += 1
        ptr += incptrx
    return visits


def update_hole_map(hm: HoleMapT, points, pose, hole_width, quality) -> int:
    """:496-534"""
    with np.errstate(all="ignore"):
        px = F(pose[0]) * hm.Scale + F(0.5)
        py = F(pose[1]) * hm.Scale + F(0.5)
        c = cosf(F(pose[2])) * hm.Scale
        s = sinf(F(pose[2])) * hm.Scale
        x1 = to_int(px)
        y1 = to_int(py)
        if x1 < 0 or x1 >= hm.Size or y1 < 0 or y1 >= hm.Size:
            return 0
        visits = 0
        for X, Y in points:
            X = F(X)
            Y = F(Y)
            x2p = c * X - s * Y
            y2p = s * X + c * Y
            xp = to_int(px + x2p)
            yp = to_int(py + y2p)
            dist = sqrtf(x2p * x2p + y2p * y2p)
            add = F(hole_width) * hm.Scale / F(2.0) / dist
            x2p = x2p * (F(1.0) + add)
            y2p = y2p * (F(1.0) + add)
            x2 = to_int(px + x2p)
            y2 = to_int(py + y2p)
            visits += draw_laser_ray(hm, x1, y1, x2, y2, xp, yp, TS_OBSTACLE, quality)
        return visits


class ObstacleMapT:
    """CoreSLAM/ObstacleMap.cs:11-44"""

    def __init__(self, size_pixels: int, size_meters: float):
        self.Size = size_pixels
        self.Scale = F(size_pixels) / F(size_meters)  # ObstacleMap.cs:20
        self.Pixels = np.zeros((size_pixels, size_pixels), dtype=np.int8)  # [y, x]


def draw_laser_ray_on_obstacle_map(om: ObstacleMapT, no_hit_map, x1, y1, x2, y2, max_obstacle_hits) -> int:
    """:456-490.  Returns how many map cells the loop touched; raises OverflowError where Math.Abs would throw."""
    if i32(x2 - x1) == INT_MIN or i32(y2 - y1) == INT_MIN:
        raise OverflowError("Math.Abs(int.MinValue)")
    dx = abs(i32(x2 - x1))
    sx = sign(i32(x2 - x1))
    dy = abs(i32(y2 - y1))
    sy = sign(i32(y2 - y1))
    err = cdiv(dx if dx > dy else -dy, 2)
    touched = 0
    while True:
        if x1 < 0 or x1 >= om.Size or y1 < 0 or y1 >= om.Size:
            break
        elif x1 == x2 and y1 == y2:
            if om.Pixels[y1, x1] < max_obstacle_hits:
                om.Pixels[y1, x1] += 1
            touched += 1
            break
        else:
            no_hit_map[y1, x1] = True
            touched += 1
        e2 = err
        if e2 > -dx:
            err = i32(err - dy)
            x1 = i32(x1 + sx)
        if e2 < dy:
            err = i32(err + dx)
            y1 = i32(y1 + sy)
    return touched


def update_obstacle_map(om: ObstacleMapT, points, pose, max_obstacle_hits) -> int:
    """:540-593"""
    with np.errstate(all="ignore"):
        no_hit_map = np.zeros((om.Size, om.Size), dtype=bool)
        px = F(pose[0]) * om.Scale + F(0.5)
        py = F(pose[1]) * om.Scale + F(0.5)
        c = cosf(F(pose[2])) * om.Scale
        s = sinf(F(pose[2])) * om.Scale
        x1 = to_int(px)
        y1 = to_int(py)
        if x1 < 0 or x1 >= om.Size or y1 < 0 or y1 >= om.Size:
            return 0
        touched = 0
        for X, Y in points:
            X = F(X)
            Y = F(Y)
            x2 = to_int(px + c * X - s * Y)
            y2 = to_int(py + s * X + c * Y)
            touched += draw_laser_ray_on_obstacle_map(om, no_hit_map, x1, y1, x2, y2, max_obstacle_hits)
        for y in range(om.Size):
            for x in range(om.Size):
                if no_hit_map[y, x]:
                    if om.Pixels[y, x] < 0:
                        om.Pixels[y, x] += 1
                    elif om.Pixels[y, x] > 0:
                        om.Pixels[y, x] -= 1
        return touched


def monte_carlo_search(hm, points, search_pose, offsets):
    """:624-653 — offsets: iterable of (dx, dy, dtheta) in dequeue order"""
    best_pose = tuple(F(v) for v in search_pose)
    current = calculate_distance(hm, points, search_pose)
    best = current
    for off in offsets:
        cur = (F(search_pose[0]) + F(off[0]), F(search_pose[1]) + F(off[1]), F(search_pose[2]) + F(off[2]))
        current = calculate_distance(hm, points, cur)
        if current < best:
            best = current
            best_pose = cur
    return best_pose, best


def parallel_monte_carlo_search(hm, points, search_pose, offsets, iterations, threads):
    """:674-710 with the worker bodies run in index order"""
    distances = []
    poses = []
    for t in range(threads):
        p, d = monte_carlo_search(hm, points, search_pose, offsets[t * iterations:(t + 1) * iterations])
        poses.append(p)
        distances.append(d)
    best_distance = INT_MAX
    best_pose = tuple(F(v) for v in search_pose)
    for t in range(threads):
        if distances[t] < best_distance:
            best_distance = distances[t]
            best_pose = poses[t]
    return best_pose, best_distance


def scan_segments_to_cloud(segments, odometry_pose):
    """:187-207.  segments: iterable of (pose (x, y, theta), rays [(angle, radius), ...]); returns the cloud's points as an
    (n, 2) float32 array in the order cloud.Points.Add (:203) builds them."""
    odo = tuple(F(v) for v in odometry_pose)
    points = []
    for seg_pose, rays in segments:  # :191
        pose = tuple(F(seg_pose[k]) - odo[k] for k in range(3))  # :194 Vector3 subtraction, per component
        for angle, radius in rays:  # :196
            a = F(angle) + pose[2]
            x = pose[0] + F(radius) * cosf(a)  # :200
            y = pose[1] + F(radius) * sinf(a)  # :201
            points.append((x, y))
    return np.array(points, dtype=np.float32).reshape(-1, 2)


class ProcessorT:
    """ctor :119-162, Reset :167-175, Update :717-752"""

    def __init__(self, physical_map_size, hole_map_size, start_pose, sigma_xy, sigma_theta, iterations, threads,
                 obstacle_map_size=0):
        self.HoleMap = HoleMapT(hole_map_size, physical_map_size)
        self.ObstacleMap = ObstacleMapT(obstacle_map_size, physical_map_size) if obstacle_map_size > 0 else None
        self.UnmappedObstacleHits = -5  # :98
        self.MaxObstacleHits = 10  # :103
        self.start_pose = tuple(F(v) for v in start_pose)
        self.iterations = iterations
        self.threads = threads
        self.Quality = 50
        self.HoleWidth = F(0.6)
        self.PositionSearchBeginning = 5
        self.Reset()

    def Reset(self):
        self.HoleMap.Pixels[:] = (TS_OBSTACLE + TS_NO_OBSTACLE) // 2
        if self.ObstacleMap is not None:
            self.ObstacleMap.Pixels[:] = self.UnmappedObstacleHits  # :170
        self.Pose = self.start_pose
        self.lastOdometryPose = (F(0), F(0), F(0))
        self.scanCount = 0

    def Update(self, points, odo_pose, offsets):
        odo = tuple(F(v) for v in odo_pose)
        if self.scanCount >= self.PositionSearchBeginning:
            search_pose = tuple(self.Pose[k] + (odo[k] - self.lastOdometryPose[k]) for k in range(3))
            if self.threads > 0:
                new_pose, _ = parallel_monte_carlo_search(self.HoleMap, points, search_pose, offsets,
                                                          self.iterations, self.threads)
            else:
                new_pose, _ = monte_carlo_search(self.HoleMap, points, search_pose, offsets[:self.iterations])
        else:
            self.scanCount += 1
            new_pose = odo
        self.lastOdometryPose = odo
        new_pose = (new_pose[0], new_pose[1], normalize_angle(new_pose[2]))
        self.Pose = new_pose
        update_hole_map(self.HoleMap, points, self.Pose, self.HoleWidth, self.Quality)
        if self.ObstacleMap is not None:
            update_obstacle_map(self.ObstacleMap, points, self.Pose, self.MaxObstacleHits)  # :751

    def UpdateSegments(self, segments, offsets):
        """Update(List<ScanSegment>) as declared (:717-723): odoPose = segments.Last().Pose, cloud from ScanSegmentsToCloud."""
        segments = list(segments)
        if not segments:
            raise ValueError("Sequence contains no elements")  # segments.Last() (:719)
        odo = segments[-1][0]
        self.Update(scan_segments_to_cloud(segments, odo), odo, offsets)

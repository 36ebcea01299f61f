"""Host-side model of the wedge decomposition used by cs_wedge_kernel (design validation, not product code).

Takes the per-ray integer tuples (x1,y1,x2,y2,xp,yp) the oracle derives from a scan, restates the closed forms of the
draw loop (cs_make_ray / cs_ray_minor / cs_ray_pixval) and checks the claims the kernel rests on:

  1. position of ray r on ring k:  p = c*k + g*m(k),  m(k) = min(k, ceil(k*s - 1/2)),  |p - k*kappa| <= 1/2, kappa = c + g*s
  2. wedge w owns positions ceil(k*beta_w) <= p < ceil(k*beta_{w+1}); a ray visits wedge w at some ring >= k0 only if
     beta_w - 1/(2 k0) <= kappa < beta_{w+1} + 1/(2 k0)   (plus the p = 8k -> 0 wrap into wedge 0)
  3. processing every (ring range, wedge) task independently, candidates in ray order, gives the oracle's map.
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc

NO_OBST = 65500


def clip(size, xyc, yxc, xy, yx):
    if xyc < 0:
        if xyc == xy:
            return None
        yxc += int((yxc - yx) * (-xyc) / (xyc - xy)) if False else _cdiv((yxc - yx) * (-xyc), (xyc - xy))
        xyc = 0
    if xyc >= size:
        if xyc == xy:
            return None
        yxc += _cdiv((yxc - yx) * (size - 1 - xyc), (xyc - xy))
        xyc = size - 1
    return xyc, yxc


def _cdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def make_ray(size, x1, y1, x2, y2, xp, yp):
    """dict of the closed-form parameters, or None for a ray that draws nothing"""
    r = clip(size, x2, y2, x1, y1)
    if r is None:
        return None
    x2c, y2c = r
    r = clip(size, y2c, x2c, y1, x1)
    if r is None:
        return None
    y2c, x2c = r
    dx, dy = abs(x2 - x1), abs(y2 - y1)
    dxc, dyc = abs(x2c - x1), abs(y2c - y1)
    sx, sy = np.sign(x2 - x1), np.sign(y2 - y1)
    if dx > dy:
        steep, D, smaj, smin = False, abs(xp - x2), sx, sy
    else:
        steep, dx = True, dy
        dxc, dyc = dyc, dxc
        D, smaj, smin = abs(yp - y2), sy, sx
    if D == 0 or dxc >= size or dyc >= size:
        return None
    incv = -(NO_OBST // D)            # (0 - 65500) / D truncated toward zero
    rem = NO_OBST - D * (NO_OBST // D)
    t2 = dx - 2 * D + 1
    t1 = dx - D
    a0 = max(t2, 0)
    lim = size + 1
    a0 = min(a0, lim); t1 = min(t1, lim); t1 = max(t1, -1)
    nd_total = max(t1 - a0 + 1, 0)
    e0 = D // 2 - nd_total * rem
    need = -e0 - rem
    kc = 0
    if need > 0:
        kc = (need + rem + D - 1) // (rem + D)
    kc = min(kc, lim)
    return dict(dxc=dxc, dyc=dyc, a0=a0, b0=t1, incv=incv, kc=kc, nd_total=nd_total, steep=steep, majneg=smaj < 0, minneg=smin < 0)


def pixval(r, k):
    if k <= r["b0"]:
        nd = max(k - r["a0"] + 1, 0)
        return NO_OBST + nd * r["incv"]
    c0 = max(r["a0"], r["b0"] + 1)
    if k < c0:
        return NO_OBST
    j = k - c0 + 1
    return NO_OBST + (r["nd_total"] - j) * r["incv"] + min(j, r["kc"])


def minor(r, k):
    if k == 0:
        return 0
    return min(k, (2 * r["dyc"] * k + r["dxc"] - 1) // (2 * r["dxc"]))


def side_c_g(r):
    """p = c*k + g*m"""
    if not r["steep"]:
        if r["majneg"]:
            return 5, (+1 if r["minneg"] else -1)   # p = 5k - oy, oy = -m if minneg
        return 1, (-1 if r["minneg"] else +1)       # p = k + oy
    if r["majneg"]:
        return 7, (-1 if r["minneg"] else +1)       # p = 7k + ox
    return 3, (+1 if r["minneg"] else -1)           # p = 3k - ox


def cell_of(r, k, x1, y1):
    m = minor(r, k)
    dmaj = -k if r["majneg"] else k
    dmin = -m if r["minneg"] else m
    ox, oy = (dmin, dmaj) if r["steep"] else (dmaj, dmin)
    return x1 + ox, y1 + oy


def kappa(r):
    c, g = side_c_g(r)
    s = min(r["dyc"], r["dxc"]) / r["dxc"] if r["dxc"] > 0 else 0.0
    return c + g * s


def blend(v, pv, alpha):
    return (((256 - alpha) * v + alpha * pv) >> 8) & 0xFFFF


FIX = 14


def bound(k, beta_fix):
    return (k * beta_fix + (1 << FIX) - 1) >> FIX


def run_model(size, rays6, alpha, pixels, levels, wedges_of_level, verbose=False):
    """pixels: flat row-major u16 list/array (modified in place).  levels: list of (k0, k1) inclusive ring ranges covering
    1..max_ring.  wedges_of_level(level_index, k0) -> W.  Returns stats."""
    x1, y1 = rays6[0][0], rays6[0][1]
    rays = [make_ray(size, *t) for t in rays6]
    keys = [kappa(r) if r else None for r in rays]
    max_ring = max([r["dxc"] for r in rays if r] + [-1])
    stats = dict(tasks=0, cand=0, visits=0, max_cand=0, multi=0)
    # ring 0: the start cell, every valid ray in order
    for r in rays:
        if r:
            pixels[y1 * size + x1] = blend(pixels[y1 * size + x1], pixval(r, 0), alpha)
            stats["visits"] += 1
    covered = 0
    for li, (k0, k1) in enumerate(levels):
        if k0 > max_ring:
            break
        k1 = min(k1, max_ring)
        W = wedges_of_level(li, k0)
        eps = 0.5 / k0 + 1e-4
        for w in range(W):
            blo = (8 << FIX) * w // W
            bhi = (8 << FIX) * (w + 1) // W
            flo, fhi = blo / (1 << FIX), bhi / (1 << FIX)
            cand = [i for i, r in enumerate(rays) if r and r["dxc"] >= k0 and
                    ((flo - eps <= keys[i] < fhi + eps) or (w == 0 and keys[i] >= 8 - eps))]
            stats["tasks"] += 1
            stats["cand"] += len(cand)
            stats["max_cand"] = max(stats["max_cand"], len(cand))
            stats["multi"] += len(cand) > 32
            for k in range(k0, k1 + 1):
                lo, hi = bound(k, blo), bound(k, bhi)
                for i in cand:
                    r = rays[i]
                    if k > r["dxc"]:
                        continue
                    c, g = side_c_g(r)
                    p = c * k + g * minor(r, k)
                    assert abs(p - k * keys[i]) <= 0.5 + 1e-9, (p, k, keys[i])
                    if p == 8 * k:
                        p = 0
                    if lo <= p < hi:
                        x, y = cell_of(r, k, x1, y1)
                        pixels[y * size + x] = blend(pixels[y * size + x], pixval(r, k), alpha)
                        covered += 1
    stats["visits"] += covered
    # completeness: every visit of every ray was claimed by exactly one task
    total = sum(min(r["dxc"], max(k1 for _, k1 in levels)) for r in rays if r)
    stats["expected_ring_visits"] = total
    assert covered == total, (covered, total)
    return stats


def default_levels(max_ring):
    lv = [(1, 1), (2, 3), (4, 7), (8, 15), (16, 31), (32, 63)]
    k = 64
    while k <= max_ring:
        lv.append((k, k + 63))
        k += 64
    return lv


if __name__ == "__main__":
    from slam.net_b200 import synth
    n_pts = int(sys.argv[1]) if len(sys.argv) > 1 else 360
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    phys = 40.0
    rp = synth.make_replay(12, n_pts, phys)
    m = orc.HoleMap(size, phys)
    m.fill(32750)
    mine = np.array(m.pixels).astype(np.int64).tolist()
    for k in range(6):
        pose = rp.odometry[k].copy()
        pts = rp.points[k]
        if k == 4:
            pts = pts[np.random.default_rng(3).permutation(len(pts))]  # any ray order
        if k == 5:
            pts = np.concatenate([pts, pts[::2] * 0.9], axis=0)  # more than one turn
        v, r6 = orc.update_hole_map(m, pts, pose, 0.6, 50, rays=True)
        r6 = [tuple(int(x) for x in t) for t in r6]
        n_alive = len(r6)
        st = run_model(size, r6, 50, mine, default_levels(size), lambda li, k0: max(1, min(n_alive // 24, 8 * k0)))
        ok = mine == np.array(m.pixels).astype(np.int64).tolist()
        print("scan", k, "visits", v, st, "map equal:", ok)
        assert ok and st["visits"] == v

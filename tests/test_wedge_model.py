"""The decomposition cs_wedge_kernel rests on, checked on the CPU against the oracle with the host model in
tools/wedge_model.py: on ring k a ray sits at position p = c k + g m(k) with |p - k kappa| <= 1/2; the wedges of a level
partition every ring's positions exactly; a task that filters its candidates by key +- 1/(2 k0) misses no visit; and drawing
the (ring range x wedge) tasks independently, each with its candidates in ray order, gives the oracle's map — for any ray
order, any number of turns, and any cut of the circle into wedges (the kernel's balance rule is only one of them)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import wedge_model as wm  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from slam.net_b200 import synth  # noqa: E402


def kernel_wedges(alive, k0):
    """cs_w_wedges (cs_wedge.cuh): ~26 candidates per wedge, of which alive / (8 k0) come from the margins."""
    if alive <= 0:
        return 0
    halo2 = alive // (8 * k0)
    own = halo2 if halo2 > 13 else 26 - halo2
    return max(1, min((alive + own - 1) // own, 8 * k0, 8192))


CUTS = {
    "kernel rule": lambda n: (lambda li, k0: kernel_wedges(n, k0)),
    "one wedge": lambda n: (lambda li, k0: 1),
    "a wedge per cell of the first ring": lambda n: (lambda li, k0: 8 * k0),
    "prime count": lambda n: (lambda li, k0: min(7, 8 * k0)),
}


@pytest.mark.parametrize("cut", sorted(CUTS))
def test_wedge_tasks_reproduce_the_oracle_map(cut):
    n_pts, size, phys = 150, 256, 40.0
    rp = synth.make_replay(8, n_pts, phys, seed=11)
    m = orc.HoleMap(size, phys)
    m.fill(32750)
    mine = np.array(m.pixels).astype(np.int64).tolist()
    rng = np.random.default_rng(5)
    for k in range(6):
        pts = rp.points[k]
        if k == 3:
            pts = pts[rng.permutation(len(pts))]                 # any ray order
        if k == 4:
            pts = np.concatenate([pts, pts[::2] * 0.9], axis=0)  # more than one turn: rays that overlap in angle
        if k == 5:
            pts = pts[::-1].copy()                               # reversed
        visits, r6 = orc.update_hole_map(m, pts, rp.odometry[k].copy(), 0.6, 50 + 30 * (k % 3), rays=True)
        r6 = [tuple(int(x) for x in t) for t in r6]
        st = wm.run_model(size, r6, 50 + 30 * (k % 3), mine, wm.default_levels(size), CUTS[cut](len(r6)))
        assert st["visits"] == visits                             # every visit claimed by exactly one task (asserted inside too)
        assert mine == np.array(m.pixels).astype(np.int64).tolist(), (cut, k)


def test_ring_position_closed_form_and_key_bound():
    """m(k) = min(k, ceil(k s - 1/2)) is the Bresenham walk of the reference (:394-396, :433-441), and the position it gives
    stays within half a cell of k * kappa — the bound the candidate filter's margins are derived from."""
    rng = np.random.default_rng(9)
    for _ in range(300):
        dxc = int(rng.integers(1, 700))
        dyc = int(rng.integers(0, dxc + 1))
        r = dict(dxc=dxc, dyc=dyc, steep=bool(rng.integers(2)), majneg=bool(rng.integers(2)), minneg=bool(rng.integers(2)))
        c, g = wm.side_c_g(r)
        kap = wm.kappa(r)
        # the reference's walk (:394-396, :433-441): error = 2 dyc - dxc; after the cell of step x is written, error > 0 takes
        # a minor step and adds 2 (dyc - dxc), otherwise it adds 2 dyc
        error, minor = 2 * dyc - dxc, 0
        for k in range(1, dxc + 1):
            if error > 0:
                minor += 1
                error += 2 * (dyc - dxc)
            else:
                error += 2 * dyc
            assert wm.minor(r, k) == minor, (dxc, dyc, k)
            p = c * k + g * minor
            assert abs(p - k * kap) <= 0.5 + 1e-9

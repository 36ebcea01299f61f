"""Seeded synthetic lidar workloads.

The reference has no recorded scan logs: its only workload is the WPF simulator, which ray-casts a
fixed room with Box2D (Simulation/Field.cs:45-69, Simulation/MainWindow.xaml.cs:380-407).  This module
rebuilds that room analytically (ray / segment intersection instead of Box2D) and scripts a closed-loop
trajectory through it, so the same inputs can be fed to the CUDA path and to the CPU oracle.
Host-side numpy only; nothing here is on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

# Simulation/Field.cs:45-59 (outer wall) and :63-69 (inner obstacle), unit coordinates
OUTER = np.array([(0.00, 0.0), (1.00, 0.0), (1.00, 0.2), (0.80, 0.3), (0.80, 0.5), (1.00, 0.4), (1.00, 1.0),
                  (0.6, 1.0), (0.6, 0.8), (0.5, 0.8), (0.5, 1.0), (0.0, 1.0)], dtype=np.float64)
INNER = np.array([(0.2, 0.3), (0.3, 0.3), (0.4, 0.7), (0.3, 0.7)], dtype=np.float64)


class Room:
    """The simulator's field scaled into a square map of `physical_size` metres: the simulator uses
    scale 30 m and offset (5, 5) in a 40 m map (MainWindow.xaml.cs:97, :69) — the same 3/4 and 1/8 here."""

    def __init__(self, physical_size: float = 40.0):
        self.physical_size = float(physical_size)
        scale, off = 0.75 * physical_size, 0.125 * physical_size
        segs = []
        for poly in (OUTER, INNER):
            p = poly * scale + off
            for i in range(len(p)):
                segs.append((p[i], p[(i + 1) % len(p)]))  # closed loops (Field.cs:100-112)
        self.a = np.array([s[0] for s in segs])
        self.b = np.array([s[1] for s in segs])
        self.scale, self.offset = scale, off

    def cast(self, pose, angles: np.ndarray, max_range: float = 40.0) -> np.ndarray:
        """Distance to the nearest wall along each lidar angle (lidar frame), inf when nothing is hit."""
        th = angles.astype(np.float64) + float(pose[2])
        d = np.stack([np.cos(th), np.sin(th)], axis=1)  # (N,2)
        o = np.array([float(pose[0]), float(pose[1])])
        e = self.b - self.a  # (M,2)
        w = self.a - o       # (M,2)
        den = d[:, None, 0] * e[None, :, 1] - d[:, None, 1] * e[None, :, 0]  # cross(d, e)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (w[None, :, 0] * e[None, :, 1] - w[None, :, 1] * e[None, :, 0]) / den
            u = (w[None, :, 0] * d[:, None, 1] - w[None, :, 1] * d[:, None, 0]) / den
        ok = (np.abs(den) > 1e-12) & (t > 1e-9) & (u >= 0.0) & (u <= 1.0)
        t = np.where(ok, t, np.inf)
        r = t.min(axis=1)
        return np.where(r <= max_range, r, np.inf)

    def trajectory(self, n: int, step: float = 0.05) -> np.ndarray:
        """Closed elliptical loop in the free space right of the inner obstacle, <= `step` metres and
        < 2 degrees of heading per scan; heading follows the tangent."""
        cx, cy = 0.62 * self.scale + self.offset, 0.5 * self.scale + self.offset
        rx, ry = 0.12 * self.scale, 0.2 * self.scale
        # arc-length parametrisation by dense sampling
        m = 20000
        phi = np.linspace(0.0, 2 * np.pi, m, endpoint=False)
        xs, ys = cx + rx * np.cos(phi), cy + ry * np.sin(phi)
        seg = np.hypot(np.diff(xs, append=xs[0]), np.diff(ys, append=ys[0]))
        cum = np.concatenate([[0.0], np.cumsum(seg)])
        total = cum[-1]
        s = (np.arange(n) * step) % total
        idx = np.searchsorted(cum, s, side="right") - 1
        idx = np.clip(idx, 0, m - 1)
        x, y = xs[idx], ys[idx]
        heading = np.arctan2(ry * np.cos(phi[idx]), -rx * np.sin(phi[idx]))
        return np.stack([x, y, heading], axis=1)


@dataclass
class Replay:
    physical_size: float
    truth: np.ndarray            # (n,3) float64
    odometry: np.ndarray         # (n,3) float32 — what Update receives as segment pose
    points: List[np.ndarray]     # per scan (P,2) float32, lidar frame relative to the odometry pose
    n_points: int

    def __len__(self):
        return len(self.points)


def lidar_scan(room: Room, pose, n_rays: int, rng: np.random.Generator, noise: bool = True) -> np.ndarray:
    """One 360-degree scan -> (P,2) float32 points in the lidar frame (ScanSegmentsToCloud with a single
    segment whose pose equals the odometry pose, CoreSLAMProcessor.cs:194-204).  Range noise follows
    MainWindow.xaml.cs:397: k/100 * 0.02 m, k in {-100..99}; rays with no hit are dropped (:395-400)."""
    ang = (np.arange(n_rays, dtype=np.float64) * (2.0 * np.pi / n_rays)).astype(np.float32)
    r = room.cast(pose, ang)
    if noise:
        r = r + rng.integers(-100, 100, n_rays) / 100.0 * 0.02
    keep = np.isfinite(r)
    r = np.maximum(r[keep], 1e-3).astype(np.float32)  # radius > 0 (a zero range has no direction to extend)
    a = ang[keep]
    return np.stack([r * np.cos(a), r * np.sin(a)], axis=1).astype(np.float32)


def make_replay(n_scans: int, n_points: int, physical_size: float = 40.0, seed: int = 0x5EED0000,
                step: float = 0.05, drift_sigma: float = 0.002, noise: bool = True) -> Replay:
    """A scripted drive through the room: truth trajectory, odometry = truth + slow random-walk drift,
    one n_points-ray scan per pose."""
    rng = np.random.default_rng(seed)
    room = Room(physical_size)
    truth = room.trajectory(n_scans, step)
    drift = np.cumsum(rng.normal(0.0, drift_sigma, (n_scans, 3)) * np.array([1.0, 1.0, 0.2]), axis=0)
    drift[0] = 0.0
    odo = (truth + drift).astype(np.float32)
    pts = [lidar_scan(room, truth[k], n_points, rng, noise) for k in range(n_scans)]
    return Replay(physical_size, truth, odo, pts, n_points)


def candidate_offsets(seed: int, scan_index: int, n: int, sigma_xy: float, sigma_theta: float) -> np.ndarray:
    """Verification-mode candidate table for one scan: N(0, sigma) deviates in the reference's dequeue
    order (X, Y, Theta), from a host PCG stream seeded per scan."""
    rng = np.random.default_rng([seed & 0xFFFFFFFF, scan_index])
    off = np.empty((n, 3), dtype=np.float32)
    off[:, 0:2] = rng.normal(0.0, sigma_xy, (n, 2))
    off[:, 2] = rng.normal(0.0, sigma_theta, n)
    return off


def random_map(size: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 65536, size * size, dtype=np.uint16)

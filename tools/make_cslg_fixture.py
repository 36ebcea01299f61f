#!/usr/bin/env python3
"""Writes a CSLG scan log (include/coreslam_b200.h) with candidate tables plus its parameter side-car, the input of
dotnet/ReplayHarness.cs and of tests/test_golden_reference.py.

    python tools/make_cslg_fixture.py                      # the small committed fixture tests/golden/cfg2_mini.cslg
    python tools/make_cslg_fixture.py --cfg2 out.cslg      # BASELINE configs[1] in full: 1000 scans x 1024 points x 4096 candidates (57 MB)

Host-side numpy only (slam.net_b200.scanlog_file + synth): no GPU, no oracle."""
import argparse, json, os, sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam.net_b200 import scanlog_file, synth  # noqa: E402

SIGMA_XY, SIGMA_THETA = 0.1, 0.17453292


def write(path, n_scans, points, size, phys, iters, threads, seed):
    rp = synth.make_replay(n_scans, points, phys, seed=seed)
    n_cand = iters * threads
    offs = [synth.candidate_offsets(seed, k, n_cand, SIGMA_XY, SIGMA_THETA) for k in range(n_scans)]
    scanlog_file.write_scanlog(path, rp.points, rp.odometry, offs, max_points=points)
    json.dump({"physical_map_size": phys, "hole_map_size": size, "obstacle_map_size": 64, "iterations_per_thread": iters,
               "num_search_threads": threads, "sigma_xy": SIGMA_XY, "sigma_theta": SIGMA_THETA, "scans": n_scans, "points": points,
               "seed": seed, "note": "each scan is fed as ONE segment at the odometry pose with rays (atan2f(y, x), |p|)"},
              open(path + ".json", "w"), indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg2", metavar="PATH", help="write the full cfg2 replay (1000 scans) to PATH instead of the mini fixture")
    a = ap.parse_args()
    if a.cfg2:
        write(a.cfg2, 1000, 1024, 2048, 40.0, 1024, 4, 0x5EED0000)
    else:
        write(os.path.join(ROOT, "tests", "golden", "cfg2_mini.cslg"), 24, 180, 512, 40.0, 100, 4, 0x5EED0001)

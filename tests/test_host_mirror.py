"""The C++ host mirror (slam.net_b200/csrc/host/coreslam.hpp) used the way the reference's simulator
uses CoreSLAMProcessor; its printed poses and map checksum are checked against the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

import slam.net_b200 as sn
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "slam.net_b200", "_build", "coreslam_demo")


def _build_demo():
    sn.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "slam.net_b200", "csrc", "host"), "-s"])


def hash01(a, b):
    a = np.asarray(a, dtype=np.uint32)
    b = np.uint32(b)
    with np.errstate(over="ignore"):
        h = (a * np.uint32(2654435761)) ^ (b + np.uint32(0x9E3779B9) + (a << np.uint32(6)) + (a >> np.uint32(2)))
        h = h ^ (h >> np.uint32(15))
        h = h * np.uint32(0x85EBCA6B)
        h = h ^ (h >> np.uint32(13))
    return (h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def test_demo_builds_and_fails_loudly_without_gpu():
    _build_demo()
    if sn.lib().cs_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([DEMO, "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle():
    _build_demo()
    scans, rays, T, I = 12, 180, 2, 64
    out = subprocess.run([DEMO, str(scans), str(rays), str(T), str(I)], capture_output=True, text=True, check=True).stdout
    got = re.findall(r"scan (\d+) pose (\S+) (\S+) (\S+) distance (-?\d+) index (\d+)", out)
    assert len(got) == scans
    o = orc.Processor(16.0, 256, [8.0, 8.0, 0.0], 0.1, 0.1, I, T)
    o.hole_width = 1.0
    ang = np.arange(rays, dtype=np.float32) * (np.float32(6.2831855) / np.float32(rays))
    rad = np.float32(3.0) + hash01(np.arange(rays), 7)
    for k in range(scans):
        pose = np.array([8.0 + 0.03125 * k, 8.0 - 0.015625 * k, 0.0078125 * k], dtype=np.float32)
        j = np.arange(T * I * 3)
        scale = np.where(j % 3 == 2, np.float32(0.125), np.float32(0.25)).astype(np.float32)
        off = ((hash01(j, k + 100) - np.float32(0.5)) * scale).astype(np.float32).reshape(-1, 3)
        cloud = orc.segment_to_cloud(np.stack([ang, rad], axis=1), pose, pose)
        o.update(cloud, pose, off)
        p = np.array([float.fromhex(v) for v in got[k][1:4]], dtype=np.float32)
        assert np.array_equal(p, o.pose), "scan %d" % k
        if k >= 5:
            assert int(got[k][4]) == o.last_distance and int(got[k][5]) == o.last_index
    checksum = int(re.search(r"map checksum (\d+)", out).group(1))
    assert checksum == sn.host_map_checksum(np.array(o.map.pixels), 256)

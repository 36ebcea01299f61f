"""The executed-reference pin: golden files written by dotnet/ReplayHarness.cs — the
REFERENCE mikkleini/slam.net CoreSLAMProcessor (with dotnet/reference_verification_hook.patch) replaying a CSLG log — are
checked here against the CPU oracle and against the CUDA path.  No .NET runtime exists in this image, so no
tests/golden/*.cslg.golden is committed yet: those tests skip, and the day a maintainer drops one in they run.  What always
runs: the committed fixture replays through the oracle, a golden file written FROM THE ORACLE round-trips through the same
parser and comparison (so the day-one path is exercised), and on a GPU the CUDA path is compared with that same file."""
import ctypes
import glob
import json
import os
import zlib

import numpy as np
import pytest

import slam.net_b200 as sn
from slam.net_b200 import scanlog_file
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.cslg")))
_libm = ctypes.CDLL("libm.so.6")
_libm.atan2f.restype = ctypes.c_float
_libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]


def as_segment_rays(points):
    """The harness feeds every scan as one segment at the odometry pose with rays (MathF.Atan2(y, x), Vector2.Length()):
    the platform's atan2f and a float sqrt of the float dot product."""
    p = np.asarray(points, dtype=np.float32)
    ang = np.array([_libm.atan2f(float(y), float(x)) for x, y in p], dtype=np.float32)
    rad = np.sqrt((p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]).astype(np.float32)).astype(np.float32)
    return np.stack([ang, rad], axis=1).astype(np.float32)


def bits(f):
    return "%08x" % int(np.float32(f).view(np.uint32))


def crc_of(pixels):
    return "%08x" % (zlib.crc32(np.ascontiguousarray(pixels, dtype="<u2").tobytes()) & 0xFFFFFFFF)


def parse_golden(path):
    rows = []
    for line in open(path):
        if line.startswith("#") or not line.strip():
            continue
        k, px, py, pz, dist, crc = line.split()
        rows.append((int(k), px.lower(), py.lower(), pz.lower(), int(dist), crc.lower()))
    return rows


def replay_oracle(log_path):
    cfg = json.load(open(log_path + ".json"))
    pts, odo, offs, mp = scanlog_file.read_scanlog(log_path)
    o = orc.Processor(cfg["physical_map_size"], cfg["hole_map_size"], odo[0], cfg["sigma_xy"], cfg["sigma_theta"],
                      cfg["iterations_per_thread"], cfg["num_search_threads"])
    rows = []
    for k in range(len(pts)):
        cloud = orc.segment_to_cloud(as_segment_rays(pts[k]), odo[k], odo[k])
        o.update(cloud, odo[k], offs[k])
        rows.append((k, bits(o.pose[0]), bits(o.pose[1]), bits(o.pose[2]), o.last_distance if k >= 5 else 2147483647,
                     crc_of(np.array(o.map.pixels))))
    return rows


def replay_gpu(log_path):
    cfg = json.load(open(log_path + ".json"))
    pts, odo, offs, mp = scanlog_file.read_scanlog(log_path)
    p = sn.Processor(cfg["physical_map_size"], cfg["hole_map_size"], odo[0], cfg["sigma_xy"], cfg["sigma_theta"],
                     cfg["iterations_per_thread"], cfg["num_search_threads"], max_points=mp)
    rows = []
    for k in range(len(pts)):
        seg = sn.ScanSegment(Rays=as_segment_rays(pts[k]), Pose=odo[k], IsLast=True)
        r = p.update_segments([seg], offs[k])
        rows.append((k, bits(r.pose[0]), bits(r.pose[1]), bits(r.pose[2]), r.distance if r.searched else 2147483647,
                     crc_of(p.map_download())))
    p.close()
    return rows


def compare(got, want, what):
    assert len(got) == len(want), what
    for g, w in zip(got, want):
        assert g[:4] == w[:4], "%s: pose differs at scan %d: %s vs %s" % (what, g[0], g[1:4], w[1:4])
        if g[0] >= 5:
            assert g[4] == w[4], "%s: winning distance differs at scan %d" % (what, g[0])
        assert g[5] == w[5], "%s: HoleMap CRC differs at scan %d" % (what, g[0])


def write_golden(path, rows):
    with open(path, "w") as f:
        f.write("# scan pose_x pose_y pose_theta (float bits, hex) distance holemap_crc32\n")
        for r in rows:
            f.write("%d %s %s %s %d %s\n" % r)


def test_fixture_is_committed_with_its_parameters():
    assert FIXTURES, "tests/golden/*.cslg missing: python tools/make_cslg_fixture.py"
    for f in FIXTURES:
        cfg = json.load(open(f + ".json"))
        pts, odo, offs, mp = scanlog_file.read_scanlog(f)
        assert offs is not None and offs[0].shape[0] == cfg["iterations_per_thread"] * cfg["num_search_threads"]
        assert len(pts) == cfg["scans"]


@pytest.mark.parametrize("log", FIXTURES, ids=[os.path.basename(f) for f in FIXTURES])
def test_oracle_golden_roundtrip(log, tmp_path):
    """A golden file written from the oracle parses back and compares equal: the path a reference-written file will take."""
    rows = replay_oracle(log)
    g = str(tmp_path / "oracle.golden")
    write_golden(g, rows)
    compare(replay_oracle(log), parse_golden(g), "oracle vs its own golden")
    assert len(set(r[5] for r in rows)) == len(rows)  # the map really changes every scan


@pytest.mark.parametrize("log", FIXTURES, ids=[os.path.basename(f) for f in FIXTURES])
def test_oracle_equals_reference_golden(log):
    golden = log + ".golden"
    if not os.path.exists(golden):
        pytest.skip("no executed-reference golden file (%s): needs a .NET box, see dotnet/ReplayHarness.cs" % os.path.basename(golden))
    compare(replay_oracle(log), parse_golden(golden), "CPU oracle vs the reference's golden file")


@pytest.mark.gpu
@pytest.mark.parametrize("log", FIXTURES, ids=[os.path.basename(f) for f in FIXTURES])
def test_gpu_equals_reference_golden_or_oracle(log, tmp_path):
    """The CUDA path through cs_update_segments (one segment, polar rays, candidate table) against the reference's golden
    file when there is one, else against the same file written from the oracle."""
    golden = log + ".golden"
    if os.path.exists(golden):
        compare(replay_gpu(log), parse_golden(golden), "CUDA path vs the reference's golden file")
    else:
        g = str(tmp_path / "oracle.golden")
        write_golden(g, replay_oracle(log))
        compare(replay_gpu(log), parse_golden(g), "CUDA path vs the oracle's golden file")

"""cs_cosf/cs_sinf (slam.net_b200/csrc/cs_math.h) must equal the host libm bit for bit — that is what
lets the device compute MathF.Cos/MathF.Sin itself (CoreSLAMProcessor.cs:234-235, 501-502)."""
import numpy as np
import pytest

import slam.net_b200 as sn
from oracle import oracle as orc


def _bits(stride, offset=0):
    return np.arange(offset, 2 ** 32, stride, dtype=np.uint64).astype(np.uint32)


def _compare(c1, s1, c2, s2):
    for a, b in ((c1, c2), (s1, s2)):
        both_nan = np.isnan(a) & np.isnan(b)
        ok = both_nan | (a.view(np.uint32) == b.view(np.uint32))
        assert ok.all(), "first mismatch at %d" % int(np.argmin(ok))


def _host(angles):
    c = np.empty_like(angles)
    s = np.empty_like(angles)
    import ctypes as C
    fp = C.POINTER(C.c_float)
    sn.lib().cs_host_sincos(angles.ctypes.data_as(fp), angles.size, c.ctypes.data_as(fp), s.ctypes.data_as(fp))
    return c, s


def test_host_sincos_matches_libm_dense_sample():
    ang = _bits(509).view(np.float32)  # ~8.4M floats across all exponents, both signs, inf/NaN included
    c1, s1 = _host(ang)
    c2, s2 = orc.libm_sincos(ang)
    _compare(c1, s1, c2, s2)


def test_host_sincos_pose_range_exhaustive():
    # every float in [-4, 4]: the range poses actually live in
    lo, hi = np.float32(0.0).view(np.uint32), np.float32(4.0).view(np.uint32)
    pos = np.arange(lo, hi + 1, 5, dtype=np.uint32).view(np.float32)
    ang = np.concatenate([pos, -pos])
    c1, s1 = _host(ang)
    c2, s2 = orc.libm_sincos(ang)
    _compare(c1, s1, c2, s2)


def test_normalize_angle_matches_libm_fmodf():
    rng = np.random.default_rng(0)
    ang = np.concatenate([rng.uniform(-50, 50, 200000), rng.uniform(-1e6, 1e6, 50000),
                          [0.0, -0.0, np.pi, -np.pi, 2 * np.pi, -2 * np.pi, 1e-30, -1e-30, 3.1415927, 3.1415925]]).astype(np.float32)
    want = orc.normalize_angle_array(ang)
    L = sn.lib()
    got = np.array([L.cs_host_normalize_angle(float(a)) for a in ang[:20000]], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), want[:20000].view(np.uint32))
    tail = np.array([L.cs_host_normalize_angle(float(a)) for a in ang[-10:]], dtype=np.float32)
    assert np.array_equal(tail.view(np.uint32), want[-10:].view(np.uint32))


def test_host_sincos_all_2_32_inputs(tmp_path):
    """Every float bit pattern (2^32) against the host libm — the claim DESIGN.md section 2 rests on."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "trig_exhaustive")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-include", "stdlib.h", "-o", exe,
                           os.path.join(root, "tests", "ctools_trig_exhaustive.c"), "-lm", "-lpthread"])
    out = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "mismatches 0 " in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_device_sincos_matches_libm():
    import ctypes as C
    fp = C.POINTER(C.c_float)
    ang = np.concatenate([_bits(1021).view(np.float32), np.random.default_rng(1).uniform(-7, 7, 1 << 20).astype(np.float32)])
    c1 = np.empty_like(ang)
    s1 = np.empty_like(ang)
    st = sn.lib().cs_device_sincos(0, ang.ctypes.data_as(fp), ang.size, c1.ctypes.data_as(fp), s1.ctypes.data_as(fp))
    assert st == 0
    c2, s2 = orc.libm_sincos(ang)
    _compare(c1, s1, c2, s2)

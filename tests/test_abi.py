"""The C-ABI library builds, loads and exports every symbol include/coreslam_b200.h declares.
No compute calls here (no GPU in this tier)."""
import ctypes
import os
import re

import pytest

import slam.net_b200 as sn
from slam.net_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "coreslam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    L = sn.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "missing export " + n


def test_binding_covers_header():
    assert sorted(N.SIGNATURES) == _declared()


def test_abi_version_and_struct_sizes():
    assert sn.lib().cs_abi_version() == 1
    assert ctypes.sizeof(N.Result) == 32
    assert ctypes.sizeof(N.Config) == 72


def test_no_cpu_fallback_without_device():
    L = sn.lib()
    if L.cs_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(sn.CoreSlamError) as e:
        sn.Processor(40.0, 256, [20, 20, 0], 0.1, 0.1, 10, 2)
    assert e.value.status == 2  # CS_ERR_NO_DEVICE
    with pytest.raises(sn.CoreSlamError):
        sn.ScanLog(4, 64)


def test_product_does_not_import_oracle():
    """The oracle is the checker, never part of the shipped path."""
    pkg = os.path.join(ROOT, "slam.net_b200")
    banned = ("import oracle", "from oracle", "coreslam_oracle", "oracle/", "libcoreslam_oracle")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                for b in banned:
                    assert b not in src, (f, b)


def test_flag_values_agree_between_header_python_and_csharp():
    """CS_FLAG_* of include/coreslam_b200.h = FLAG_* of the ctypes loader = CsFlags of the C# stubs (value by value)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "coreslam_b200.h")).read()
    c_flags = {m.group(1): int(m.group(2), 16) for m in re.finditer(r"#define CS_FLAG_(\w+)\s+0x([0-9a-fA-F]+)u", hdr)}
    assert len(c_flags) >= 9
    from slam.net_b200 import _native as N
    for name, value in c_flags.items():
        assert getattr(N, "FLAG_" + name) == value, name
    cs = open(os.path.join(root, "dotnet", "CoreSlamNative.cs")).read()
    enum = cs[cs.index("public enum CsFlags"):]
    enum = enum[:enum.index("}")]
    cs_flags = {m.group(1).lower(): int(m.group(2), 16) for m in re.finditer(r"(\w+)\s*=\s*0x([0-9a-fA-F]+)", enum)}
    for name, value in c_flags.items():
        assert cs_flags.get(name.replace("_", "").lower()) == value, name

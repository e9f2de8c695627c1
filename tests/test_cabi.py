"""The C-ABI library builds, loads and exports exactly what include/v2e2v_b200.h declares.

CPU-only: no compute entry point is called (there is no GPU here); the host-only entry points
(`v2e_abi_version`, `v2e_error_string`, `v2e_host_pack_scalars`) are.
"""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

from v2e2v_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "v2e2v_b200.h")


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _cabi.lib()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(v2e_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_all_bound_and_exported(lib):
    names = declared_functions()
    assert names, "no functions parsed from the header"
    assert sorted(_cabi.SYMBOLS) == names
    for n in names:
        assert hasattr(lib, n), f"{n} not exported by {_cabi.LIB_PATH}"
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = set(re.findall(r" T (v2e_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_struct_layouts_match_header(lib):
    """sizeof/offsetof from the C compiler == the ctypes mirror."""
    fields = {
        "v2e_config": (_cabi.Config, ["abi_version", "engine", "leak_rate", "philox_rounds", "lut"]),
        "v2e_frame_scalars": (_cabi.FrameScalars, ["t_bin_prev", "sdt", "shot_c"]),
        "v2e_pack_scalars": (_cabi.PackScalars, ["Tr", "shot_f64_mask", "num_frames", "fr"]),
        "v2e_pack": (_cabi.Pack, ["frames", "voxel", "raw_capacity", "shot_imax", "dbg_pol",
                                  "dbg_shot_imax", "philox_seed", "sc"]),
        "v2e_host_params": (_cabi.HostParams, ["num_frames", "t_prev", "qs",
                                               "refractory_period_s"]),
    }
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(){']
    for s, (_, fl) in fields.items():
        prog.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in fl:
            prog.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    prog.append('return 0;}')
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "l.c"), os.path.join(d, "l")
        open(c, "w").write("\n".join(prog))
        subprocess.run(["/usr/bin/gcc", c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = dict(line.split() for line in out.strip().splitlines())
    for s, (ct, fl) in fields.items():
        assert int(got[s]) == ctypes.sizeof(ct), s
        for f in fl:
            assert int(got[f"{s}.{f}"]) == getattr(ct, f).offset, f"{s}.{f}"


def test_version_and_error_strings(lib):
    assert lib.v2e_abi_version() == _cabi.ABI_VERSION
    assert b"later than previous" in lib.v2e_error_string(_cabi.ERR_TIME_ORDER)
    assert lib.v2e_error_string(0) == b"ok"
    assert lib.v2e_launch_count() == 0


def test_constants_match_header():
    src = open(HEADER).read()
    defs = dict(re.findall(r"#define\s+(V2E_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", src))
    assert int(defs["V2E_ABI_VERSION"]) == _cabi.ABI_VERSION
    assert int(defs["V2E_MAX_FRAMES"]) == _cabi.MAX_FRAMES
    assert int(defs["V2E_MAX_BINS"]) == _cabi.MAX_BINS
    assert int(defs["V2E_MAX_SEQ_PER_LAUNCH"]) == _cabi.MAX_SEQ_PER_LAUNCH
    for k in ("OK", "ERR_BAD_ARG", "ERR_UNSUPPORTED", "ERR_TIME_ORDER", "ERR_CUDA",
              "ERR_WORKSPACE", "ERR_NO_DEVICE", "NOISE_OFF", "NOISE_INJECTED", "NOISE_PHILOX",
              "DT_F32", "DT_U8", "DT_BF16", "OUT_VOXEL", "OUT_RAW", "OUT_VOXEL_UNNORM",
              "ENGINE_AUTO", "ENGINE_STEPWISE", "ENGINE_PERSISTENT",
              "ENGINE_RUNAHEAD", "ENGINE_FIXPOINT", "STATUS_OK", "STATUS_NONINTEGER_FRAMES",
              "STATUS_NOT_CONVERGED"):
        assert int(defs["V2E_" + k]) == getattr(_cabi, k), k


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libv2e2v_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _cabi.lib()


def test_emulator_rejects_cpu_device(lib):
    from v2e2v_b200 import EventEmulator
    with pytest.raises(RuntimeError, match="CUDA device only"):
        EventEmulator("voxel_grid", device="cpu")

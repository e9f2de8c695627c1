"""ctypes mirror of include/v2e2v_b200.h and the loader of the in-tree CUDA library.

This is the reference-side binding a maintainer would add: plain ctypes, no
pybind, no torch headers.  The product path has NO CPU fallback: if the CUDA library is missing,
`lib()` raises, and every op in this package goes through `lib()`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

ABI_VERSION = 2
MAX_FRAMES = 32
MAX_BINS = 8
MAX_SEQ_PER_LAUNCH = 16

OK = 0
ERR_BAD_ARG = -1
ERR_UNSUPPORTED = -2
ERR_TIME_ORDER = -3
ERR_CUDA = -4
ERR_WORKSPACE = -5
ERR_NO_DEVICE = -6

STATUS_OK, STATUS_NONINTEGER_FRAMES, STATUS_NOT_CONVERGED = 0, 1, 2
NOISE_OFF, NOISE_INJECTED, NOISE_PHILOX = 0, 1, 2
DT_F32, DT_U8, DT_BF16 = 0, 1, 2
OUT_VOXEL, OUT_RAW, OUT_VOXEL_UNNORM = 0, 1, 2
ENGINE_AUTO, ENGINE_STEPWISE, ENGINE_PERSISTENT, ENGINE_RUNAHEAD, ENGINE_FIXPOINT = 0, 1, 2, 3, 4


class Config(C.Structure):
    """struct v2e_config"""
    _fields_ = [
        ("abi_version", C.c_int32),
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("num_bins", C.c_int32),
        ("cutoff_on", C.c_int32),
        ("leak_on", C.c_int32),
        ("shot_on", C.c_int32),
        ("noise_mode", C.c_int32),
        ("frame_dtype", C.c_int32),
        ("voxel_dtype", C.c_int32),
        ("output_mode", C.c_int32),
        ("engine", C.c_int32),
        ("leak_rate", C.c_float),
        ("leak_jitter", C.c_float),
        ("rn_pos", C.c_float),
        ("rn_neg", C.c_float),
        ("shot_inten_m1", C.c_float),
        ("philox_rounds", C.c_int32),
        ("frames_integer", C.c_int32),
        ("reserved_cfg", C.c_int32),
        ("lut", C.c_void_p),
    ]


class FrameScalars(C.Structure):
    """struct v2e_frame_scalars"""
    _fields_ = [
        ("t_bin_prev", C.c_float),
        ("s_large", C.c_float),
        ("s_small", C.c_float),
        ("sdt", C.c_float),
        ("shot_c", C.c_double),
    ]


class PackScalars(C.Structure):
    """struct v2e_pack_scalars"""
    _fields_ = [
        ("Tr", C.c_float),
        ("duration", C.c_float),
        ("bins_m1", C.c_float),
        ("first_pack", C.c_int32),
        ("shot_f64_mask", C.c_uint32),
        ("num_frames", C.c_int32),
        ("reserved0", C.c_int32),
        ("reserved1", C.c_int32),
        ("fr", FrameScalars * MAX_FRAMES),
    ]


class Pack(C.Structure):
    """struct v2e_pack"""
    _fields_ = [
        ("frames", C.c_void_p),
        ("base_log_frame", C.c_void_p),
        ("lp_log_frame0", C.c_void_p),
        ("timestamp_mem", C.c_void_p),
        ("pos_thres", C.c_void_p),
        ("neg_thres", C.c_void_p),
        ("noise_rate", C.c_void_p),
        ("voxel", C.c_void_p),
        ("raw_events", C.c_void_p),
        ("raw_capacity", C.c_int64),
        ("num_events", C.c_void_p),
        ("status", C.c_void_p),
        ("max_iters", C.c_void_p),
        ("leak_randn", C.c_void_p),
        ("shot_rand", C.c_void_p),
        ("shot_imax", C.c_int32),
        ("reserved", C.c_int32),
        ("dbg_pol", C.c_void_p),
        ("dbg_cnt", C.c_void_p),
        ("dbg_final", C.c_void_p),
        ("dbg_leak_randn", C.c_void_p),
        ("dbg_shot", C.c_void_p),
        ("dbg_shot_imax", C.c_int32),
        ("reserved2", C.c_int32),
        ("philox_seed", C.c_uint64),
        ("philox_offset", C.c_uint64),
        ("sc", PackScalars),
    ]


class HostParams(C.Structure):
    """struct v2e_host_params"""
    _fields_ = [
        ("num_frames", C.c_int32),
        ("num_bins", C.c_int32),
        ("n_stamps", C.c_int32),
        ("stamps_f64", C.c_int32),
        ("first_pack", C.c_int32),
        ("t_prev_f64", C.c_int32),
        ("t_range_f64", C.c_int32),
        ("reserved", C.c_int32),
        ("t_prev", C.c_double),
        ("t_start", C.c_double),
        ("t_end", C.c_double),
        ("cutoff_hz", C.c_double),
        ("ql", C.c_double),
        ("qs", C.c_double),
        ("shot_noise_rate_hz", C.c_double),
        ("refractory_period_s", C.c_float),
        ("reserved_f", C.c_float),
    ]


# Every symbol include/v2e2v_b200.h declares, with its ctypes signature.
SYMBOLS = {
    "v2e_abi_version": (C.c_int, []),
    "v2e_error_string": (C.c_char_p, [C.c_int]),
    "v2e_last_cuda_error": (C.c_int, []),
    "v2e_last_cuda_error_string": (C.c_char_p, []),
    "v2e_workspace_bytes": (C.c_size_t, [C.POINTER(Config), C.c_int]),
    "v2e_resolve_engine": (C.c_int, [C.POINTER(Config), C.c_int]),
    "v2e_launch_count": (C.c_int64, []),
    "v2e_host_pack_scalars": (C.c_int, [C.POINTER(HostParams), C.POINTER(C.c_double),
                                        C.POINTER(C.c_float), C.POINTER(PackScalars)]),
    "v2e_host_clip_prepare": (C.c_int, [C.POINTER(HostParams), C.POINTER(C.c_void_p), C.c_size_t,
                                        C.POINTER(C.c_float), C.POINTER(Pack), C.POINTER(C.c_void_p),
                                        C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                        C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(Pack)]),
    "v2e_pack_forward": (C.c_int, [C.POINTER(Config), C.POINTER(Pack), C.c_int,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "v2e_clip_forward": (C.c_int, [C.POINTER(Config), C.POINTER(Pack), C.c_int,
                                   C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]),
    "v2e_batch_clip_forward": (C.c_int, [C.POINTER(Config), C.POINTER(Pack), C.c_int, C.c_int,
                                         C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]),
    "v2e_events_to_voxel": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_double, C.c_double, C.c_int,
                                      C.c_void_p, C.c_void_p]),
    "v2e_voxel_normalize": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_float,
                                      C.c_void_p, C.c_void_p]),
}

LIB_NAME = "libv2e2v_b200.so"
_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# V2E_LIB: development aid (kernel variants built side by side under v2e2v_b200/_build)
LIB_PATH = os.environ.get("V2E_LIB") or os.path.join(_PKG_DIR, LIB_NAME)

_lib = None
_lock = threading.Lock()


class V2EError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where}: error {code} ({detail})")


def lib() -> C.CDLL:
    """Load the in-tree CUDA library; raise loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (or `python -m v2e2v_b200.build`). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        got = handle.v2e_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH}: ABI version {got}, expected {ABI_VERSION}")
        _lib = handle
        return _lib


def check(code: int, where: str) -> None:
    if code == OK:
        return
    handle = lib()
    detail = handle.v2e_error_string(code).decode()
    if code == ERR_CUDA:
        detail += ": " + handle.v2e_last_cuda_error_string().decode()
    raise V2EError(code, where, detail)

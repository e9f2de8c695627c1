"""Python driver of the CPU oracle (oracle/v2e_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under v2e2v_b200/ imports this module.

`OracleEmulator` mirrors the reference `EventEmulator` (v2e/v2e_model.py:32) call pattern for a
single sequence (B = 1) on numpy arrays: `forward(frames, t_frames) -> (events, num_events)`,
`reset()`, and the state attributes `base_log_frame`, `lp_log_frame0`, `timestamp_mem`.
Randomness is never drawn here: thresholds / noise-rate planes and the per-frame noise draws are
injected by the caller, so a run is a pure function of its inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import torch

from v2e2v_b200 import _cabi
from .host_scalars_torch import pack_scalars_torch

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "_build", "libv2e_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/v2e_oracle.c with gcc (see oracle/Makefile)."""
    src = os.path.join(_DIR, "v2e_oracle.c")
    hdr = os.path.join(_DIR, "..", "include", "v2e2v_b200.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.run(["make", "-C", _DIR, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()   # no-op unless the source or the ABI header is newer than the library
        h = C.CDLL(_LIB_PATH)
        h.v2e_oracle_pack.restype = C.c_int
        h.v2e_oracle_pack.argtypes = [C.POINTER(_cabi.Config), C.POINTER(_cabi.Pack), C.c_int]
        h.v2e_oracle_build_lut.restype = None
        h.v2e_oracle_build_lut.argtypes = [C.c_void_p]
        h.v2e_oracle_events_to_voxel.restype = C.c_int
        h.v2e_oracle_events_to_voxel.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int,
                                                 C.c_int, C.c_int, C.c_void_p]
        h.v2e_oracle_voxel_normalize.restype = C.c_int
        h.v2e_oracle_voxel_normalize.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_float]
        _lib = h
    return _lib


def build_lut() -> np.ndarray:
    lut = np.zeros((2, 256), np.float32)
    lib().v2e_oracle_build_lut(lut.ctypes.data)
    return lut


def _ptr(a):
    return None if a is None else a.ctypes.data


_FRAME_DT = {np.dtype(np.float32): _cabi.DT_F32, np.dtype(np.uint8): _cabi.DT_U8,
             np.dtype(np.uint16): _cabi.DT_BF16}


class OracleEmulator:
    """Single-sequence CPU oracle with the reference emulator's call pattern."""

    def __init__(self, output_mode, pl=1, ps=1, ql=1, qs=1, num_bins=5, pos_thres=0.2,
                 neg_thres=0.2, sigma_thres=0.03, cutoff_hz=0, leak_rate_hz=0.1,
                 refractory_period_s=0, shot_noise_rate_hz=0, leak_jitter_fraction=0.1,
                 noise_rate_cov_decades=0.1, threads=1, voxel_bf16=False, normalize=True):
        self.output_mode = output_mode
        self.pl, self.ps, self.ql, self.qs = pl, ps, ql, qs
        self.num_bins = num_bins
        self.pos_nominal, self.neg_nominal = pos_thres, neg_thres
        self.sigma_thres = sigma_thres
        self.cutoff_hz = cutoff_hz
        self.leak_rate_hz = leak_rate_hz
        self.refractory_period_s = refractory_period_s
        self.shot_noise_rate_hz = shot_noise_rate_hz
        self.leak_jitter_fraction = leak_jitter_fraction
        self.noise_rate_cov_decades = noise_rate_cov_decades
        self.threads = threads
        self.voxel_bf16 = voxel_bf16
        self.normalize = normalize
        self.lut = build_lut()
        # injected per-sequence planes (set before the first forward after reset())
        self.pos_thres = None
        self.neg_thres = None
        self.noise_rate_array = None
        self.reset()

    def reset(self):
        self.base_log_frame = None
        self.lp_log_frame0 = None
        self.timestamp_mem = None
        self.t_previous = None
        self.last = {}

    def set_planes(self, pos_thres, neg_thres, noise_rate_array=None):
        self.pos_thres = np.ascontiguousarray(pos_thres, np.float32).reshape(-1)
        self.neg_thres = np.ascontiguousarray(neg_thres, np.float32).reshape(-1)
        self.noise_rate_array = (None if noise_rate_array is None else
                                 np.ascontiguousarray(noise_rate_array, np.float32).reshape(-1))

    def config(self, H, W, frame_dtype) -> _cabi.Config:
        cfg = _cabi.Config()
        cfg.abi_version = _cabi.ABI_VERSION
        cfg.height, cfg.width, cfg.num_bins = H, W, self.num_bins
        cfg.cutoff_on = int(self.cutoff_hz > 0)
        cfg.leak_on = int(self.leak_rate_hz > 0)
        cfg.shot_on = int(self.shot_noise_rate_hz > 0)
        cfg.noise_mode = (_cabi.NOISE_INJECTED if (cfg.leak_on or cfg.shot_on)
                          else _cabi.NOISE_OFF)
        cfg.frame_dtype = frame_dtype
        cfg.voxel_dtype = _cabi.DT_BF16 if self.voxel_bf16 else _cabi.DT_F32
        if self.output_mode == 'raw':
            cfg.output_mode = _cabi.OUT_RAW
        else:
            cfg.output_mode = _cabi.OUT_VOXEL if self.normalize else _cabi.OUT_VOXEL_UNNORM
        f32 = torch.float32
        cfg.leak_rate = float(torch.tensor(self.leak_rate_hz, dtype=f32))
        cfg.leak_jitter = float(torch.tensor(self.leak_jitter_fraction, dtype=f32))
        # 1./nominal on an f32 tensor == reciprocal()*1. (v2e_model.py:217-218)
        cfg.rn_pos = float(1. / torch.tensor([self.pos_nominal], dtype=f32))
        cfg.rn_neg = float(1. / torch.tensor([self.neg_nominal], dtype=f32))
        cfg.shot_inten_m1 = float(torch.tensor(0.25 - 1, dtype=f32))
        cfg.lut = self.lut.ctypes.data
        return cfg

    def forward(self, frames, t_frames, leak_randn=None, shot_rand=None, debug=False,
                scalars=None, raw_capacity=None):
        """frames: [1,T,H,W] or [T,H,W] numpy/torch (f32, u8, or bf16-as-uint16).
        t_frames: torch CPU tensor [1,2] or [1,>=T].  leak_randn: f32 [T-1,P];
        shot_rand: f32 [T-1,Imax,P].  scalars: optional precomputed _cabi.PackScalars."""
        if isinstance(frames, torch.Tensor):
            frames = frames.detach().cpu().numpy()
        frames = np.ascontiguousarray(frames)
        if frames.ndim == 4:
            assert frames.shape[0] == 1, "oracle sequences are B=1"
            frames = frames[0]
        T, H, W = frames.shape
        P = H * W
        first = self.base_log_frame is None
        if scalars is None:
            sc, t_prev = pack_scalars_torch(
                t_frames, T, self.num_bins, self.refractory_period_s, self.cutoff_hz, self.ql,
                self.qs, self.shot_noise_rate_hz, self.t_previous, first)
            self.t_previous = t_prev
        else:
            sc = scalars
        cfg = self.config(H, W, _FRAME_DT[frames.dtype])
        if first:
            self.base_log_frame = np.zeros(P, np.float32)
            self.lp_log_frame0 = np.zeros(P, np.float32)
            self.timestamp_mem = np.zeros(P, np.float32)
        pk = _cabi.Pack()
        pk.sc = sc
        pk.frames = frames.ctypes.data
        pk.base_log_frame = self.base_log_frame.ctypes.data
        pk.lp_log_frame0 = self.lp_log_frame0.ctypes.data
        pk.timestamp_mem = self.timestamp_mem.ctypes.data
        pk.pos_thres = self.pos_thres.ctypes.data
        pk.neg_thres = self.neg_thres.ctypes.data
        pk.noise_rate = _ptr(self.noise_rate_array)
        keep = [frames]
        nb = self.num_bins
        voxel = raw = None
        if self.output_mode == 'raw':
            cap = raw_capacity if raw_capacity is not None else P * (T - 1) * 4
            raw = np.zeros((cap, 5), np.float32)
            pk.raw_events, pk.raw_capacity = raw.ctypes.data, cap
        else:
            voxel = np.zeros((nb, H, W), np.uint16 if self.voxel_bf16 else np.float32)
            pk.voxel = voxel.ctypes.data
        nev = np.zeros(1, np.int64)
        mxi = np.zeros(T, np.int32)
        pk.num_events, pk.max_iters = nev.ctypes.data, mxi.ctypes.data
        if leak_randn is not None:
            leak_randn = np.ascontiguousarray(leak_randn, np.float32).reshape(T - 1, P)
            pk.leak_randn = leak_randn.ctypes.data
        if shot_rand is not None:
            shot_rand = np.ascontiguousarray(shot_rand, np.float32)
            shot_rand = shot_rand.reshape(T - 1, -1, P)
            pk.shot_rand, pk.shot_imax = shot_rand.ctypes.data, shot_rand.shape[1]
        dbg = {}
        if debug:
            dbg = dict(pol=np.zeros((T - 1, P), np.int8), cnt=np.zeros((T - 1, P), np.int32),
                       final=np.zeros((T - 1, P), np.int32))
            pk.dbg_pol, pk.dbg_cnt = dbg['pol'].ctypes.data, dbg['cnt'].ctypes.data
            pk.dbg_final = dbg['final'].ctypes.data
        rc = lib().v2e_oracle_pack(C.byref(cfg), C.byref(pk), int(self.threads))
        if rc != 0:
            raise RuntimeError(f"v2e_oracle_pack failed: {rc}")
        self.last = dict(max_iters=mxi, scalars=sc, **dbg)
        n = int(nev[0])
        if self.output_mode == 'raw':
            return raw[:min(n, raw.shape[0])], n
        return voxel[None], n


def events_to_voxel(events: np.ndarray, num_bins, width, height, torch_semantics=False):
    """Oracle for utils/event_process.py:15-63 / 66-129 (does not mutate `events`)."""
    events = np.ascontiguousarray(events)
    assert events.ndim == 2 and events.shape[1] == 4 or len(events) == 0
    f64 = events.dtype == np.float64
    if not f64:
        events = events.astype(np.float32, copy=False)
    vox = np.zeros((num_bins, height, width), np.float32)
    rc = lib().v2e_oracle_events_to_voxel(events.ctypes.data, int(f64), len(events), num_bins,
                                          width, height, int(torch_semantics), vox.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"v2e_oracle_events_to_voxel failed: {rc}")
    return vox


def voxel_normalize(voxel: np.ndarray, mode='std', hot_pixel_thresh=0.0):
    """Oracle for utils/event_process.py:132-154 / 157-177 (returns a new array)."""
    out = np.ascontiguousarray(voxel, np.float32).copy()
    rc = lib().v2e_oracle_voxel_normalize(out.ctypes.data, out.size, 0 if mode == 'std' else 1,
                                          float(hot_pixel_thresh))
    if rc != 0:
        raise RuntimeError(f"v2e_oracle_voxel_normalize failed: {rc}")
    return out

"""Host-side per-pack scalar preparation (product path).

Wraps ``v2e_host_pack_scalars`` (csrc/host_scalars.c), which restates the scalar part of
``EventEmulator.forward`` (v2e/v2e_model.py:303-318,335,353,521) with the dtype promotion torch
applies to 0-dim tensors.  The two ``torch.linspace`` tables the reference builds are taken from
torch itself (linspace's rounding is implementation-defined) and cached.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache

import numpy as np
import torch

from . import _cabi


@lru_cache(maxsize=64)
def time_frames_table(num_frames: int, num_bins: int) -> np.ndarray:
    """time_frames = linspace(0, duration*(T-1), T, f32)  (v2e_model.py:313-314)."""
    duration = (num_bins - 1) / (num_frames - 1)
    t = torch.linspace(start=0, end=duration * (num_frames - 1), steps=num_frames,
                       dtype=torch.float32)
    return np.ascontiguousarray(t.numpy())


class TimeState:
    """``self.t_previous`` of the reference: a value and the dtype torch would carry it in."""
    __slots__ = ("value", "is_f64")

    def __init__(self, value: float = 0.0, is_f64: bool = True):
        self.value = value
        self.is_f64 = is_f64


def pack_scalars(t_frames: torch.Tensor, num_frames: int, num_bins: int,
                 refractory_period_s: float, cutoff_hz: float, ql: float, qs: float,
                 shot_noise_rate_hz: float, t_prev: TimeState, first_pack: bool,
                 out: _cabi.PackScalars = None) -> _cabi.PackScalars:
    """Fill a PackScalars for one pack and advance ``t_prev`` in place.

    t_frames: [B>=1, 2] or [B, >= num_frames], float32/float64, any device (row 0 is used, like
    the reference).  Raises ValueError when frame 1 is not later than the previous frame time.
    """
    if t_frames.dim() != 2:
        raise ValueError("t_frames must be [batch, 2] or [batch, num_frames]")
    tf = t_frames.detach()
    if tf.is_cuda:
        tf = tf.cpu()                     # the reference reads these on the host as well
    if tf.dtype not in (torch.float32, torch.float64):
        tf = tf.to(torch.float32)         # integer stamps behave like default-dtype floats
    range_f64 = tf.dtype == torch.float64
    # (numpy views from here on: a torch indexing op costs microseconds, and this runs per pack)
    row_np = tf.numpy()[0]
    n_cols = row_np.shape[0]
    t_start, t_end = float(row_np[0]), float(row_np[-1])
    if n_cols == 2:
        # v2e_model.py:307-309: f32 linspace between the two stamps
        row = tf[0]
        tfl = torch.linspace(start=row[0], end=row[-1], steps=num_frames, dtype=torch.float32)
        stamps = tfl.to(torch.float64).numpy()
        stamps_f64 = False
    else:
        if n_cols < num_frames:
            raise ValueError(f"need {num_frames} stamps, got {n_cols}")
        stamps = row_np.astype(np.float64, copy=False)   # (f32 -> f64 is exact)
        stamps_f64 = range_f64
    stamps = np.ascontiguousarray(stamps)
    hp = _cabi.HostParams()
    hp.num_frames, hp.num_bins = num_frames, num_bins
    hp.n_stamps = len(stamps)
    hp.stamps_f64 = int(stamps_f64)
    hp.first_pack = int(first_pack)
    hp.t_prev_f64 = int(t_prev.is_f64)
    hp.t_range_f64 = int(range_f64)
    hp.t_prev = t_prev.value
    hp.t_start, hp.t_end = t_start, t_end
    hp.cutoff_hz, hp.ql, hp.qs = float(cutoff_hz), float(ql), float(qs)
    hp.shot_noise_rate_hz = float(shot_noise_rate_hz)
    hp.refractory_period_s = float(refractory_period_s)
    tfr = time_frames_table(num_frames, num_bins)
    sc = out if out is not None else _cabi.PackScalars()
    rc = _cabi.lib().v2e_host_pack_scalars(
        C.byref(hp), stamps.ctypes.data_as(C.POINTER(C.c_double)),
        tfr.ctypes.data_as(C.POINTER(C.c_float)), C.byref(sc))
    if rc == _cabi.ERR_TIME_ORDER:
        prev = hp.t_start if first_pack else t_prev.value
        raise ValueError("this frame time={} must be later than previous frame time={}".format(
            stamps[1], prev))
    _cabi.check(rc, "v2e_host_pack_scalars")
    t_prev.value = float(stamps[num_frames - 1])
    t_prev.is_f64 = stamps_f64
    return sc


class PackScalarPrep:
    """``pack_scalars`` for one emulator: the parameters that never change are set once, and the
    per-pack work is what the stamps require (it runs once per pack on the enqueue path, where a
    microsecond of Python is a microsecond the device may wait)."""

    def __init__(self, num_frames: int, num_bins: int, refractory_period_s: float, cutoff_hz: float,
                 ql: float, qs: float, shot_noise_rate_hz: float):
        self.key = (num_frames, num_bins, float(refractory_period_s), float(cutoff_hz), float(ql),
                    float(qs), float(shot_noise_rate_hz))
        self.raw_key = None    # the caller's unconverted attribute values this key was built from
        hp = self.hp = _cabi.HostParams()
        hp.num_frames, hp.num_bins = num_frames, num_bins
        hp.cutoff_hz, hp.ql, hp.qs = float(cutoff_hz), float(ql), float(qs)
        hp.shot_noise_rate_hz = float(shot_noise_rate_hz)
        hp.refractory_period_s = float(refractory_period_s)
        self._hp_ref = C.byref(hp)
        self._tfr = time_frames_table(num_frames, num_bins)
        self._tfr_ptr = self._tfr.ctypes.data_as(C.POINTER(C.c_float))
        self._fn = _cabi.lib().v2e_host_pack_scalars
        self._dbl_p = C.POINTER(C.c_double)

    def fill(self, t_frames: torch.Tensor, t_prev: TimeState, first_pack: bool,
             out: _cabi.PackScalars) -> _cabi.PackScalars:
        """Same contract as ``pack_scalars`` (fills ``out``, advances ``t_prev``)."""
        tf = t_frames
        if (tf.dim() != 2 or tf.is_cuda or tf.dtype != torch.float64 or tf.requires_grad
                or tf.shape[1] < self.hp.num_frames or tf.shape[1] == 2):
            k = self.key   # anything but a host f64 [B, >= T] stamp tensor: the general routine
            return pack_scalars(t_frames, k[0], k[1], k[2], k[3], k[4], k[5], k[6], t_prev, first_pack,
                                out=out)
        stamps = tf.numpy()[0]
        if not stamps.flags["C_CONTIGUOUS"]:
            stamps = np.ascontiguousarray(stamps)
        hp = self.hp
        hp.n_stamps = stamps.shape[0]
        hp.stamps_f64 = 1
        hp.first_pack = int(first_pack)
        hp.t_prev_f64 = int(t_prev.is_f64)
        hp.t_range_f64 = 1
        hp.t_prev = t_prev.value
        hp.t_start, hp.t_end = float(stamps[0]), float(stamps[-1])
        rc = self._fn(self._hp_ref, C.cast(stamps.__array_interface__["data"][0], self._dbl_p),
                      self._tfr_ptr, C.byref(out))
        if rc == _cabi.ERR_TIME_ORDER:
            prev = hp.t_start if first_pack else t_prev.value
            raise ValueError("this frame time={} must be later than previous frame time={}".format(
                stamps[1], prev))
        _cabi.check(rc, "v2e_host_pack_scalars")
        t_prev.value = float(stamps[hp.num_frames - 1])
        t_prev.is_f64 = True
        return out

"""`EventEmulator`: drop-in for the reference DVS emulator, backed by the sm_100a kernels.

Mirrors `v2e.v2e_model.EventEmulator` of lsying009/V2E2V (v2e/v2e_model.py:32-536): same
constructor keywords, `forward(frames, t_frames) -> (events, num_events)`, `reset()`, and the
state attributes other code reads (`base_log_frame`, `lp_log_frame0`, `timestamp_mem`,
`pos_thres`, `neg_thres`, `noise_rate_array`, `t_previous`, `num_bins`, `output_mode`).

What runs where
  * per-sequence initialisation (`_init`, v2e_model.py:156-240): the random per-pixel thresholds
    and noise-rate plane are drawn with the same torch calls, in the same order, as the
    reference, so a given torch seed gives the same sensor.
  * per-pack scalars (v2e_model.py:303-318): host, `host_scalars.pack_scalars`.
  * everything per pixel (lin-log, low-pass, leak, thresholding, refractory, voxel binning,
    normalisation): one call of `v2e_pack_forward` through the `v2e2v::emulate_pack` op.

Differences from the reference, all documented in DESIGN.md:
  * batch_size must be 1 (the reference's B>1 mixes sequences, SURVEY.md trap T1); use
    `EventEmulatorBatch` for many independent sequences in one launch.
  * the per-frame noise draws come from an in-kernel Philox generator (`noise_mode='philox'`),
    not from torch's global generator; `noise_mode='injected'` consumes caller-supplied draws
    and is bit-exact to the reference given the same draws.
  * refractory_period_s == 0 is accepted (the reference crashes on the second pack, trap T2).
"""
from __future__ import annotations

import ctypes as C
import logging
import math
import random
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _cabi, ops
from .host_scalars import PackScalarPrep, TimeState, pack_scalars

logger = logging.getLogger(__name__)

_LUT_CACHE = {}


def lin_log_lut(device) -> torch.Tensor:
    """[2,256] f32: lin_log(v) (v2e/emulator_utils.py:13-37) and (v+20)/275 (:40-45), v=0..255.

    Built once per device with the reference's own arithmetic: fp64 piecewise lin/log rounded to
    1e-8, then cast to f32; f32 add and divide for the rescale.
    """
    key = str(device)
    lut = _LUT_CACHE.get(key)
    if lut is None:
        v = torch.arange(256, dtype=torch.float32)
        x = v.double()
        f = (1. / 20) * math.log(20)
        y = torch.where(x <= 20, x * f, torch.log(x))
        y = torch.round(y * 1e8) / 1e8
        lut = torch.stack([y.float(), (v + 20) / 275.]).contiguous().to(device)
        _LUT_CACHE[key] = lut
    return lut


def draw_thresholds(nominal: torch.Tensor, coef_large, coef_small, sigma, H, W, device):
    """Per-pixel contrast thresholds (v2e_model.py:183-198): N(pl*C, sigma) everywhere, pixels
    at even row and even column re-drawn from N(ps*C, sigma), clamped to >= 0.01.  Uses the
    global CPU generator with the same draw order as the reference."""
    B = nominal.numel()
    large = torch.stack([torch.normal(coef_large * nominal[i], sigma, size=[1, H, W],
                                      dtype=torch.float32).to(device) for i in range(B)], 0)
    small = torch.stack([torch.normal(coef_small * nominal[i], sigma, size=[1, H, W],
                                      dtype=torch.float32).to(device) for i in range(B)], 0)
    large[:, :, 0::2, 0::2] = small[:, :, 0::2, 0::2]
    return torch.clamp(large, min=0.01).contiguous()


class EventEmulator(nn.Module):
    """compute events based on the input frame (B200 kernels; see module docstring)."""

    def __init__(
            self,
            output_mode,
            pl=1, ps=1, ql=1, qs=1,
            num_bins=5,
            pos_thres=0.2,
            neg_thres=0.2,
            sigma_thres=0.03,
            cutoff_hz=0,
            leak_rate_hz=0.1,
            refractory_period_s=0,
            shot_noise_rate_hz=0,
            leak_jitter_fraction=0.1,
            noise_rate_cov_decades=0.1,
            seed=0,
            show_dvs_model_state: str = None,
            device="cuda",
            # ---- extensions (keyword-only in spirit; defaults keep the reference behaviour) ----
            noise_mode: Optional[str] = None,     # 'philox' | 'injected' | 'off'
            engine: str = "auto",                 # 'auto' | 'fixpoint' | 'runahead' | 'persistent' | 'stepwise'
            voxel_dtype: torch.dtype = torch.float32,
            philox_rounds: int = 10,
            normalize: bool = True,
            raw_capacity: Optional[int] = None,
            debug_maps: bool = False,
            debug_shot_imax: int = 24,
            lazy_num_events: bool = False,
            integer_frames: Optional[bool] = None):
        super().__init__()
        if output_mode not in ("raw", "voxel_grid"):
            raise ValueError("output_mode must be 'raw' or 'voxel_grid'")
        if show_dvs_model_state:
            raise NotImplementedError(
                "show_dvs_model_state needs a display (cv2.imshow); use debug_maps=True and read "
                "`last_debug` instead")
        if not (1 <= num_bins <= _cabi.MAX_BINS):
            raise ValueError(f"num_bins must be in [1,{_cabi.MAX_BINS}]")
        logger.info("ON/OFF log_e temporal contrast thresholds: {} / {} +/- {}".format(
            pos_thres, neg_thres, sigma_thres))
        self.base_log_frame = None
        self.t_previous = None
        self.lp_log_frame0 = None
        self.device = device
        self.output_mode = output_mode
        self.num_bins = num_bins
        self.ql, self.qs, self.cutoff_hz = ql, qs, cutoff_hz
        self.pl, self.ps, self.sigma_thres = pl, ps, sigma_thres
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("v2e2v_b200.EventEmulator runs on a CUDA device only "
                               "(no CPU fallback); got device=%r" % (device,))
        f32 = torch.float32
        self.pos_thres = torch.tensor(pos_thres, dtype=f32, device=dev)
        self.neg_thres = torch.tensor(neg_thres, dtype=f32, device=dev)
        self.pos_thres_nominal = torch.tensor(pos_thres, dtype=f32, device=dev)
        self.neg_thres_nominal = torch.tensor(neg_thres, dtype=f32, device=dev)
        self._pos_nominal_host = torch.tensor(pos_thres, dtype=f32)
        self._neg_nominal_host = torch.tensor(neg_thres, dtype=f32)
        self.leak_rate_hz = leak_rate_hz
        self.refractory_period_s = torch.tensor(refractory_period_s, dtype=f32, device=dev)
        self._refractory_host = float(torch.tensor(refractory_period_s, dtype=f32))
        self.shot_noise_rate_hz = shot_noise_rate_hz
        self.leak_jitter_fraction = leak_jitter_fraction
        self.noise_rate_cov_decades = noise_rate_cov_decades
        self.SHOT_NOISE_INTEN_FACTOR = 0.25
        self.show_input = show_dvs_model_state
        if seed != 0:
            torch.manual_seed(seed)
            np.random.seed(seed)
            random.seed(seed)
        self.num_events = 0
        self.frame_counter = 0

        noisy = leak_rate_hz > 0 or shot_noise_rate_hz > 0
        if noise_mode is None:
            noise_mode = "philox" if noisy else "off"
        if noise_mode not in ("philox", "injected", "off"):
            raise ValueError("noise_mode must be 'philox', 'injected' or 'off'")
        if noise_mode == "off" and noisy:
            raise ValueError("noise_mode='off' requires leak_rate_hz == shot_noise_rate_hz == 0")
        self.noise_mode = noise_mode
        self.engine = engine
        self.voxel_dtype = voxel_dtype
        self.philox_rounds = philox_rounds
        self.normalize = normalize
        self.raw_capacity = raw_capacity
        self.debug_maps = debug_maps
        self.debug_shot_imax = debug_shot_imax
        self.lazy_num_events = lazy_num_events
        self._clip_trace = None   # development aid: list that receives (label, perf_counter) marks
        self.clip_ramp = None   # development aid: packs per enqueue chunk of forward_clip (None: Fibonacci ramp)
        # None: uint8 frames are integer by construction, float frames are not assumed to be.
        # True: the caller guarantees integer-valued frames in 0..255 (what every reader of the
        # reference produces); a violation is detected on the device and raised as ValueError.
        self.integer_frames = integer_frames
        self.last_debug = {}
        self.noise_rate_array = None
        self.timestamp_mem = None
        # injected-noise inputs for the next forward() (noise_mode='injected')
        self.inject_leak_randn = None
        self.inject_shot_rand = None
        self.philox_seed = (int(torch.initial_seed()) ^ 0x5851F42D4C957F2D) & (2 ** 64 - 1)
        self._pack_counter = 0
        self._first_pending = False
        self._clip_tmpl = _cabi.Pack()
        self._t_prev = TimeState()
        self._launch = None
        self._launch_key = None
        self._handle = None
        self._workspace = None
        self._dev = dev
        self._scalars = _cabi.PackScalars()

    # ------------------------------------------------------------------------------------------
    def _init(self, height, width):
        """Per-sequence state (v2e_model.py:156-240).  base/lp/timestamp_mem are filled by the
        kernel on the first pack (base = lp = lin_log(frame 0), timestamp_mem = -Tr)."""
        dev = self._dev
        shape = (1, 1, height, width)
        self.base_log_frame = torch.empty(shape, dtype=torch.float32, device=dev)
        self.lp_log_frame0 = torch.empty(shape, dtype=torch.float32, device=dev)
        self.pos_thres_real_nominal = self.pos_thres_nominal * torch.ones(1, device=dev)
        self.neg_thres_real_nominal = self.neg_thres_nominal * torch.ones(1, device=dev)
        if self.sigma_thres > 0:
            self.pos_thres = draw_thresholds(self._pos_nominal_host.reshape(1), self.pl, self.ps,
                                             self.sigma_thres, height, width, dev)
            self.neg_thres = draw_thresholds(self._neg_nominal_host.reshape(1), self.pl, self.ps,
                                             self.sigma_thres, height, width, dev)
        else:
            # the reference crashes for sigma_thres == 0 (trap T2); use the nominal thresholds
            pos = torch.full(shape, float(self.pl * self._pos_nominal_host), device=dev)
            pos[:, :, 0::2, 0::2] = float(self.ps * self._pos_nominal_host)
            neg = torch.full(shape, float(self.pl * self._neg_nominal_host), device=dev)
            neg[:, :, 0::2, 0::2] = float(self.ps * self._neg_nominal_host)
            self.pos_thres = torch.clamp(pos, min=0.01)
            self.neg_thres = torch.clamp(neg, min=0.01)
        # shot-noise pre-probabilities (v2e_model.py:217-218) are recomputed in-kernel
        if self.leak_rate_hz > 0:
            z = torch.randn(shape, dtype=torch.float32, device=dev)
            self.noise_rate_array = torch.exp(math.log(10) * self.noise_rate_cov_decades * z)
        self.timestamp_mem = torch.empty(shape, dtype=torch.float32, device=dev)

    def set_planes(self, pos_thres, neg_thres, noise_rate_array=None):
        """Install externally generated per-pixel planes (tests, checkpointed sensors)."""
        dev = self._dev

        def prep(a):
            t = torch.as_tensor(a, dtype=torch.float32).to(dev)
            return t.reshape(1, 1, *t.shape[-2:]).contiguous()
        self.pos_thres, self.neg_thres = prep(pos_thres), prep(neg_thres)
        if noise_rate_array is not None:
            self.noise_rate_array = prep(noise_rate_array)
        self._planes_external = True

    def reset(self):
        """resets so that next use will reinitialize the base frame (v2e_model.py:242-250)"""
        self.num_events_total = 0
        self.num_events_on = 0
        self.num_events_off = 0
        self.base_log_frame = None
        self.lp_log_frame0 = None
        self.frame_counter = 0
        self._first_pending = False

    # ------------------------------------------------------------------------------------------
    def _config(self, H, W, frame_dtype) -> _cabi.Config:
        cfg = _cabi.Config()
        cfg.abi_version = _cabi.ABI_VERSION
        cfg.height, cfg.width, cfg.num_bins = H, W, self.num_bins
        cfg.cutoff_on = int(self.cutoff_hz > 0)
        cfg.leak_on = int(self.leak_rate_hz > 0)
        cfg.shot_on = int(self.shot_noise_rate_hz > 0)
        cfg.noise_mode = {"off": _cabi.NOISE_OFF, "injected": _cabi.NOISE_INJECTED,
                          "philox": _cabi.NOISE_PHILOX}[self.noise_mode]
        cfg.frame_dtype = ops._FRAME_DT[frame_dtype]
        cfg.voxel_dtype = ops._VOX_DT[self.voxel_dtype]
        if self.output_mode == "raw":
            cfg.output_mode = _cabi.OUT_RAW
        else:
            cfg.output_mode = _cabi.OUT_VOXEL if self.normalize else _cabi.OUT_VOXEL_UNNORM
        cfg.engine = {"auto": _cabi.ENGINE_AUTO, "stepwise": _cabi.ENGINE_STEPWISE,
                      "persistent": _cabi.ENGINE_PERSISTENT,
                      "runahead": _cabi.ENGINE_RUNAHEAD,
                      "fixpoint": _cabi.ENGINE_FIXPOINT}[self.engine]
        f32 = torch.float32
        cfg.leak_rate = float(torch.tensor(self.leak_rate_hz, dtype=f32))
        cfg.leak_jitter = float(torch.tensor(self.leak_jitter_fraction, dtype=f32))
        cfg.rn_pos = float(1. / self._pos_nominal_host.reshape(1))   # reciprocal()*1.
        cfg.rn_neg = float(1. / self._neg_nominal_host.reshape(1))
        cfg.shot_inten_m1 = float(torch.tensor(self.SHOT_NOISE_INTEN_FACTOR - 1, dtype=f32))
        cfg.philox_rounds = self.philox_rounds
        cfg.frames_integer = int(bool(self.integer_frames) or frame_dtype == torch.uint8)
        self._lut = lin_log_lut(self._dev)
        cfg.lut = self._lut.data_ptr()
        return cfg

    def _get_launch(self, H, W, frame_dtype) -> ops.PackLaunch:
        key = (H, W, frame_dtype)
        if self._launch_key != key:
            if self._handle is not None:
                ops.release_launch(self._handle)
            self._launch = ops.PackLaunch(self._config(H, W, frame_dtype), 1)
            self._handle = ops.register_launch(self._launch)
            self._launch_key = key
            self._workspace = torch.zeros(self._launch.workspace_bytes, dtype=torch.uint8,
                                          device=self._dev)
        return self._launch

    @staticmethod
    def _raise_on_status(status: int) -> None:
        if status == _cabi.STATUS_NONINTEGER_FRAMES:
            raise ValueError("integer_frames=True was asserted but a frame value is not an "
                             "integer in 0..255: the outputs of this pack are invalid")
        if status == _cabi.STATUS_NOT_CONVERGED:
            raise RuntimeError("fixed-point engine: the per-frame maxima did not settle "
                               "(internal error): the outputs of this pack are invalid")

    def check_status(self) -> None:
        """Raise if the last pack reported a contract violation (synchronises)."""
        clip = getattr(self, "_clip_meta", None)
        if clip is not None:
            self._clip_meta = None
            self._raise_on_status(int(clip[:, 2].max().item()))
        meta = getattr(self, "_last_meta", None)
        if meta is not None:
            self._raise_on_status(int(meta[2].item()))

    @property
    def resolved_engine(self) -> Optional[str]:
        if self._launch is None:
            return None
        return {_cabi.ENGINE_STEPWISE: "stepwise",
                _cabi.ENGINE_PERSISTENT: "persistent",
                _cabi.ENGINE_RUNAHEAD: "runahead",
                _cabi.ENGINE_FIXPOINT: "fixpoint"}.get(self._launch.resolve_engine())

    def _first_pack_init(self, height, width) -> bool:
        """True (after creating the per-sequence state, v2e_model.py:325-327) when the next pack
        is the first one after reset()."""
        first = self.base_log_frame is None
        if first:
            ext = getattr(self, "_planes_external", False)
            keep = (self.pos_thres, self.neg_thres, self.noise_rate_array) if ext else None
            self._init(height, width)
            if ext:
                self.pos_thres, self.neg_thres, self.noise_rate_array = keep
        elif self._first_pending:
            # (a refused chunk preparation had already created the state)
            first, self._first_pending = True, False
        return first

    def _get_scalar_prep(self, num_frames) -> PackScalarPrep:
        # (the raw attribute values first: the float() conversions of the full key cost more than
        # the C routine they feed)
        raw = (num_frames, self.num_bins, self._refractory_host, self.cutoff_hz, self.ql, self.qs,
               self.shot_noise_rate_hz)
        prep = self.__dict__.get("_scalar_prep")
        if prep is not None and prep.raw_key == raw:
            return prep
        key = (num_frames, self.num_bins, float(self._refractory_host), float(self.cutoff_hz),
               float(self.ql), float(self.qs), float(self.shot_noise_rate_hz))
        if prep is None or prep.key != key:
            prep = PackScalarPrep(*key)
            object.__setattr__(self, "_scalar_prep", prep)
        prep.raw_key = raw
        return prep

    def _prep_chunk_c(self, pk0, n, fptrs, rows, n_stamps, num_frames, height, width,
                      vox_ptr, vox_stride, meta_ptr, meta_stride=16, frame_offset=0, row_offset=0,
                      out_stride=1) -> bool:
        """`_prep_pack` for n consecutive packs in ONE C call (`v2e_host_clip_prepare`): frames by
        device address, stamps as host float64 rows of n_stamps >= num_frames entries.  False
        (nothing changed) when the C routine refuses - the caller then takes the per-pack path,
        which raises what has to be raised."""
        first = self._first_pack_init(height, width)
        prep = self._get_scalar_prep(num_frames)
        hp = prep.hp
        hp.n_stamps = n_stamps
        hp.first_pack = int(first)
        hp.t_prev = self._t_prev.value
        hp.t_prev_f64 = int(self._t_prev.is_f64)
        t = self._clip_tmpl
        t.base_log_frame = self.base_log_frame.data_ptr()
        t.lp_log_frame0 = self.lp_log_frame0.data_ptr()
        t.timestamp_mem = self.timestamp_mem.data_ptr()
        t.pos_thres = self.pos_thres.data_ptr()
        t.neg_thres = self.neg_thres.data_ptr()
        t.noise_rate = ops._ptr(self.noise_rate_array)
        t.voxel = vox_ptr
        t.num_events = meta_ptr
        t.status = meta_ptr + 8
        t.philox_seed = self.philox_seed
        t.philox_offset = self._pack_counter
        tp, bad = C.c_double(0.0), C.c_int32(-1)
        rc = _cabi.lib().v2e_host_clip_prepare(prep._hp_ref, rows, row_offset, prep._tfr_ptr, C.byref(t),
                                               fptrs, frame_offset, vox_stride, meta_stride, n,
                                               out_stride, C.byref(tp), C.byref(bad), pk0)
        if rc != _cabi.OK:
            if first:   # (the per-pack path must see a first pack again)
                self._first_pending = True
            return False
        self._t_prev.value, self._t_prev.is_f64 = tp.value, True
        # (plain attributes: written past nn.Module.__setattr__, which costs ~2.5 us a piece)
        d = self.__dict__
        d["t_previous"] = tp.value
        d["frame_counter"] += n * num_frames
        d["_pack_counter"] += n
        return True

    def _prep_pack(self, pk, fr, t_frames, vox_ptr, meta_ptr):
        """Fill one v2e_pack (voxel output, no debug maps) for frames `fr` [T,H,W] and advance the
        host-side state (time, pack counter).  meta_ptr: 16 zeroed bytes (num_events i64, status)."""
        num_frames, height, width = fr.shape
        first = self._first_pack_init(height, width)
        prep = self._get_scalar_prep(num_frames)
        prep.fill(t_frames, self._t_prev, first, pk.sc)
        self.t_previous = self._t_prev.value
        self.frame_counter += num_frames
        pk.frames = fr.data_ptr()
        pk.base_log_frame = self.base_log_frame.data_ptr()
        pk.lp_log_frame0 = self.lp_log_frame0.data_ptr()
        pk.timestamp_mem = self.timestamp_mem.data_ptr()
        pk.pos_thres = self.pos_thres.data_ptr()
        pk.neg_thres = self.neg_thres.data_ptr()
        pk.noise_rate = ops._ptr(self.noise_rate_array)
        pk.voxel = vox_ptr
        pk.num_events = meta_ptr
        pk.status = meta_ptr + 8
        pk.philox_seed = self.philox_seed
        pk.philox_offset = self._pack_counter
        self._pack_counter += 1

    def forward_clip(self, frames, t_frames, events=None):
        """Emulate K consecutive packs in one call (extension; voxel_grid output only).

        Equivalent to ``[self(frames[k:k+1], t_frames[k]) for k in range(K)]`` (the loop
        test.py:115-121 runs around the reference emulator), but the K launches are enqueued
        back to back by one C call (`v2e_clip_forward`) with every host-side scalar prepared
        ahead, so the device never waits for Python between packs.

        frames: [K, T, H, W] on the CUDA device (float32 / uint8 / bfloat16), or a sequence of K
        tensors [1, T, H, W] / [T, H, W] of one shape and dtype; t_frames: sequence of K stamp
        tensors ([1,2] or [1,T]).  events: optional list of 2K torch.cuda.Event
        (timing enabled) recorded before / after each pack's launch.
        Returns (voxels [K, num_bins, H, W], num_events int64[K] on the device).  Nothing
        synchronises; call ``check_status()`` at a convenient point.
        """
        if self.output_mode != "voxel_grid":
            raise NotImplementedError("forward_clip serves output_mode='voxel_grid'")
        if self.noise_mode == "injected" or self.debug_maps:
            raise NotImplementedError("forward_clip: injected noise / debug maps go through forward()")
        if isinstance(frames, torch.Tensor):
            if frames.dim() != 4:
                raise ValueError("frames must be [num_packs, num_frames, height, width]")
            frames = list(frames.unbind(0))
        frames = list(frames)
        K = len(frames)
        if K < 1:
            raise ValueError("forward_clip needs K >= 1 packs of one shape and dtype")

        def norm(fr):
            # (per pack, inside the enqueue loop: the device starts on the first pack while the
            # host is still looking at the others)
            if not fr.is_cuda:
                raise RuntimeError("frames must be on the CUDA device (no CPU fallback)")
            if fr.dtype not in ops._FRAME_DT:
                fr = fr.float()
            fr = fr.contiguous()
            return fr.reshape(fr.shape[-3:])

        frames[0] = norm(frames[0])
        num_frames, height, width = frames[0].shape
        if len(t_frames) != K:
            raise ValueError("need one stamp tensor per pack")
        if not (2 <= num_frames <= _cabi.MAX_FRAMES):
            raise ValueError(f"num_frames must be in [2,{_cabi.MAX_FRAMES}]")
        dev = self._dev
        tr = self._clip_trace
        if tr is not None:
            import time as _t
            tr.append(("entry+norm", _t.perf_counter()))
        pl = self._get_launch(height, width, frames[0].dtype)
        packs = (_cabi.Pack * K)()
        out = torch.empty((K, self.num_bins, height, width), dtype=self.voxel_dtype, device=dev)
        # one 16-byte record per pack: num_events (i64), status (i32), pad
        meta = torch.zeros((K, 2), dtype=torch.int64, device=dev)
        ostride = out.stride(0) * out.element_size()
        evp = None
        if events is not None:
            if len(events) != 2 * K:
                raise ValueError("need 2 events per pack")
            evp = (C.c_void_p * (2 * K))()
            for i, e in enumerate(events):
                e.record()                      # forces creation of the cudaEvent_t
                evp[i] = e.cuda_event
        # Packs are prepared and enqueued in growing chunks, so that the device works on chunk i
        # while the host prepares chunk i+1.  A chunk whose stamps are host float64 rows (what the
        # reference's readers produce) and whose frames are contiguous is prepared by ONE C call
        # (`v2e_host_clip_prepare`, well under a microsecond per pack); anything else goes pack
        # by pack through `_prep_pack` (the torch expressions of v2e_model.py:303-318, ~10 us).
        if tr is not None:
            tr.append(("alloc+events", _t.perf_counter()))
        stream = torch.cuda.current_stream(dev).cuda_stream
        out_ptr, meta_ptr = out.data_ptr(), meta.data_ptr()
        dt0, shp0 = frames[0].dtype, frames[0].shape
        f64 = torch.float64
        sizeof_pack, sizeof_vp = C.sizeof(_cabi.Pack), C.sizeof(C.c_void_p)
        # chunks: one pack first (the device starts after ~70 us of host time), then as many as a
        # launch chains when the C preparation serves the clip; 1, 1, 2, 3, 5, 8, 13, 16, ... packs
        # when the packs are prepared one by one (the host then needs about a fifth of the
        # device's time per pack)
        k0, chunk, prev = 0, 1, 0
        ramp = self.clip_ramp
        if ramp is None:
            tf0 = t_frames[0]
            if (tf0.dim() == 2 and tf0.dtype == f64 and not tf0.is_cuda and tf0.shape[1] != 2
                    and tf0.shape[1] >= num_frames):
                ramp = (1, _cabi.MAX_SEQ_PER_LAUNCH)
        ci = 0
        while k0 < K:
            if ramp is not None:
                chunk = ramp[min(ci, len(ramp) - 1)]
                ci += 1
            k1 = min(K, k0 + chunk)
            n = k1 - k0
            pk0 = C.cast(C.addressof(packs) + k0 * sizeof_pack, C.POINTER(_cabi.Pack))
            # ---- the chunk in one C call, when its inputs have the usual layout
            fptrs, rows, n_st = [], [], t_frames[k0].shape[-1] if t_frames[k0].dim() == 2 else 0
            fast = n_st >= num_frames and n_st != 2
            if fast:
                for k in range(k0, k1):
                    fr, tf = frames[k], t_frames[k]
                    if not (fr.is_cuda and fr.dtype == dt0 and fr.shape[-3:] == shp0
                            and fr.is_contiguous() and tf.dtype == f64 and tf.dim() == 2
                            and tf.shape[1] == n_st and tf.shape[0] >= 1 and not tf.is_cuda
                            and tf.is_contiguous() and not tf.requires_grad):
                        fast = False
                        break
                    fptrs.append(fr.data_ptr())
                    rows.append(tf.data_ptr())
            if fast:
                fast = self._prep_chunk_c(pk0, n, (C.c_void_p * n)(*fptrs), (C.c_void_p * n)(*rows),
                                          n_st, num_frames, height, width,
                                          out_ptr + k0 * ostride, ostride, meta_ptr + 16 * k0)
            failed = None
            if not fast:
                k = k0
                try:
                    for k in range(k0, k1):
                        if k:
                            frames[k] = norm(frames[k])
                            if frames[k].shape != shp0 or frames[k].dtype != dt0:
                                raise ValueError("forward_clip needs K >= 1 packs of one shape and dtype")
                        self._prep_pack(packs[k], frames[k], t_frames[k], out_ptr + k * ostride,
                                        meta_ptr + 16 * k)
                except Exception as e:
                    # like the caller's loop around the reference (test.py:115-121): the packs
                    # before the offending one are emulated, then the error surfaces
                    failed, n = e, k - k0
            if failed is not None and n == 0:
                self._clip_keep = (frames, packs)
                raise failed
            ev0 = None
            if evp is not None:
                ev0 = C.cast(C.addressof(evp) + 2 * k0 * sizeof_vp, C.POINTER(C.c_void_p))
            if tr is not None:
                tr.append((f"prep {n}", _t.perf_counter()))
            rc = pl._lib.v2e_clip_forward(C.byref(pl.cfg), pk0, n, self._workspace.data_ptr(),
                                          self._workspace.numel(), stream, ev0)
            _cabi.check(rc, "v2e_clip_forward")
            if tr is not None:
                tr.append((f"enqueue {n}", _t.perf_counter()))
            if failed is not None:
                self._clip_keep = (frames, packs)
                raise failed
            k0, chunk, prev = k1, min(chunk + prev if prev else 1, 16), chunk
        self._clip_keep = (frames, packs)       # keep the inputs alive until the stream is done
        meta32 = meta.view(torch.int32)         # [K, 4]: num_events lo/hi, status, pad
        self._last_meta = meta32[-1]
        self._clip_meta = meta32
        num_events = meta[:, 0]
        self.num_events = num_events
        return out, num_events

    def forward_host(self, frames, t_frames):
        """``forward`` for frames that live in HOST memory (extension; what a caller does with the
        frames of a video reader, test.py:115-116: ``frames.to(device)`` then the emulator).

        frames: CPU tensor [1, num_frames, H, W], float32 / uint8 / bfloat16, ideally pinned.
        The pack is staged through two device buffers on a copy stream, so that the copy of pack
        k+1 runs while the kernel of pack k does (with ``lazy_num_events`` nothing blocks the
        host).  Frame 0 of every pack but the first after ``reset()`` is never read by the
        emulator (v2e_model.py:263,350: the loops start at n = 1; the first pack takes its
        initial state from frame 0, :325-327), so it is not copied: T-1 frames per pack cross the
        bus.  Returns what ``forward`` returns.
        """
        if frames.is_cuda:
            return self.forward(frames, t_frames)
        if frames.dim() != 4 or frames.size(0) != 1:
            raise ValueError("frames must be [batch_size=1, num_frames, height, width]")
        if frames.dtype not in ops._FRAME_DT:
            frames = frames.float()
        dev = self._dev
        st = getattr(self, "_host_stage", None)
        if st is None or st["shape"] != tuple(frames.shape) or st["dtype"] != frames.dtype:
            st = self._host_stage = {
                "shape": tuple(frames.shape), "dtype": frames.dtype, "k": 0,
                "buf": [torch.empty(frames.shape, dtype=frames.dtype, device=dev) for _ in range(2)],
                "ready": [torch.cuda.Event() for _ in range(2)],
                "freed": [torch.cuda.Event() for _ in range(2)],
                "stream": torch.cuda.Stream(dev)}
        b = st["k"] & 1
        st["k"] += 1
        lo = 0 if self.base_log_frame is None else 1
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(st["stream"]):
            st["stream"].wait_event(st["freed"][b])
            st["buf"][b][:, lo:].copy_(frames[:, lo:], non_blocking=True)
            st["ready"][b].record(st["stream"])
        cur.wait_event(st["ready"][b])
        out = self.forward(st["buf"][b], t_frames)
        st["freed"][b].record(cur)
        return out

    def forward(self, frames, t_frames):
        """Compute events based on a sequence of frames (v2e_model.py:290-536).

        frames: [1, num_frames, H, W] on the CUDA device; float32 (the reference contract),
            uint8 or bfloat16 (lossless for the integer-valued 0..255 frames every reference
            reader produces).
        t_frames: [1, 2] or [1, num_frames] float seconds.
        Returns (events, num_events): events is the normalised voxel grid [1, num_bins, H, W]
        or raw rows [N, 5] = [t, x, y, polarity, batch]; num_events is a python int.
        """
        if frames.dim() != 4:
            raise ValueError("frames must be [batch_size=1, num_frames, height, width]")
        batch_size, num_frames, height, width = frames.size()
        if batch_size != 1:
            raise NotImplementedError(
                "batch_size must be 1 per EventEmulator (the reference mixes sequences for "
                "B>1); use EventEmulatorBatch for independent sequences")
        if not frames.is_cuda:
            raise RuntimeError("frames must be on the CUDA device (no CPU fallback)")
        if frames.dtype not in ops._FRAME_DT:
            frames = frames.float()
        frames = frames.contiguous()
        if not (2 <= num_frames <= _cabi.MAX_FRAMES):
            raise ValueError(f"num_frames must be in [2,{_cabi.MAX_FRAMES}]")
        self.frame_counter += num_frames
        first = self.base_log_frame is None
        if first:
            ext = getattr(self, "_planes_external", False)
            keep = (self.pos_thres, self.neg_thres, self.noise_rate_array) if ext else None
            self._init(height, width)
            if ext:
                self.pos_thres, self.neg_thres, self.noise_rate_array = keep
        # per-pack scalars on the host; raises ValueError on non-increasing time (:335)
        sc = pack_scalars(t_frames, num_frames, self.num_bins, self._refractory_host,
                          self.cutoff_hz, self.ql, self.qs, self.shot_noise_rate_hz,
                          self._t_prev, first, out=self._scalars)
        self.t_previous = self._t_prev.value

        pl = self._get_launch(height, width, frames.dtype)
        pk = pl.packs[0]
        P = height * width
        dev = self._dev
        pk.sc = sc
        pk.frames = frames.data_ptr()
        pk.base_log_frame = self.base_log_frame.data_ptr()
        pk.lp_log_frame0 = self.lp_log_frame0.data_ptr()
        pk.timestamp_mem = self.timestamp_mem.data_ptr()
        pk.pos_thres = self.pos_thres.data_ptr()
        pk.neg_thres = self.neg_thres.data_ptr()
        pk.noise_rate = ops._ptr(self.noise_rate_array)
        # one zeroed scratch tensor: [0:2] num_events (int64), [2] status, [4:4+T] max_iters
        meta = torch.zeros(4 + num_frames, dtype=torch.int32, device=dev)
        num_events = meta[0:2].view(torch.int64)
        pk.num_events = meta.data_ptr()
        pk.status = meta.data_ptr() + 8
        keep = [frames]
        if self.output_mode == "raw":
            cap = self.raw_capacity or 4 * P * (num_frames - 1)
            out = torch.empty((cap, 5), dtype=torch.float32, device=dev)
            pk.raw_events, pk.raw_capacity, pk.voxel = out.data_ptr(), cap, None
        else:
            out = torch.empty((1, self.num_bins, height, width), dtype=self.voxel_dtype,
                              device=dev)
            pk.voxel, pk.raw_events, pk.raw_capacity = out.data_ptr(), None, 0
        if self.noise_mode == "injected":
            if self.leak_rate_hz > 0:
                lk = self.inject_leak_randn.to(dev, torch.float32).reshape(num_frames - 1, P)
                lk = lk.contiguous()
                pk.leak_randn = lk.data_ptr()
                keep.append(lk)
            if self.shot_noise_rate_hz > 0:
                sh = self.inject_shot_rand.to(dev, torch.float32).reshape(num_frames - 1, -1, P)
                sh = sh.contiguous()
                pk.shot_rand, pk.shot_imax = sh.data_ptr(), sh.size(1)
                keep.append(sh)
        pk.philox_seed = self.philox_seed
        pk.philox_offset = self._pack_counter
        self._pack_counter += 1
        mx = meta[4:]
        pk.max_iters = meta.data_ptr() + 16
        dbg = {}
        if self.debug_maps:
            dbg = dict(pol=torch.zeros((num_frames - 1, P), dtype=torch.int8, device=dev),
                       cnt=torch.zeros((num_frames - 1, P), dtype=torch.int32, device=dev),
                       final=torch.zeros((num_frames - 1, P), dtype=torch.int32, device=dev))
            pk.dbg_pol, pk.dbg_cnt = dbg["pol"].data_ptr(), dbg["cnt"].data_ptr()
            pk.dbg_final = dbg["final"].data_ptr()
            if self.noise_mode == "philox":
                imax = self.debug_shot_imax
                dbg["leak_randn"] = torch.zeros((num_frames - 1, P), device=dev)
                dbg["shot"] = torch.zeros((num_frames - 1, imax, P), dtype=torch.uint8, device=dev)
                pk.dbg_leak_randn = dbg["leak_randn"].data_ptr()
                pk.dbg_shot, pk.dbg_shot_imax = dbg["shot"].data_ptr(), imax
        else:
            pk.dbg_pol = pk.dbg_cnt = pk.dbg_final = pk.dbg_leak_randn = pk.dbg_shot = None

        ops.emulate_pack(self._handle, frames, self.base_log_frame, self.lp_log_frame0,
                         self.timestamp_mem, out, num_events, self._workspace)

        self.last_debug = dict(max_iters=mx, **dbg)
        self._last_meta = meta
        if self.lazy_num_events and self.output_mode != "raw":
            # extension: keep the pipeline asynchronous; num_events is a 1-element device tensor
            # (call check_status() at a convenient sync point when integer_frames=True)
            self.num_events = num_events
            return out, num_events
        n, status = meta[0:4].view(torch.int64).tolist()   # python int, like the reference (sync)
        self._raise_on_status(status)
        self.num_events = n
        if self.output_mode == "raw":
            rows = out[:min(n, out.size(0))]
            if rows.size(0) > 1:
                # v2e_model.py:527-534: time order; ties are broken by pixel for determinism
                # (t >= 0, so its bit pattern is monotone)
                tbits = rows[:, 0].contiguous().view(torch.int32).to(torch.int64)
                key = (tbits << 32) | (rows[:, 2] * width + rows[:, 1]).to(torch.int64)
                rows = rows[torch.argsort(key)]
            return rows, n
        return out, n


class _FrameView:
    """[T, H, W] slice of a contiguous device tensor, by address (no torch indexing op on the
    enqueue path)."""
    __slots__ = ("_ptr", "shape")

    def __init__(self, ptr, shape):
        self._ptr, self.shape = ptr, shape

    def data_ptr(self):
        return self._ptr


class EventEmulatorBatch(nn.Module):
    """S independent sequences emulated together (extension): one `EventEmulator` (batch size 1,
    the only well-defined batch semantics of the reference, SURVEY.md trap T1) per sequence, but
    up to 16 sequences share ONE launch of `v2e_pack_forward` (their CTAs are grouped per
    sequence; barriers, maxima and normalisation statistics stay per sequence).  Results are
    those of running the S emulators one after another.

    forward(frames, t_frames): frames [S, T, H, W] (or a sequence of S tensors [1,T,H,W] /
    [T,H,W]) on the CUDA device, t_frames [S, 2] / [S, T] (or a sequence of S [1,·] tensors).
    Returns (voxels [S, num_bins, H, W], num_events int64[S] on the device); nothing
    synchronises (call `check_status()`).  voxel_grid output, PHILOX / no noise only.
    """

    MAX_PER_LAUNCH = _cabi.MAX_SEQ_PER_LAUNCH

    def __init__(self, n_sequences, output_mode="voxel_grid", **kwargs):
        super().__init__()
        if output_mode != "voxel_grid":
            raise NotImplementedError("EventEmulatorBatch serves output_mode='voxel_grid'")
        if n_sequences < 1:
            raise ValueError("n_sequences must be >= 1")
        self.emulators = nn.ModuleList([EventEmulator(output_mode, **kwargs)
                                        for _ in range(n_sequences)])
        for i, em in enumerate(self.emulators):
            if em.noise_mode == "injected" or em.debug_maps:
                raise NotImplementedError("injected noise / debug maps: use EventEmulator")
            # one Philox key per sequence (the stream is indexed by pixel and pack only)
            em.philox_seed = (em.philox_seed + i * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        self._groups = None
        self._key = None
        self._meta = None

    def __len__(self):
        return len(self.emulators)

    def __getitem__(self, i):
        return self.emulators[i]

    def reset(self, indices=None):
        for i in (range(len(self.emulators)) if indices is None else indices):
            self.emulators[i].reset()

    def _get_groups(self, H, W, dtype):
        key = (H, W, dtype)
        if self._key != key:
            S, g = len(self.emulators), self.MAX_PER_LAUNCH
            self._groups = []
            for a in range(0, S, g):
                n = min(g, S - a)
                pl = ops.PackLaunch(self.emulators[a]._config(H, W, dtype), n)
                ws = torch.zeros(pl.workspace_bytes, dtype=torch.uint8,
                                 device=self.emulators[a]._dev)
                self._groups.append((a, n, pl, ws))
            self._key = key
        return self._groups

    @property
    def resolved_engine(self):
        if not self._groups:
            return None
        return {_cabi.ENGINE_STEPWISE: "stepwise", _cabi.ENGINE_PERSISTENT: "persistent",
                _cabi.ENGINE_RUNAHEAD: "runahead",
                _cabi.ENGINE_FIXPOINT: "fixpoint"}.get(self._groups[0][2].resolve_engine())

    def check_status(self):
        if self._meta is not None:
            meta, self._meta = self._meta, None
            EventEmulator._raise_on_status(int(meta[:, 2].max().item()))

    def forward_clip(self, frames, t_frames, events=None):
        """K consecutive steps of the batch in one call (extension): step k is what
        ``forward(frames[k], t_frames[k])`` does - one launch per group of up to 16 sequences -
        but the launches of a group are enqueued back to back by one C call
        (`v2e_batch_clip_forward`) with every host-side scalar prepared ahead.

        frames: [K, S, T, H, W] on the CUDA device, or a sequence of K tensors [S, T, H, W];
        t_frames: a sequence of K stamp tensors, each [S, 2] / [S, T] (one row per sequence) or
        [1, 2] / [1, T] (one time base for the whole batch: its scalars are computed once per
        step and shared by the sequences that are in the same time state).
        events: optional list of 2K torch.cuda.Event per GROUP of 16 sequences (a flat list of
        2K * n_groups), recorded before / after each step's launch.
        Returns (voxels [K, S, num_bins, H, W], num_events int64 [K, S] on the device); nothing
        synchronises (call `check_status()`).
        """
        S = len(self.emulators)
        frames = list(frames)
        K = len(frames)
        if K < 1 or len(t_frames) != K:
            raise ValueError("need K >= 1 steps and one stamp tensor per step")
        f0 = frames[0]
        if f0.dim() != 4 or f0.size(0) != S:
            raise ValueError("each step's frames must be [n_sequences, num_frames, height, width]")
        if not f0.is_cuda:
            raise RuntimeError("frames must be on the CUDA device (no CPU fallback)")
        dtype = f0.dtype if f0.dtype in ops._FRAME_DT else torch.float32
        T, H, W = f0.shape[1:]
        if not (2 <= T <= _cabi.MAX_FRAMES):
            raise ValueError(f"num_frames must be in [2,{_cabi.MAX_FRAMES}]")
        em0 = self.emulators[0]
        dev = em0._dev
        out = torch.empty((K, S, em0.num_bins, H, W), dtype=em0.voxel_dtype, device=dev)
        meta = torch.zeros((K, S, 4), dtype=torch.int32, device=dev)
        ostride = out.stride(1) * out.element_size()
        out_ptr, meta_ptr = out.data_ptr(), meta.data_ptr()
        stream = torch.cuda.current_stream(dev).cuda_stream
        groups = self._get_groups(H, W, dtype)
        lib = _cabi.lib()
        packs = [(_cabi.Pack * (K * n))() for _, n, _, _ in groups]
        evp = None
        if events is not None:
            if len(events) != 2 * K * len(groups):
                raise ValueError("need 2 events per step and group")
            evp = (C.c_void_p * len(events))()
            for i, e in enumerate(events):
                e.record()                      # forces creation of the cudaEvent_t
                evp[i] = e.cuda_event
        keep = []
        # Steps are prepared and enqueued in chunks (see EventEmulator.forward_clip).  When the
        # stamps are host float64 rows and the frames contiguous, a sequence's share of a chunk
        # is prepared by one C call (`v2e_host_clip_prepare`, records interleaved by sequence).
        f64 = torch.float64
        sizeof_pack = C.sizeof(_cabi.Pack)
        tf0 = t_frames[0]
        n_st = tf0.shape[1] if tf0.dim() == 2 else 0
        fast_clip = (n_st >= T and n_st != 2 and tf0.dtype == f64 and not tf0.is_cuda)
        ramp = (1, _cabi.MAX_SEQ_PER_LAUNCH) if fast_clip else None
        fstride = T * H * W * {torch.float32: 4, torch.uint8: 1, torch.bfloat16: 2}[dtype]
        k0, chunk, prev, ci = 0, 1, 0, 0
        while k0 < K:
            if ramp is not None:
                chunk = ramp[min(ci, len(ramp) - 1)]
                ci += 1
            k1 = min(K, k0 + chunk)
            nst = k1 - k0
            fptrs, rows, fast, shared = [], [], fast_clip, t_frames[k0].size(0) == 1 and S > 1
            for k in range(k0, k1):
                fr = frames[k]
                if fr.shape != f0.shape or not fr.is_cuda:
                    raise ValueError("all steps share one frame shape and device")
                if fr.dtype != dtype:
                    fr = fr.to(dtype)
                fr = fr.contiguous()
                frames[k] = fr
                keep.append(fr)
                tf = t_frames[k]
                if fast and not (tf.dim() == 2 and tf.dtype == f64 and not tf.is_cuda
                                 and tf.shape[1] == n_st and tf.is_contiguous()
                                 and not tf.requires_grad
                                 and tf.shape[0] == (1 if shared else S)):
                    fast = False
                fptrs.append(fr.data_ptr())
                rows.append(tf.data_ptr() if fast else 0)
            if fast:
                fp_arr, row_arr = (C.c_void_p * nst)(*fptrs), (C.c_void_p * nst)(*rows)
            for g, (a, n, pl, ws) in enumerate(groups):
                pk = packs[g]
                for j in range(n):
                    i = a + j
                    em = self.emulators[i]
                    done = False
                    if fast:
                        pk0 = C.cast(C.addressof(pk) + (k0 * n + j) * sizeof_pack, C.POINTER(_cabi.Pack))
                        done = em._prep_chunk_c(pk0, nst, fp_arr, row_arr, n_st, T, H, W,
                                                out_ptr + (k0 * S + i) * ostride, S * ostride,
                                                meta_ptr + 16 * (k0 * S + i), 16 * S, i * fstride,
                                                0 if shared else i * n_st * 8, n)
                    if not done:      # pack by pack (raises what has to be raised)
                        for k in range(k0, k1):
                            tf = t_frames[k]
                            sh = tf.size(0) == 1 and S > 1
                            fr_i = _FrameView(fptrs[k - k0] + i * fstride, (T, H, W))
                            em._prep_pack(pk[k * n + j], fr_i, tf if sh else tf[i:i + 1],
                                          out_ptr + (k * S + i) * ostride, meta_ptr + 16 * (k * S + i))
            for g, (a, n, pl, ws) in enumerate(groups):
                pk0 = C.cast(C.addressof(packs[g]) + k0 * n * sizeof_pack, C.POINTER(_cabi.Pack))
                ev0 = None
                if evp is not None:
                    ev0 = C.cast(C.addressof(evp) + (g * 2 * K + 2 * k0) * C.sizeof(C.c_void_p),
                                 C.POINTER(C.c_void_p))
                rc = lib.v2e_batch_clip_forward(C.byref(pl.cfg), pk0, nst, n, ws.data_ptr(),
                                                ws.numel(), stream, ev0)
                _cabi.check(rc, "v2e_batch_clip_forward")
            k0, chunk, prev = k1, min(chunk + prev if prev else 1, 16), chunk
        self._keep = (keep, packs)
        self._meta = meta.reshape(K * S, 4)
        num_events = meta[:, :, 0:2].contiguous().view(torch.int64).reshape(K, S)
        return out, num_events

    def forward_host(self, frames, t_frames):
        """``forward`` for a step whose frames live in HOST memory (extension, the batch's
        counterpart of `EventEmulator.forward_host`): frames is a CPU tensor [S, T, H, W] (ideally
        pinned), staged through two device buffers on a copy stream so that the copy of step k+1
        runs while the kernels of step k do.  Frame 0 of a pack is read only by a sequence in its
        first pack after `reset()` (v2e_model.py:263,350,325-327): when no sequence is, the T-1
        other frames of every sequence cross the bus.  Returns what `forward` returns."""
        if frames.is_cuda:
            return self.forward(frames, t_frames)
        S = len(self.emulators)
        if frames.dim() != 4 or frames.size(0) != S:
            raise ValueError("frames must be [n_sequences, num_frames, height, width]")
        if frames.dtype not in ops._FRAME_DT:
            frames = frames.float()
        dev = self.emulators[0]._dev
        st = getattr(self, "_host_stage", None)
        if st is None or st["shape"] != tuple(frames.shape) or st["dtype"] != frames.dtype:
            st = self._host_stage = {
                "shape": tuple(frames.shape), "dtype": frames.dtype, "k": 0,
                "buf": [torch.empty(frames.shape, dtype=frames.dtype, device=dev) for _ in range(2)],
                "ready": [torch.cuda.Event() for _ in range(2)],
                "freed": [torch.cuda.Event() for _ in range(2)],
                "stream": torch.cuda.Stream(dev)}
        b = st["k"] & 1
        st["k"] += 1
        lo = 0 if any(em.base_log_frame is None for em in self.emulators) else 1
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(st["stream"]):
            st["stream"].wait_event(st["freed"][b])
            if lo == 0:
                st["buf"][b].copy_(frames, non_blocking=True)
            else:
                # (one contiguous copy per sequence: a strided [S, T-1, H, W] copy_ from pinned
                # memory goes through a synchronous staging path in torch)
                buf = st["buf"][b]
                for i in range(S):
                    buf[i, lo:].copy_(frames[i, lo:], non_blocking=True)
            st["ready"][b].record(st["stream"])
        cur.wait_event(st["ready"][b])
        out = self.forward(st["buf"][b], t_frames)
        st["freed"][b].record(cur)
        return out

    def forward(self, frames, t_frames):
        S = len(self.emulators)
        if isinstance(frames, torch.Tensor):
            if frames.dim() != 4 or frames.size(0) != S:
                raise ValueError("frames must be [n_sequences, num_frames, height, width]")
            frames = list(frames.unbind(0))
        if isinstance(t_frames, torch.Tensor):
            t_frames = [t_frames[i:i + 1] for i in range(t_frames.size(0))]
        if len(frames) != S or len(t_frames) != S:
            raise ValueError("need one pack of frames and one stamp tensor per sequence")
        flist = []
        for fr in frames:
            if not fr.is_cuda:
                raise RuntimeError("frames must be on the CUDA device (no CPU fallback)")
            if fr.dtype not in ops._FRAME_DT:
                fr = fr.float()
            fr = fr.contiguous()
            flist.append(fr.reshape(fr.shape[-3:]))
        if any(f.shape != flist[0].shape or f.dtype != flist[0].dtype for f in flist):
            raise ValueError("all sequences of a batch share one frame shape and dtype")
        T, H, W = flist[0].shape
        if not (2 <= T <= _cabi.MAX_FRAMES):
            raise ValueError(f"num_frames must be in [2,{_cabi.MAX_FRAMES}]")
        em0 = self.emulators[0]
        dev = em0._dev
        out = torch.empty((S, em0.num_bins, H, W), dtype=em0.voxel_dtype, device=dev)
        meta = torch.zeros((S, 4), dtype=torch.int32, device=dev)
        ostride = out.stride(0) * out.element_size()
        stream = torch.cuda.current_stream(dev).cuda_stream
        for a, n, pl, ws in self._get_groups(H, W, flist[0].dtype):
            for j in range(n):
                i = a + j
                self.emulators[i]._prep_pack(pl.packs[j], flist[i], t_frames[i],
                                             out.data_ptr() + i * ostride, meta.data_ptr() + 16 * i)
            pl.launch(ws, stream)
        self._keep = flist
        self._meta = meta
        num_events = meta[:, 0:2].contiguous().view(torch.int64).reshape(S)
        return out, num_events

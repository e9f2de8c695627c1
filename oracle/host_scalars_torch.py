"""Torch-literal restatement of the reference's per-pack scalar preparation.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Each expression below is the reference's own
torch expression, evaluated on CPU 0-dim tensors so that torch's type promotion and rounding
decide every bit (SURVEY.md traps T5, T6, T9).  The product's C implementation
(`v2e_host_pack_scalars`, v2e2v_b200/csrc/host_scalars.c) is tested for bit-equality against
this file in tests/test_host_scalars.py.

Reference lines restated (paths relative to the reference repo):
  v2e/v2e_model.py:303-318   t_float_frames, duration, time_frames, Tr
  v2e/v2e_model.py:259,268   IIR delta_time
  v2e/v2e_model.py:335-338   time-order ValueError
  v2e/v2e_model.py:353,521   delta_time = t_float_frames[n] - t_previous ; t_previous update
  v2e/emulator_utils.py:73-86  eps = inten01 * (delta_time / tau)
  v2e/emulator_utils.py:120    delta_leak = delta_time * ...
  v2e/emulator_utils.py:174    (shot_noise_rate_hz / 2) * delta_time / num_iters
"""
from __future__ import annotations

import math

import torch

from v2e2v_b200 import _cabi


def _as_f32_multiplier(zero_dim: torch.Tensor) -> float:
    """Value a 0-dim tensor takes when multiplied into a dimensioned f32 tensor."""
    return float((torch.ones(1, dtype=torch.float32) * zero_dim)[0])


def pack_scalars_torch(t_frames: torch.Tensor, num_frames: int, num_bins: int,
                       refractory_period_s: float, cutoff_hz: float, ql: float, qs: float,
                       shot_noise_rate_hz: float, t_previous, first_pack: bool):
    """Returns (PackScalars, new_t_previous[0-dim tensor]); raises ValueError like the reference.

    t_frames: CPU tensor [1,2] or [1,>=num_frames], float32 or float64.
    t_previous: 0-dim CPU tensor carried from the previous pack (ignored when first_pack).
    """
    t_frames = t_frames.detach().cpu()
    refr = torch.tensor(refractory_period_s, dtype=torch.float32)
    if t_frames.size()[1] == 2:
        t_float_frames = torch.linspace(start=t_frames[0, 0], end=t_frames[0, -1],
                                        steps=num_frames, dtype=torch.float32)
    else:
        t_float_frames = t_frames[0]
    duration = (num_bins - 1) / (num_frames - 1)
    time_frames = torch.linspace(start=0, end=duration * (num_frames - 1), steps=num_frames,
                                 dtype=torch.float32)
    Tr = torch.einsum('i,ij->ij', (num_bins - 1) * refr.repeat(1),
                      1 / (t_frames[:, -1:] - t_frames[:, 0:1])).float()
    if first_pack:
        t_previous = t_frames[0, 0]
    if t_float_frames[1] <= t_previous:
        raise ValueError("this frame time={} must be later than previous frame time={}".format(
            t_float_frames[1], t_previous))

    sc = _cabi.PackScalars()
    sc.Tr = float(Tr[0, 0])
    # ts_step = duration / num_iters == num_iters.reciprocal() * duration (f32 tensor * scalar)
    sc.duration = _as_f32_multiplier(torch.tensor(duration, dtype=torch.float64))
    sc.bins_m1 = float(torch.tensor(float(num_bins - 1), dtype=torch.float32))
    sc.first_pack = int(first_pack)
    sc.num_frames = num_frames
    mask = 0
    delta_iir = t_float_frames[1:] - t_float_frames[:-1]
    for n in range(1, num_frames):
        fr = sc.fr[n]
        fr.t_bin_prev = float(time_frames[n - 1])
        if cutoff_hz > 0:
            dt = delta_iir[n - 1]
            if ql > 0:
                tau0 = 1 / (math.pi * 2 * cutoff_hz * ql)
                fr.s_large = _as_f32_multiplier(dt / tau0)
            else:
                fr.s_large = math.inf
            if qs > 0:
                tau1 = 1 / (math.pi * 2 * cutoff_hz * qs)
                fr.s_small = _as_f32_multiplier(dt / tau1)
            else:
                fr.s_small = math.inf
        else:
            fr.s_large = fr.s_small = math.inf
        delta_time = t_float_frames[n] - t_previous
        fr.sdt = _as_f32_multiplier(delta_time)
        c = (shot_noise_rate_hz / 2) * delta_time
        # dtype of c / int32[B] decides the precision of the shot thresholds (trap T5)
        probe = c / torch.ones(1, dtype=torch.int32)
        if probe.dtype == torch.float64:
            mask |= 1 << n
        fr.shot_c = float(c)
        t_previous = t_float_frames[n]
    sc.shot_f64_mask = mask
    return sc, t_previous

"""Helpers shared by the `-m gpu` parity tests (CUDA path vs oracle / golden fixtures)."""
from __future__ import annotations

import numpy as np
import torch

from oracle.oracle import OracleEmulator
from tests.refharness import noise_arrays, synth_walk_pack, stamps

ENGINES = ("stepwise", "persistent", "runahead", "fixpoint")
# engines that share the (4-pixel group, frame) Philox stage-0 stream; the fixed-point engine
# indexes its stream by pixel (v2e_fixpoint.cu, noise_px)
LOCKSTEP_ENGINES = ("stepwise", "persistent", "runahead")

SD = dict(pl=1.5, ps=0.5, ql=1, qs=0, cutoff_hz=200)
CFG2 = dict(output_mode="voxel_grid", num_bins=5, pos_thres=0.6, neg_thres=0.6, sigma_thres=0.03,
            refractory_period_s=0.001, leak_rate_hz=0.1, shot_noise_rate_hz=1, **SD)
CFG1 = dict(output_mode="voxel_grid", num_bins=5, pos_thres=0.6, neg_thres=0.6, sigma_thres=0.03,
            refractory_period_s=0.001, leak_rate_hz=0.1, shot_noise_rate_hz=1,
            pl=1, ps=1, ql=1, qs=1, cutoff_hz=0)


def make_planes(seed, H, W, emu):
    """Seeded per-pixel planes with the reference's distributions (numpy; shared by both sides)."""
    rng = np.random.Generator(np.random.PCG64(seed))

    def thr(nominal):
        a = rng.normal(emu["pl"] * nominal, emu["sigma_thres"], (H, W)).astype(np.float32)
        b = rng.normal(emu["ps"] * nominal, emu["sigma_thres"], (H, W)).astype(np.float32)
        a[0::2, 0::2] = b[0::2, 0::2]
        return np.maximum(a, np.float32(0.01))
    pos, neg = thr(emu["pos_thres"]), thr(emu["neg_thres"])
    nra = np.exp(np.float32(np.log(10) * 0.1) * rng.standard_normal((H, W), dtype=np.float32))
    return pos, neg, nra.astype(np.float32)


def cuda_emulator(emu, engine, noise_mode=None, **kw):
    from v2e2v_b200 import EventEmulator
    noisy = emu["leak_rate_hz"] > 0 or emu["shot_noise_rate_hz"] > 0
    if noise_mode is None:
        noise_mode = "injected" if noisy else "off"
    return EventEmulator(device="cuda", noise_mode=noise_mode, engine=engine, **emu, **kw)


def assert_state_equal(em, orc, cutoff, msg=""):
    np.testing.assert_array_equal(em.base_log_frame.cpu().numpy().reshape(-1),
                                  orc.base_log_frame, err_msg="base " + msg)
    if cutoff:
        np.testing.assert_array_equal(em.lp_log_frame0.cpu().numpy().reshape(-1),
                                      orc.lp_log_frame0, err_msg="lp " + msg)
    np.testing.assert_array_equal(em.timestamp_mem.cpu().numpy().reshape(-1),
                                  orc.timestamp_mem, err_msg="tmem " + msg)


def video(seq, packs, T, H, W, fps=1000.0, sigma=6.0, dtype=torch.float64, two_col=False):
    prev = None
    for k in range(packs):
        fr = synth_walk_pack(seq, k, T, H, W, prev, sigma)
        prev = fr[:, -1:].clone()
        yield k, fr, stamps(k, T, fps, dtype, two_col)


def shot_rand_from_decisions(shot_u8: np.ndarray, pol: np.ndarray) -> np.ndarray:
    """Turn the kernel's exported shot decisions [T-1,I,P] into a rand tensor that makes the
    reference's comparisons (emulator_utils.py:202-205) reproduce them."""
    r = np.full(shot_u8.shape, 0.5, np.float32)
    on = (shot_u8 != 0) & (pol[:, None, :] > 0)
    off = (shot_u8 != 0) & (pol[:, None, :] < 0)
    r[on] = np.float32(0.99999994)
    r[off] = np.float32(0.0)
    return r

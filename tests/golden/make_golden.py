"""Generate the golden fixtures by running the UNMODIFIED reference (needs /root/reference).

    python tests/golden/make_golden.py            # writes tests/golden/<case>.npz

For each case in cases.py the reference `EventEmulator` (v2e/v2e_model.py:32) is driven on CPU
over several packs of seeded synthetic video with *injected* noise (torch.randn / torch.rand are
patched to serve pre-generated arrays, see tests/refharness.NoiseInjector).  Stored per case:

  inputs : frames (uint8), stamps, per-sequence planes (pos/neg thresholds, noise-rate array) as
           the reference's own _init produced them, the noise seeds + checksums of the
           regenerated noise arrays, and the host scalar tables (v2e_pack_scalars bytes).
  outputs: after every pack  base_log_frame, lp_log_frame0, timestamp_mem (bit-exact targets),
           the returned voxel grid / raw event rows, and num_events.

Also writes voxel_cases.npz: reference outputs of events_to_voxel_grid / event_preprocess
(utils/event_process.py) on seeded event lists.
"""
from __future__ import annotations

import ctypes
import json
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import refharness as rh                      # noqa: E402
from tests.golden.cases import CASES, SHOT_IMAX         # noqa: E402
from oracle.host_scalars_torch import pack_scalars_torch  # noqa: E402


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def run_case(name: str, case: dict) -> dict:
    v2e_model, _, _ = rh.import_reference()
    emu_kw = dict(case["emu"])
    H, W, T, packs, fps = case["H"], case["W"], case["T"], case["packs"], case["fps"]
    P = H * W
    kind, dt = case["stamps"]
    dtype = torch.float64 if dt == "f64" else torch.float32
    leak = emu_kw["leak_rate_hz"] > 0
    shot = emu_kw["shot_noise_rate_hz"] > 0
    seed0 = zlib.crc32(name.encode()) & 0xFFFF

    torch.manual_seed(seed0)
    em = v2e_model.EventEmulator(device="cpu", **emu_kw)
    em.reset()
    rng = np.random.Generator(np.random.PCG64(seed0))
    nra_z = rng.standard_normal(P, dtype=np.float32) if leak else None
    inj = rh.NoiseInjector(nra_z)

    out = {"config": json.dumps(dict(case=case, seed0=seed0, shot_imax=SHOT_IMAX))}
    prev_last = None
    t_prev = None
    for k in range(packs):
        frames = rh.synth_walk_pack(seed0, k, T, H, W, prev_last, sigma=case.get("sigma", 6.0))
        prev_last = frames[:, -1:].clone()
        t_frames = rh.stamps(k, T, fps, dtype, two_col=(kind == "two"))
        fr_in = frames
        if case.get("drop_first") and k > 0:
            fr_in = frames[:, 1:]            # T-1 frames, T stamps (video_readers.py:85-101)
        nseed = seed0 * 100 + k
        lk, sh = rh.noise_arrays(nseed, fr_in.shape[1], P, SHOT_IMAX, leak, shot)
        inj.begin_pack(lk, sh)
        first = em.base_log_frame is None
        with inj.patched():
            ev, nev = em(fr_in.clone(), t_frames)
        sc, t_prev = pack_scalars_torch(
            t_frames, fr_in.shape[1], emu_kw["num_bins"], emu_kw["refractory_period_s"],
            emu_kw["cutoff_hz"], emu_kw["ql"], emu_kw["qs"], emu_kw["shot_noise_rate_hz"],
            t_prev, first)
        if k == 0:
            out["pos_thres"] = em.pos_thres.numpy().reshape(P).copy()
            out["neg_thres"] = em.neg_thres.numpy().reshape(P).copy()
            if leak:
                out["noise_rate"] = em.noise_rate_array.numpy().reshape(P).copy()
        pre = f"p{k}_"
        out[pre + "frames"] = fr_in[0].numpy().astype(np.uint8)
        assert np.array_equal(out[pre + "frames"].astype(np.float32), fr_in[0].numpy())
        out[pre + "t_frames"] = t_frames.numpy()
        out[pre + "noise_seed"] = np.int64(nseed)
        if leak:
            out[pre + "leak_crc"] = np.uint32(crc(lk))
        if shot:
            out[pre + "shot_crc"] = np.uint32(crc(sh))
        out[pre + "scalars"] = np.frombuffer(ctypes.string_at(ctypes.addressof(sc),
                                                               ctypes.sizeof(sc)), np.uint8).copy()
        out[pre + "base"] = em.base_log_frame.numpy().reshape(P).copy()
        out[pre + "lp"] = em.lp_log_frame0.numpy().reshape(P).copy()
        out[pre + "tmem"] = em.timestamp_mem.numpy().reshape(P).copy()
        out[pre + "num_events"] = np.int64(nev)
        if emu_kw["output_mode"] == "raw":
            rows = ev.numpy().reshape(-1, 5).astype(np.float32)
            order = np.lexsort((rows[:, 1], rows[:, 2], rows[:, 0]))   # canonical (t, y, x)
            out[pre + "events"] = rows[order]
        else:
            out[pre + "voxel"] = ev.numpy().reshape(emu_kw["num_bins"], P).astype(np.float32)
        print(f"  {name} pack {k}: num_events={nev}")
    return out


def voxel_cases() -> dict:
    """Reference outputs of the standalone voxel functions (utils/event_process.py)."""
    _, _, ep = rh.import_reference()
    out = {}
    rng = np.random.Generator(np.random.PCG64(77))
    specs = [("a", 2000, 5, 24, 16, np.float64), ("b", 500, 5, 20, 12, np.float32),
             ("c", 1, 5, 8, 8, np.float64), ("d", 300, 3, 10, 6, np.float64),
             ("e", 0, 5, 8, 8, np.float64), ("f", 5000, 5, 12, 10, np.float64)]
    for tag, n, nb, W, H, dt in specs:
        t = np.sort(rng.random(n)) * 0.05 + 10.0
        if tag == "f":                       # duplicated stamps + dense pixels (hot-pixel filter)
            t = np.round(t, 3)
        ev = np.stack([t, rng.integers(0, W, n).astype(np.float64),
                       rng.integers(0, H, n).astype(np.float64),
                       rng.integers(0, 2, n).astype(np.float64)], 1).astype(dt).reshape(n, 4)
        out[f"{tag}_events"] = ev.copy()
        out[f"{tag}_dims"] = np.array([nb, W, H], np.int64)
        vox = ep.events_to_voxel_grid(ev.copy(), nb, W, H)
        out[f"{tag}_voxel_np"] = vox.copy()
        out[f"{tag}_norm_np"] = np.asarray(ep.event_preprocess(vox.copy(), 'std', False))
        out[f"{tag}_norm_np_hot"] = np.asarray(ep.event_preprocess(vox.copy(), 'std', True))
        out[f"{tag}_maxmin_np"] = np.asarray(ep.event_preprocess(vox.copy(), 'maxmin', False))
        evt = torch.from_numpy(ev.copy())
        vt = ep.events_to_voxel_grid_pytorch(evt, nb, W, H)
        out[f"{tag}_voxel_torch"] = vt.numpy().copy()
        out[f"{tag}_norm_torch"] = ep.event_preprocess_pytorch(vt.clone(), 'std', False).numpy()
        out[f"{tag}_norm_torch_hot"] = ep.event_preprocess_pytorch(vt.clone(), 'std', True).numpy()
    return out


def main():
    assert rh.have_reference(), "needs the reference checkout"
    torch.set_num_threads(1)
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        print(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(name, case))
    if not only or "voxel_cases" in only:
        np.savez_compressed(os.path.join(HERE, "voxel_cases.npz"), **voxel_cases())


if __name__ == "__main__":
    main()

"""Loader for tests/golden/<case>.npz (see tests/golden/make_golden.py)."""
from __future__ import annotations

import ctypes
import json
import os
import zlib

import numpy as np
import torch

from v2e2v_b200 import _cabi
from tests.refharness import noise_arrays

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


class GoldenCase:
    def __init__(self, name: str):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        meta = json.loads(str(self.z["config"]))
        self.case = meta["case"]
        self.shot_imax = meta["shot_imax"]
        self.emu = self.case["emu"]
        self.H, self.W = self.case["H"], self.case["W"]
        self.P = self.H * self.W
        self.packs = self.case["packs"]
        self.leak = self.emu["leak_rate_hz"] > 0
        self.shot = self.emu["shot_noise_rate_hz"] > 0
        self.raw = self.emu["output_mode"] == "raw"

    def planes(self):
        nra = self.z["noise_rate"] if self.leak else None
        return self.z["pos_thres"], self.z["neg_thres"], nra

    def pack_inputs(self, k: int):
        """-> frames u8 [T,H,W], t_frames torch [1,n], leak_randn, shot_rand, PackScalars."""
        pre = f"p{k}_"
        frames = self.z[pre + "frames"]
        t_frames = torch.from_numpy(self.z[pre + "t_frames"].copy())
        lk, sh = noise_arrays(int(self.z[pre + "noise_seed"]), frames.shape[0], self.P,
                              self.shot_imax, self.leak, self.shot)
        # the noise is regenerated, not stored: fail loudly if numpy's stream ever drifts
        if self.leak:
            assert crc(lk) == int(self.z[pre + "leak_crc"]), "leak noise stream drifted"
        if self.shot:
            assert crc(sh) == int(self.z[pre + "shot_crc"]), "shot noise stream drifted"
        sc = _cabi.PackScalars()
        raw = self.z[pre + "scalars"].tobytes()
        assert len(raw) == ctypes.sizeof(sc)
        ctypes.memmove(ctypes.addressof(sc), raw, len(raw))
        return frames, t_frames, lk, sh, sc

    def pack_outputs(self, k: int):
        pre = f"p{k}_"
        out = dict(base=self.z[pre + "base"], lp=self.z[pre + "lp"], tmem=self.z[pre + "tmem"],
                   num_events=int(self.z[pre + "num_events"]))
        if self.raw:
            out["events"] = self.z[pre + "events"]
        else:
            out["voxel"] = self.z[pre + "voxel"]
        return out


def scalars_equal(a: _cabi.PackScalars, b: _cabi.PackScalars) -> bool:
    """Bitwise comparison of the meaningful fields (frames 1..T-1)."""
    def key(s):
        T = s.num_frames
        head = (np.float32(s.Tr).tobytes(), np.float32(s.duration).tobytes(),
                np.float32(s.bins_m1).tobytes(), s.first_pack, s.shot_f64_mask, T)
        fr = tuple((np.float32(s.fr[n].t_bin_prev).tobytes(), np.float32(s.fr[n].s_large).tobytes(),
                    np.float32(s.fr[n].s_small).tobytes(), np.float32(s.fr[n].sdt).tobytes(),
                    np.float64(s.fr[n].shot_c).tobytes()) for n in range(1, T))
        return head, fr
    return key(a) == key(b)

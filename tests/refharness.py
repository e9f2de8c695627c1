"""Harness around the UNMODIFIED reference (only usable where /root/reference exists).

Used by tests/golden/make_golden.py (fixture generation) and by the CPU-only tests that compare
the oracle / host code with the live reference.  Nothing here is imported on the GPU box at
run time: `have_reference()` gates every use.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("V2E2V_REFERENCE_ROOT", "/root/reference")


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "v2e", "v2e_model.py"))


def import_reference():
    """Import the reference's modules (matplotlib is display-only there and absent here)."""
    for m in ("matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import v2e.v2e_model as v2e_model          # noqa: E402
    import v2e.emulator_utils as emulator_utils  # noqa: E402
    import utils.event_process as event_process  # noqa: E402
    return v2e_model, emulator_utils, event_process


class NoiseInjector:
    """Serves the reference's torch.randn / torch.rand calls from caller-supplied arrays.

    Call order inside the reference (v2e/v2e_model.py): the first randn after reset() is the
    noise_rate_array draw (:232); then per frame one randn (leak, emulator_utils.py:113) and
    one rand of shape [max_iters,B,1,H,W] (shot, emulator_utils.py:191).
    """

    def __init__(self, nra_z=None):
        self.nra_z = nra_z
        self.nra_served = nra_z is None
        self.leak = None
        self.shot = None
        self.n_leak = 0
        self.n_shot = 0

    def begin_pack(self, leak_randn, shot_rand):
        self.leak, self.shot = leak_randn, shot_rand
        self.n_leak = self.n_shot = 0

    def randn(self, *size, dtype=None, device=None, **kw):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else size
        if not self.nra_served:
            self.nra_served = True
            return torch.from_numpy(self.nra_z.reshape(shape).copy())
        z = self.leak[self.n_leak]
        self.n_leak += 1
        return torch.from_numpy(z.reshape(shape).copy())

    def rand(self, *args, size=None, dtype=None, device=None, **kw):
        if size is None:
            size = args[0] if len(args) == 1 and not isinstance(args[0], int) else args
        size = [int(s) for s in size]
        r = self.shot[self.n_shot][: size[0]]
        assert r.shape[0] == size[0], f"shot_rand has {r.shape[0]} iterations, need {size[0]}"
        self.n_shot += 1
        return torch.from_numpy(r.reshape(size).copy())

    @contextlib.contextmanager
    def patched(self):
        orig_randn, orig_rand = torch.randn, torch.rand
        torch.randn, torch.rand = self.randn, self.rand
        try:
            yield self
        finally:
            torch.randn, torch.rand = orig_randn, orig_rand


def synth_walk_pack(seq: int, k: int, T: int, H: int, W: int, prev_last=None, sigma=6.0):
    """synth-walk-v1: integer-valued f32 frames [1,T,H,W] for pack k."""
    g = torch.Generator().manual_seed(1000 * seq + k)
    if prev_last is None:
        base = torch.rand(1, 1, H, W, generator=g) * 255
    else:
        base = prev_last
    frames = (base + torch.cumsum(torch.randn(1, T, H, W, generator=g) * sigma, dim=1))
    frames = frames.clamp(0, 255).round()
    if prev_last is not None:
        frames[:, 0] = prev_last[:, 0]
    return frames


def stamps(k: int, T: int, fps: float, dtype=torch.float64, two_col=False):
    t = torch.linspace(k * (T - 1) / fps, (k + 1) * (T - 1) / fps, T, dtype=torch.float64)
    t = t.to(dtype)
    if two_col:
        t = t[[0, -1]]
    return t[None]


def noise_arrays(seed: int, T: int, P: int, imax: int, leak: bool, shot: bool):
    """Deterministic noise draws for one pack (numpy PCG64; regenerated, never stored)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lk = rng.standard_normal((T - 1, P), dtype=np.float32) if leak else None
    sh = rng.random((T - 1, imax, P), dtype=np.float32) if shot else None
    return lk, sh

"""Probe: how many fixed-point passes does the speculative engine need on synth-walk-v1?

Per-pack: guess the per-frame maxima g[1..T-1], run every pixel independently under the guess,
collect the maxima m actually produced; consistent (m == g) <=> the guess was the true chain.
numpy f32, noise off, cfg3 parameters.  Test infrastructure only.
"""
import sys
import numpy as np
import torch

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (360, 640)
T, NB, FPS = 10, 5, 1000.0
f32 = np.float32
rng = torch.Generator().manual_seed(0)
P = H * W
yy, xx = np.divmod(np.arange(P), W)
small = ((yy | xx) & 1) == 0
C0 = 0.6
pos = np.where(small, 0.3, 0.9).astype(f32) + (torch.randn(P, generator=rng).numpy() * 0.03).astype(f32)
neg = np.where(small, 0.3, 0.9).astype(f32) + (torch.randn(P, generator=rng).numpy() * 0.03).astype(f32)
pos = np.maximum(pos, f32(0.01)); neg = np.maximum(neg, f32(0.01))
v = np.arange(256, dtype=np.float64)
LOG = np.where(v <= 20, v * np.log(20) / 20, np.log(np.maximum(v, 1))).astype(f32)
INT = ((v.astype(f32) + f32(20)) / f32(275)).astype(f32)
s_l = f32(0.001 * 2 * np.pi * 200)
dur = f32((NB - 1) / (T - 1))
Tr = f32(f32(4 * 0.001) * f32(1 / 0.009))
tfr = np.linspace(0, float(dur) * (T - 1), T).astype(f32)


def run_pack(frames, base, lp, tmem, g, mask=None):
    """all pixels under guessed maxima g[n]; returns new state + produced maxima"""
    base, lp, tmem = base.copy(), lp.copy(), tmem.copy()
    m = np.zeros(T, np.int64)
    for n in range(1, T):
        x = frames[n]
        L, I = LOG[x], INT[x]
        eps = np.minimum(np.where(small, f32(1), I * s_l), f32(1)).astype(f32)
        lp = ((f32(1) - eps) * lp + eps * L).astype(f32)
        diff = lp - base
        diff[np.abs(diff) <= f32(1e-6)] = 0
        pol = np.sign(diff).astype(f32)
        Cc = np.where(diff > 0, pos, np.where(diff < 0, neg, f32(0))).astype(f32)
        cnt = np.floor(np.abs(diff) / (Cc + f32(1e-9))).astype(np.int64)
        m[n] = cnt.max() if mask is None else (cnt[mask].max() if mask.any() else 0)
        mx = int(g[n])
        ni = max(mx, 1)
        step = f32(f32(1) / f32(ni)) * dur
        gate = Tr > step
        final = np.zeros(P, np.int64)
        for i in range(mx):
            msk = cnt >= i + 1
            ts = f32(tfr[n - 1] + f32(step * f32(i + 1)))
            if gate:
                msk = (np.where(msk, ts, f32(0)).astype(f32) - tmem) > Tr
                tmem[msk] = ts
            final += msk
        base = (base + (pol * final.astype(f32)) * Cc).astype(f32)
    return base, lp, tmem, m


x0 = (torch.rand(P, generator=rng).numpy() * 255)
walk = x0
prev_last = None
guess = np.full(T, 8, np.int64)
base = lp = tmem = None
tot_passes = []
for k in range(int(sys.argv[3]) if len(sys.argv) > 3 else 6):
    steps = torch.randn(T, P, generator=rng).numpy() * 6
    fr = np.clip(np.round(walk + np.cumsum(steps, 0)), 0, 255).astype(np.int64)
    if prev_last is not None:
        fr[0] = prev_last
    walk = fr[-1].astype(np.float64); prev_last = fr[-1]
    if base is None:
        base = LOG[fr[0]].copy(); lp = base.copy(); tmem = np.full(P, -Tr, f32)
    else:
        tmem = np.where(tmem > 0, tmem - f32(NB - 1), tmem).astype(f32)
        tmem = np.where(tmem < 0, -Tr, tmem).astype(f32)
    # hot pixels: first crossing is at least 4 thresholds (known without any guess)
    _, _, _, _ = 0, 0, 0, 0
    # hot pixels: some frame is at least theta thresholds away from the memorised brightness the
    # pack starts with (an upper bound of every event count the pixel can reach: known without g)
    theta = max(2, int(guess[1:].min()) // 2)
    hot = np.zeros(P, bool)
    l0 = lp.copy()
    for n in range(1, T):
        x = fr[n]; L, I = LOG[x], INT[x]
        eps = np.minimum(np.where(small, f32(1), I * s_l), f32(1)).astype(f32)
        l0 = ((f32(1) - eps) * l0 + eps * L).astype(f32)
        d = np.abs(l0 - base); Cc = np.where(l0 > base, pos, neg)
        hot |= d >= theta * Cc
    _, _, _, mh = run_pack(fr, base, lp, tmem, guess, mask=hot)
    gh = np.where(mh > 0, mh, guess)
    _, _, _, mfull = run_pack(fr, base, lp, tmem, gh)
    print(f"   hot pixels {hot.sum()} ({100*hot.mean():.2f}%): hot-only maxima {mh[1:].tolist()} -> full pass under them consistent: {(mfull[1:] == gh[1:]).all()}")
    g = guess.copy(); hist = []
    for it in range(12):
        b2, l2, t2, m = run_pack(fr, base, lp, tmem, g)
        hist.append((g[1:].tolist(), m[1:].tolist()))
        if (m[1:] == g[1:]).all():
            break
        g = m.copy()
    print(f"pack {k}: passes={len(hist)}")
    for gg, mm in hist:
        print("   guess", gg, "-> produced", mm)
    tot_passes.append(len(hist))
    base, lp, tmem, guess = b2, l2, t2, m
print("passes per pack:", tot_passes)

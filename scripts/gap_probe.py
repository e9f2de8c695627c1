"""Where a K-pack forward_clip spends its device span: start latency, launches, gaps (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import gpu_common as gc
from v2e2v_b200 import EventEmulator

H, W, T = 720, 1280, 10
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfg = dict(gc.CFG2)
torch.manual_seed(0)
em = EventEmulator(device="cuda", engine="auto", noise_mode="philox", lazy_num_events=True, **cfg)
g = torch.Generator(device="cuda").manual_seed(1)
base = torch.rand(1, 1, H, W, device="cuda", generator=g) * 255
vids = []
for k in range(6):
    fr = (base + torch.cumsum(torch.randn(1, T, H, W, device="cuda", generator=g) * 6, 1)).clamp(0, 255).round()
    fr[:, 0] = base[:, 0]
    base = fr[:, -1:].clone()
    vids.append(fr.contiguous())
sts = [torch.linspace(k * (T - 1) / 1000, (k + 1) * (T - 1) / 1000, T, dtype=torch.float64)[None] for k in range(4000)]
k = 0
for name, ramp in [("fib", None), ("16", (16,)), ("1,16", (1, 16)), ("fib", None), ("16", (16,))]:
    em.clip_ramp = ramp
    for rep in range(5):
        fl = [vids[(k + j) % 6] for j in range(K)]
        sl = [sts[k + j] for j in range(K)]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K)]
        for e in ev:
            e.record()
        torch.cuda.synchronize()
        if rep % 2:
            time.sleep(0.05)      # idle device before the call
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, nev = em.forward_clip(fl, sl, events=ev)
        e1.record()
        torch.cuda.synchronize()
        k += K
        t = [e0.elapsed_time(e) * 1e3 for e in ev] + [e0.elapsed_time(e1) * 1e3]
        # launches: runs of intervals starting at a non-empty one
        segs = []
        for j in range(K):
            a, b = t[2 * j], t[2 * j + 1]
            if b - a > 10 or not segs:
                segs.append([a, b, 1])
            else:
                segs[-1][1] = b; segs[-1][2] += 1
        if rep >= 1:
            txt = " ".join(f"[{a:.0f}-{b:.0f} n={n} {(b-a)/n:.1f}/pk]" for a, b, n in segs)
            print(f"K={K} ramp {name} idle={'y' if rep % 2 else 'n'}: total {t[-1]:.0f} us ({t[-1]/K:.1f}/pack): {txt}")
        del out

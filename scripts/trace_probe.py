"""Host-side marks of forward_clip (development aid)."""
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
for name, ramp in [("fib", None), ("16", (16,)), ("1,16", (1, 16))]:
    em.clip_ramp = ramp
    for rep in range(5):
        fl = [vids[(k + j) % 6] for j in range(K)]
        sl = [sts[k + j] for j in range(K)]
        torch.cuda.synchronize()
        em._clip_trace = [] if rep >= 2 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out, nev = em.forward_clip(fl, sl)
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        k += K
        if rep >= 2:
            marks = " ".join(f"{n}@{1e6*(t-t0):.0f}" for n, t in em._clip_trace)
            print(f"K={K} ramp {name}: span {1e3*e0.elapsed_time(e1):.0f} us, host call {1e6*(t1-t0):.0f}, wall {1e6*(t2-t0):.0f}: {marks}")
        del out

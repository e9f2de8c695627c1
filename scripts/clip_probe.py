"""Host-enqueue vs device time of forward_clip (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests import gpu_common as gc
from v2e2v_b200 import EventEmulator

H, W, T = (int(v) for v in os.environ.get("V2E_PROBE_HWT", "720,1280,10").split(","))
noise = sys.argv[1] if len(sys.argv) > 1 else "philox"
engine = sys.argv[2] if len(sys.argv) > 2 else "auto"
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cfg = dict(gc.CFG2)
if noise == "off":
    cfg.update(leak_rate_hz=0, shot_noise_rate_hz=0)
torch.manual_seed(0)
em = EventEmulator(device="cuda", engine=engine, noise_mode=noise, lazy_num_events=True, **cfg)
g = torch.Generator(device="cuda").manual_seed(1)
base = torch.rand(1, 1, H, W, device="cuda", generator=g) * 255
vids = []
for k in range(6):
    fr = (base + torch.cumsum(torch.randn(1, T, H, W, device="cuda", generator=g) * 6, 1)).clamp(0, 255).round()
    fr[:, 0] = base[:, 0]
    base = fr[:, -1:].clone()
    vids.append(fr.contiguous())
st = lambda k: torch.linspace(k * (T - 1) / 1000, (k + 1) * (T - 1) / 1000, T, dtype=torch.float64)[None]
k = 0
for rep in range(4):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K)] if rep % 2 == 0 else None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out, nev = em.forward_clip([vids[(k + j) % 6] for j in range(K)], [st(k + j) for j in range(K)], events=ev)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    k += K
    kern = sum(ev[2 * j].elapsed_time(ev[2 * j + 1]) for j in range(K)) / K * 1e3 if ev else float("nan")
    print(f"rep {rep} engine={em.resolved_engine} events={'y' if ev else 'n'}: host call {1e6*(t1-t0)/K:7.1f} us/pack, "
          f"device span {1e3*e0.elapsed_time(e1)/K:7.1f} us/pack, kernel mean {kern:7.1f} us, wall {1e6*(t2-t0)/K:7.1f} us/pack")
    del out
if os.environ.get("V2E_TIMELINE"):
    ws = em._workspace
    n = 160 * 32 * 8 + 2048 + 1024
    tl = ws[ws.numel() - n:].view(torch.int64).cpu().numpy()
    hist = tl[160 * 32:160 * 32 + 256].reshape(128, 2)
    rounds = tl[160 * 32 + 256:160 * 32 + 384]
    dur_all = (hist[:, 1] - hist[:, 0]) / 1e3
    for r in sorted(set(rounds[hist[:, 0] > 0].tolist())):
        m = (hist[:, 0] > 0) & (rounds == r)
        print(f"launches with {r} rounds (pre-pass included): {int(m.sum()):3d} of {int((hist[:,0]>0).sum())}, kernel mean {dur_all[m].mean():.1f} us")
    h = hist[hist[:, 0] > 0]
    h = h[np.argsort(h[:, 0])][-40:]
    dur = (h[:, 1] - h[:, 0]) / 1e3
    gap = (h[1:, 0] - h[:-1, 1]) / 1e3
    print("last launches: kernel (first CTA start -> last CTA exit) mean %.1f us, gap to next launch mean %.1f us (min %.1f max %.1f)"
          % (dur.mean(), gap.mean(), gap.min(), gap.max()))
    print("globaltimer span of these %d launches: %.1f us  -> %.1f us/launch by globaltimer (CUDA events say ~device span above)"
          % (len(h), (h[-1, 1] - h[0, 0]) / 1e3, (h[-1, 1] - h[0, 0]) / 1e3 / len(h)))

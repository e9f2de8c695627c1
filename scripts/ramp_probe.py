"""Device span of a K-pack forward_clip for different chunk ramps (development aid)."""
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
sts = [torch.linspace(k * (T - 1) / 1000, (k + 1) * (T - 1) / 1000, T, dtype=torch.float64)[None] for k in range(12000)]
k = 0
ramps = {"fib": None, "1,3,16": (1, 3, 16), "2,16": (2, 16), "4,16": (4, 16), "1,1,2,16": (1, 1, 2, 16), "16": (16,), "1,16": (1, 16)}
import gc as pygc
for name, ramp in list(ramps.items()) * 2:
    em.clip_ramp = ramp
    spans, hosts = [], []
    for rep in range(8):
        fl = [vids[(k + j) % 6] for j in range(K)]
        sl = [sts[k + j] for j in range(K)]
        pygc.collect(); pygc.disable()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out, nev = em.forward_clip(fl, sl)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        pygc.enable()
        k += K
        spans.append(1e3 * e0.elapsed_time(e1) / K)
        hosts.append(1e6 * (t1 - t0) / K)
        del out
    spans, hosts = np.array(spans[2:]), np.array(hosts[2:])
    print(f"K={K} ramp {name:10s}: device span us/pack median {np.median(spans):6.1f} min {spans.min():6.1f} max {spans.max():6.1f}   host call us/pack median {np.median(hosts):6.1f}")

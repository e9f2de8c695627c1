"""cProfile of forward_clip's host side (development aid)."""
import sys, time, os, cProfile, pstats
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
sts = [torch.linspace(k * (T - 1) / 1000, (k + 1) * (T - 1) / 1000, T, dtype=torch.float64)[None] for k in range(2000)]
k = 0
def one():
    global k
    fl = [vids[(k + j) % 6] for j in range(K)]
    sl = [sts[k + j] for j in range(K)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, nev = em.forward_clip(fl, sl)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    k += K
    return 1e6 * (t1 - t0)
for _ in range(5):
    one()
print("host call us:", [round(one()) for _ in range(8)])
# where inside the call: wall-clock marks through a patched time source
import v2e2v_b200.emulator as E
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    one()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

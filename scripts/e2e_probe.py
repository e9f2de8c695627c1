"""Where does the e2e step time go? (development aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from v2e2v_b200 import EventEmulator
dev = torch.device("cuda", 0)
H, W, T = 720, 1280, 10
gen = torch.Generator(device=dev).manual_seed(1234)
torch.manual_seed(0)
em = EventEmulator(device=dev, noise_mode="philox", engine="auto", lazy_num_events=True, **bench.EMU)
base = torch.rand(1, 1, H, W, device=dev, generator=gen) * 255
packs = []
for _ in range(5):
    fr, base = bench.synth_pack_device(gen, base, H, W)
    packs.append(fr)
pinned = [p.cpu().pin_memory() for p in packs]
copy_stream = torch.cuda.Stream(dev)
nev_host = torch.zeros(1, dtype=torch.int64).pin_memory()
stage = [torch.empty_like(pinned[0], device=dev) for _ in range(2)]
ready = [torch.cuda.Event() for _ in range(2)]
freed = [torch.cuda.Event() for _ in range(2)]

def step(kk, do_kernel=True, do_d2h=True, do_copy=True):
    b = kk & 1
    if do_copy:
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])
            stage[b].copy_(pinned[kk % 5], non_blocking=True)
            ready[b].record(copy_stream)
        torch.cuda.current_stream(dev).wait_event(ready[b])
    if do_kernel:
        _, n = em(stage[b], bench.stamps(kk))
        if do_d2h:
            nev_host[0:1].copy_(n, non_blocking=True)
    freed[b].record()

kk = 0
for name, kw in (("full", {}), ("no d2h", dict(do_d2h=False)), ("copy only", dict(do_kernel=False)),
                 ("kernel only", dict(do_copy=False)), ("full again", {})):
    for _ in range(5):
        step(kk, **kw); kk += 1
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = 0.0
    for _ in range(200):
        a = time.perf_counter(); step(kk, **kw); th += time.perf_counter() - a; kk += 1
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 200
    print(f"{name:12s}: {dt*1e6:7.1f} us/step (host enqueue {th/200*1e6:6.1f} us/step)")

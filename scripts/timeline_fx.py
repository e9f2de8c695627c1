"""Dump the per-CTA timeline of the fixed-point kernel (V2E_TIMELINE=1): development aid."""
import os, sys
os.environ["V2E_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import gpu_common as gc
from v2e2v_b200 import EventEmulator

H, W = int(sys.argv[1]), int(sys.argv[2])
noise = sys.argv[3] if len(sys.argv) > 3 else "philox"
cfg = dict(gc.CFG2)
if noise == "off":
    cfg.update(leak_rate_hz=0, shot_noise_rate_hz=0)
em = EventEmulator(device="cuda", engine="fixpoint", noise_mode=noise, lazy_num_events=True, **cfg)
T = 10
for k, fr, tfr in gc.video(1, 5, T, H, W):
    em(fr.cuda(), tfr)
    torch.cuda.synchronize()
    ws = em._workspace
    tl = ws[ws.numel() - (160 * 32 * 8 + 2048 + 1024):].view(torch.int64).cpu().numpy()[:160 * 32].reshape(160, 32)
    tl = tl[tl[:, 0] > 0]
    t0 = tl[:, 0].min()
    npass = int(tl[0, 30])
    us = lambda a: (a - t0) / 1e3
    print(f"pack {k}: CTAs {tl.shape[0]} rounds (pre-pass + full passes) {npass}  entries/tile mean {tl[:,28].mean():.0f} max {tl[:,28].max()} "
          f"spill mean {tl[:,29].mean():.1f} max {tl[:,29].max()} hot mean {tl[:,26].mean():.0f}")
    # NOTE: %globaltimer ticks are ~1.31 ns on this pool (calibrated against CUDA events in
    # scripts/clip_probe.py): multiply the figures below by 1.31 for microseconds
    rows = [("start", 0), ("setup done", 1), ("stream done", 8), ("  entries walked", 14), ("  offsets done", 15), ("post pass done", 2)]
    for ps in range(min(npass, 4)):
        rows += [(f"resolve{ps} done", 3 + 3 * ps)] + ([(f"arrived{ps}", 5 + 3 * ps)] if ps == 2 else []) + [(f"barrier{ps} done", 4 + 3 * ps)]
    rows += [("loop exit", 27), ("commit done", 31)]
    if os.environ.get("V2E_STAMPS"):
        c = tl[:, 24]
        print(f"   first full round, cold event lanes per tile: slow(shot) {(c & 0xffff).mean():.1f} slow(other) {((c >> 16) & 0xffff).mean():.1f} classified {((c >> 32) & 0xffff).mean():.1f} simple {((c >> 48) & 0xffff).mean():.1f}")
        rows += [("x: last task end (max warp)", 15), ("x: after flush (max warp)", 22), ("x: hot task 0 start", 16), ("x: hot task 0 end", 17),
                 ("x: first cold task start", 18), ("x: first cold task end", 19), ("x: last cold task start", 20), ("x: last cold task end", 21)]
    for name, i in rows:
        x = us(tl[:, i])
        print(f"   {name:16s} min {x.min():8.1f}  mean {x.mean():8.1f}  max {x.max():8.1f} us")

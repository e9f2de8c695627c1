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
em = EventEmulator(device="cuda", engine="fixpoint", noise_mode=noise, lazy_num_events=True,
                   philox_rounds=int(os.environ.get("V2E_TL_ROUNDS", "10")), **cfg)
T = 10
chain = int(os.environ.get("V2E_TL_CHAIN", "0"))   # > 0: packs go out as one warm-up pack + chains of this length
vid = list(gc.video(1, 5 if not chain else 1 + 3 * chain, T, H, W))
steps = [(k, [fr], [tfr]) for k, fr, tfr in vid] if not chain else \
    [(0, [vid[0][1]], [vid[0][2]])] + [(1 + c * chain, [v[1] for v in vid[1 + c * chain:1 + (c + 1) * chain]],
                                        [v[2] for v in vid[1 + c * chain:1 + (c + 1) * chain]]) for c in range(3)]
for k, frs, tfrs in steps:
    if chain:
        em.clip_ramp = (16,)
        em.forward_clip([f.cuda()[0] for f in frs], tfrs)
    else:
        em(frs[0].cuda(), tfrs[0])
    torch.cuda.synchronize()
    ws = em._workspace
    tl = ws[ws.numel() - (160 * 32 * 8 + 2048 + 1024):].view(torch.int64).cpu().numpy()[:160 * 32].reshape(160, 32)
    tl = tl[tl[:, 0] > 0]
    t0 = tl[:, 0].min()
    npass = int(tl[0, 30])
    us = lambda a: (a - t0) / 1e3
    print(f"pack {k}: CTAs {tl.shape[0]} rounds (pre-pass + full passes) {npass}  entries/tile mean {tl[:,28].mean():.0f} max {tl[:,28].max()} "
          f"spill mean {tl[:,29].mean():.1f} max {tl[:,29].max()} hot mean {tl[:,26].mean():.0f} cold mean {tl[:,25].mean():.0f}")
    # NOTE: %globaltimer ticks are ~1.31 ns on this pool (calibrated against CUDA events in
    # scripts/clip_probe.py): multiply the figures below by 1.31 for microseconds
    rows = [("start", 0), ("setup done", 1), ("stream done", 8), ("  marks refined", 23), ("  entries walked", 14), ("  offsets done", 15), ("post pass done", 2)]
    for ps in range(min(npass, 4)):
        rows += [(f"resolve{ps} done", 3 + 3 * ps)] + ([(f"arrived{ps}", 5 + 3 * ps)] if ps == 2 else []) + [(f"barrier{ps} done", 4 + 3 * ps)]
    rows += [("loop exit", 27), ("commit done", 31)]
    if os.environ.get("V2E_STAMPS"):
        c = tl[:, 24]
        print(f"   first full round, cold event lanes per tile: slow(shot) {(c & 0xffff).mean():.1f} slow(other) {((c >> 16) & 0xffff).mean():.1f} classified {((c >> 32) & 0xffff).mean():.1f} simple {((c >> 48) & 0xffff).mean():.1f}")
        rows += [("x: last task end (max warp)", 15)]
    if os.environ.get("V2E_STAMPS") and npass >= 3:
        # which tiles make the first full round late?
        c = tl[:, 24]
        dur = us(tl[:, 9]) - us(tl[:, 7])   # resolve2 done - barrier1 done
        order = np.argsort(-dur)[:8]
        for o in order:
            print(f"   tile {o:3d}: full round {dur[o]:5.1f} us  entries {tl[o,28]} hot {tl[o,26]} cold {tl[o,25]}  slow(shot) {c[o] & 0xffff} "
                  f"slow(other) {(c[o] >> 16) & 0xffff} classified {(c[o] >> 32) & 0xffff} simple {(c[o] >> 48) & 0xffff}")
        w = tl[:, 16]
        lt = np.argsort(-(w >> 32))[:6]
        for o in lt:
            t = int(w[o] & 0xff); th = int(tl[o, 17]); tc = int(tl[o, 18])
            print(f"   tile {o:3d}: longest task {int(w[o] >> 32)/1e3:5.2f} us = queue index {t} ({'hot' if t < th else 'cold'}; {th} hot + {tc} cold tasks), "
                  f"of which in generic calls (max lane) {int((w[o] >> 8) & 0xffffff) * 16 / 1e3:5.2f} us; task 0 ended at {us(tl[o, 19]) - us(tl[o, 7]):5.2f} us after the barrier")
        print(f"   longest task per tile: mean {(w >> 32).mean()/1e3:.2f} us")
        q = tl[:, 20]
        print(f"   generic calls per tile (hot/cold lanes): lean refused early {(q & 0xff).mean():.2f}/{((q >> 8) & 0xff).mean():.2f}, crossing in the frame "
              f"{((q >> 16) & 0xff).mean():.2f}/{((q >> 24) & 0xff).mean():.2f}, second firing in sight {((q >> 32) & 0xff).mean():.2f}/{((q >> 40) & 0xff).mean():.2f}; "
              f"hot slow frames: marked without shot data {((q >> 48) & 0xff).mean():.2f}, unmarked {((q >> 56) & 0xff).mean():.2f}")
        print(f"   shot-event calls (first full round): longest {tl[:,22].max()/1e3:.2f} us, total per tile mean {tl[:,21].mean()/1e3:.2f} us max {tl[:,21].max()/1e3:.2f} us")
        sl = ((c & 0xffff) + ((c >> 16) & 0xffff)) > 0
        if sl.any() and (~sl).any():
            print(f"   full round: tiles with a slow cold lane {dur[sl].mean():.1f} us ({sl.sum()}), others {dur[~sl].mean():.1f} us")
    for name, i in rows:
        x = us(tl[:, i])
        print(f"   {name:16s} min {x.min():8.1f}  mean {x.mean():8.1f}  max {x.max():8.1f} us")

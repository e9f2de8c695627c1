"""Per-CTA timeline of the run-ahead kernel (V2E_TIMELINE=1): where a pack's time goes."""
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
em = EventEmulator(device="cuda", engine="runahead", noise_mode=noise, lazy_num_events=True, integer_frames=bool(int(__import__("os").environ.get("V2E_INTF","1"))), **cfg)
T = 10
for k, fr, tfr in gc.video(1, 4, T, H, W):
    em(fr.cuda(), tfr)
torch.cuda.synchronize()
ws = em._workspace
swork = 8 * 36 + 8 + 8 + 4 + 4 + 8 * 3 * 2048 + 4 * 32 * 160 + 8 * 4 * 160
vox_off = (swork + 255) & ~255
tl = ws[vox_off:vox_off + 148 * 32 * 8 * 8].view(torch.int64).cpu().numpy().reshape(148, 32, 8)
ok = tl[:, 0, 0] > 0
tl = tl[ok].astype(np.float64)
print("CTAs with data:", tl.shape[0])
t0 = tl[:, 0, 0][:, None]
def stat(name, x):
    print(f"{name:34s} mean {x.mean():9.0f} med {np.median(x):9.0f} p95 {np.percentile(x,95):9.0f} max {x.max():9.0f}")
stat("prologue (start->sync)", tl[:, 0, 1] - tl[:, 0, 0])
fr = tl[:, 1:T, :]
stat("stream: frame1 done (from start)", fr[:, 0, 0] - t0[:, 0])
stat("stream: per-frame period", np.diff(fr[:, :, 0], axis=1))
stat("stream: last frame done", fr[:, -1, 0] - t0[:, 0])
stat("resolve: wait grid barrier", fr[:, :, 4] - fr[:, :, 1])
stat("resolve: tail+events+step (thread 0)", fr[:, :, 5] - fr[:, :, 4])
stat("resolve: wait slowest + arrive", fr[:, :, 3] - fr[:, :, 5])
stat("resolve: list+load", fr[:, :, 6] - fr[:, :, 3])
stat("resolve: round period", np.diff(fr[:, :, 1], axis=1))
stat("resolve: list length", fr[:, :, 7])
stat("roles done (from start)", tl[:, 0, 2] - t0[:, 0])
stat("sync after roles", tl[:, 0, 3] - tl[:, 0, 2])
stat("  parked state write-back", tl[:, 31, 0] - tl[:, 0, 3])
stat("  statistics pass over the grid", tl[:, 31, 1] - tl[:, 31, 0])
stat("  CTA reduce + publish", tl[:, 31, 2] - tl[:, 31, 1])
stat("  wait other tiles + combine", tl[:, 0, 4] - tl[:, 31, 2])
stat("stats barrier", tl[:, 0, 4] - tl[:, 0, 3])
stat("normalise+store", tl[:, 0, 5] - tl[:, 0, 4])
stat("total", tl[:, 0, 5] - t0[:, 0])
for n in range(1, T):
    f = lambda k: np.median(tl[:, n, k] - t0[:, 0])
    print(f"  round {n}: start {f(1):8.0f} released {f(4):8.0f} own work {f(5):8.0f} arrived {f(3):8.0f} listed {f(6):8.0f} | stream done {f(0):8.0f}  list {np.median(tl[:, n, 7]):5.0f}")

for r, n in ((16, 5), (20, 6)):
    x = tl[:, r:r + 4, :].reshape(tl.shape[0], 32)
    print(f"round {n}: per-warp critical-section cycles, median over CTAs per warp id:")
    print("   " + " ".join(f"{np.median(x[:, w]):5.0f}" for w in range(32)))
    print(f"   per-CTA max: median {np.median(x.max(1)):6.0f}  p95 {np.percentile(x.max(1),95):6.0f}  max {x.max():6.0f}; all warps: median {np.median(x):6.0f}")

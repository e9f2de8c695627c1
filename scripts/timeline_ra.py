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
em = EventEmulator(device="cuda", engine="runahead", noise_mode=noise, lazy_num_events=True, **cfg)
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
stat("resolve: release -> arrive", fr[:, :, 3] - fr[:, :, 4])
stat("   slowest tail (per CTA)", fr[:, :, 2])
stat("   slowest events+step (per CTA)", fr[:, :, 5])
stat("resolve: list+load", fr[:, :, 6] - fr[:, :, 3])
stat("resolve: round period", np.diff(fr[:, :, 1], axis=1))
stat("resolve: list length", fr[:, :, 7])
stat("roles done (from start)", tl[:, 0, 2] - t0[:, 0])
stat("sync after roles", tl[:, 0, 3] - tl[:, 0, 2])
stat("stats barrier", tl[:, 0, 4] - tl[:, 0, 3])
stat("normalise+store", tl[:, 0, 5] - tl[:, 0, 4])
stat("total", tl[:, 0, 5] - t0[:, 0])
for n in range(1, T):
    f = lambda k: np.median(tl[:, n, k] - t0[:, 0])
    print(f"  round {n}: start {f(1):8.0f} released {f(4):8.0f} arrived {f(3):8.0f} listed {f(6):8.0f} | max tail {np.median(tl[:, n, 2]):6.0f} max ev {np.median(tl[:, n, 5]):6.0f} (worst CTA {tl[:, n, 2].max():6.0f} {tl[:, n, 5].max():6.0f}) | stream done {f(0):8.0f}  list {np.median(tl[:, n, 7]):5.0f}")

for n in range(1, T):
    x = tl[:, n + 16, :3]
    print(f"  round {n}: slowest thread per CTA (median over CTAs / worst): entry-load wait {np.median(x[:,0]):6.0f}/{x[:,0].max():6.0f}  events {np.median(x[:,1]):6.0f}/{x[:,1].max():6.0f}  step {np.median(x[:,2]):6.0f}/{x[:,2].max():6.0f}")

for n in range(1, T):
    x = tl[:, n + 16, :]
    k = int(np.argmax(x[:, 3]))
    print(f"  round {n}: slowest do_events over the chip: total {x[k,3]:6.0f} = fetch+load-wait {x[k,4]:6.0f} + diff_count {x[k,5]:6.0f} + iterate/stores {x[k,6]:6.0f}; cnt*1000+fin {x[k,7]:6.0f} | median CTA: {np.median(x[:,3]):6.0f} = {np.median(x[:,4]):6.0f} + {np.median(x[:,5]):6.0f} + {np.median(x[:,6]):6.0f}")

"""Dump the per-CTA, per-frame thread-0 timeline of the persistent kernel (V2E_TIMELINE=1)."""
import os, sys
os.environ["V2E_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import gpu_common as gc
from v2e2v_b200 import EventEmulator, _cabi

H, W = int(sys.argv[1]), int(sys.argv[2])
noise = sys.argv[3] if len(sys.argv) > 3 else "philox"
cfg = dict(gc.CFG2)
if noise == "off":
    cfg.update(leak_rate_hz=0, shot_noise_rate_hz=0)
em = EventEmulator(device="cuda", engine="persistent", noise_mode=noise, lazy_num_events=True, **cfg)
T = 10
for k, fr, tfr in gc.video(1, 4, T, H, W):
    em(fr.cuda(), tfr)
torch.cuda.synchronize()
# timeline lives in the stepwise accumulator area of the workspace
P = H * W
off = (( _cabi.lib().v2e_workspace_bytes.__self__ is None) and 0) if False else None
import ctypes
ws = em._workspace
# SeqWork size rounded to 256
swork = 8 * 36 + 8 + 8 + 4 + 4 + 8 * 3 * 2048 + 4 * 32 * 160 + 8 * 4 * 160
vox_off = (swork + 255) & ~255
tl = ws[vox_off:vox_off + 148*32*8*8].view(torch.int64).cpu().numpy().reshape(148, 32, 8)
tl = tl[:, 1:T, :]
ok = tl[:, :, 0] > 0
tl = tl[ok[:, 0]]
print("CTAs with data:", tl.shape[0])
names = ["start->tma", "tma->A done", "A->sync1", "sync1->arrive", "arrive->released", "released->sync2", "sync2->B done", "B->next start"]
d = np.diff(tl, axis=2).astype(np.float64)          # [cta, frame, 7]
nxt = np.concatenate([tl[:, 1:, 0] - tl[:, :-1, 7], np.zeros((tl.shape[0], 1))], axis=1)
for i in range(7):
    x = d[:, :, i]
    print(f"{names[i]:20s} mean {x.mean():9.0f}  med {np.median(x):9.0f}  p95 {np.percentile(x,95):9.0f}  max {x.max():9.0f} cycles")
x = nxt[:, :-1]
print(f"{names[7]:20s} mean {x.mean():9.0f}  med {np.median(x):9.0f}  p95 {np.percentile(x,95):9.0f}  max {x.max():9.0f} cycles")
per_frame = (tl[:, 1:, 0] - tl[:, :-1, 0]).astype(np.float64)
print("frame period: mean %.0f cycles" % per_frame.mean())

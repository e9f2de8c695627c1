"""Per-pack view of the bench workload (development aid): rounds, entries, spills, kernel time."""
import os, sys
os.environ["V2E_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from v2e2v_b200 import EventEmulator
noise = sys.argv[1] if len(sys.argv) > 1 else "philox"
H, W = 720, 1280
dev = torch.device("cuda", 0)
EMU = dict(bench.EMU)
if noise == "off":
    EMU.update(leak_rate_hz=0, shot_noise_rate_hz=0)
gen = torch.Generator(device=dev).manual_seed(1234)
torch.manual_seed(0)
em = EventEmulator(device=dev, noise_mode=noise, engine="auto", lazy_num_events=True, **EMU)
base = torch.rand(1, 1, H, W, device=dev, generator=gen) * 255
packs = []
for _ in range(5):
    fr, base = bench.synth_pack_device(gen, base, H, W)
    packs.append(fr)
for k in range(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    v, n = em(packs[k % 5], bench.stamps(k))
    e1.record()
    torch.cuda.synchronize()
    ws = em._workspace
    tl = ws[ws.numel() - (160 * 32 * 8 + 2048):].view(torch.int64).cpu().numpy()[:160 * 32].reshape(160, 32)
    tl = tl[tl[:, 0] > 0]
    print(f"pack {k:3d} ring {k%5}: kernel {1e3*e0.elapsed_time(e1):6.1f} us  rounds {int(tl[0,30])}  entries/tile mean {tl[:,28].mean():6.0f} max {tl[:,28].max():5d}  "
          f"spill mean {tl[:,29].mean():6.1f} max {tl[:,29].max():5d}  hot mean {tl[:,26].mean():5.0f}  events {int(n.item())}")

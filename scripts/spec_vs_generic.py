"""Specialised fixed-point kernel (no debug maps) vs the generic one (debug maps on): same seed,
same video => bit-identical states, event counts and voxel grids.  Development aid; the same
check lives in tests/test_emulator_gpu.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import gpu_common as gc

def run(H, W, noise, fdt=torch.float32, packs=4, T=10, cfgname="CFG2"):
    cfg = dict(getattr(gc, cfgname))
    if noise == "off":
        cfg.update(leak_rate_hz=0, shot_noise_rate_hz=0)
    pos, neg, nra = gc.make_planes(5, H, W, cfg)
    res = []
    for dbg in (True, False):
        em = gc.cuda_emulator(cfg, "fixpoint", noise_mode=noise, debug_maps=dbg)
        em.philox_seed = 4242
        em.set_planes(pos, neg, nra)
        out = []
        for k, fr, tfr in gc.video(3, packs, T, H, W):
            ev, nev = em(fr.to("cuda").to(fdt), tfr)
            out.append((ev.cpu().numpy().copy(), nev, em.base_log_frame.cpu().numpy().copy(),
                        em.lp_log_frame0.cpu().numpy().copy() if cfg["cutoff_hz"] > 0 else None,
                        em.timestamp_mem.cpu().numpy().copy()))
        res.append(out)
    ok = True
    for k, (a, b) in enumerate(zip(*res)):
        same = [np.array_equal(a[0], b[0]), a[1] == b[1], np.array_equal(a[2], b[2]),
                a[3] is None or np.array_equal(a[3], b[3]), np.array_equal(a[4], b[4])]
        if not all(same):
            ok = False
            nd = lambda x, y: int((x != y).sum())
            print(f"  pack {k}: vox/nev/base/lp/tmem same = {same}; nev {a[1]} vs {b[1]}; "
                  f"diffs vox {nd(a[0], b[0])} base {nd(a[2], b[2])} tmem {nd(a[4], b[4])}")
    print(f"{H}x{W} {cfgname} noise={noise} fdt={str(fdt)[6:]}: {'IDENTICAL' if ok else 'MISMATCH'} (events/pack {res[0][-1][1]})")
    return ok

if __name__ == "__main__":
    ok = True
    fdts = [torch.float32] if os.environ.get("V2E_DEV") else [torch.float32, torch.uint8, torch.bfloat16]
    for (H, W) in [(96, 128), (260, 346), (720, 1280)]:
        for noise in ("philox", "off"):
            for fdt in fdts:
                ok &= run(H, W, noise, fdt)
    ok &= run(180, 240, "philox", cfgname="CFG1")
    ok &= run(180, 240, "off", cfgname="CFG1")
    sys.exit(0 if ok else 1)

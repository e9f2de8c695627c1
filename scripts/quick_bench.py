"""Quick device-side timing of the emulator engines (development aid, not the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_common as gc
from v2e2v_b200 import EventEmulator

def run(H, W, engine, noise, fdt=torch.float32, packs=30, T=10, rounds=10):
    cfg = dict(gc.CFG2)
    if noise == "off":
        cfg.update(leak_rate_hz=0, shot_noise_rate_hz=0)
    torch.manual_seed(0)
    em = EventEmulator(device="cuda", engine=engine, noise_mode=noise, philox_rounds=rounds, integer_frames=bool(int(__import__("os").environ.get("V2E_INTF","1"))), **cfg)
    g = torch.Generator(device="cuda").manual_seed(1)
    # device-generated random-walk video, several distinct packs (inputs > L2 in total)
    nvid = 8
    base = torch.rand(1, 1, H, W, device="cuda", generator=g) * 255
    vids = []
    for k in range(nvid):
        fr = (base + torch.cumsum(torch.randn(1, T, H, W, device="cuda", generator=g) * 6, 1)).clamp(0, 255).round()
        fr[:, 0] = base[:, 0]
        base = fr[:, -1:].clone()
        vids.append(fr.to(fdt).contiguous())
    ts = [torch.linspace(k * (T - 1) / 1000, (k + 1) * (T - 1) / 1000, T, dtype=torch.float64)[None] for k in range(packs + 5)]
    for k in range(5):
        em(vids[k % nvid], ts[k])
    torch.cuda.synchronize()
    # (a) API loop incl. python + .item() sync
    t0 = time.perf_counter()
    nev = 0
    for k in range(packs):
        v, n = em(vids[(k + 5) % nvid], ts[k + 5]); nev += n
    torch.cuda.synchronize()
    t_api = (time.perf_counter() - t0) / packs
    # (b) device time: events around a run of back-to-back asynchronous packs
    em.reset()
    em.lazy_num_events = True
    for k in range(5):
        em(vids[k % nvid], ts[k])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    th0 = time.perf_counter()
    e0.record()
    for k in range(packs):
        v, n = em(vids[(k + 5) % nvid], ts[k + 5])
    e1.record()
    th1 = time.perf_counter()
    torch.cuda.synchronize()
    med = e0.elapsed_time(e1) / packs
    dts = [med]
    t_host = (th1 - th0) / packs
    print(f"   host enqueue {t_host*1e6:.1f} us/pack")
    pxf = H * W * T
    print(f"{H}x{W} {engine:10s} noise={noise:8s} fdt={str(fdt)[6:]:8s} R={rounds} engine={em.resolved_engine}: "
          f"api {t_api*1e6:8.1f} us/pack | dev median {med*1e3:8.1f} us min {dts[0]*1e3:8.1f} us "
          f"-> {pxf/med/1e6:8.2f} Gpx-frames/s, {92*H*W/med/1e6:7.1f} GB/s algorithmic; events/pack {nev/packs:.0f}")

if __name__ == "__main__":
    if len(sys.argv) > 1:
        H, W, engine, noise = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
        fdt = {"f32": torch.float32, "u8": torch.uint8, "bf16": torch.bfloat16}[sys.argv[5] if len(sys.argv) > 5 else "f32"]
        run(H, W, engine, noise, fdt=fdt, packs=int(sys.argv[6]) if len(sys.argv) > 6 else 30)
        sys.exit(0)
    for (H, W) in [(720, 1280), (180, 240), (260, 346)]:
        for engine in ("stepwise", "persistent"):
            for noise in ("off", "philox"):
                run(H, W, engine, noise)
    run(720, 1280, "persistent", "philox", rounds=7)
    run(720, 1280, "persistent", "philox", fdt=torch.uint8)
    run(720, 1280, "persistent", "off", fdt=torch.uint8)

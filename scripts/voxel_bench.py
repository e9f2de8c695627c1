"""Standalone voxeliser + normaliser (SURVEY §8 rows a18/a19, v2e_events_to_voxel /
v2e_voxel_normalize): events/s, algorithmic GB/s against the measured HBM peak, and the CPU oracle
port beside it.  One JSON line per case.  python scripts/voxel_bench.py > profiles/..."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as orc
from v2e2v_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
orc.build()
nb = 5
for W, H, n, dt in [(346, 260, 4_000_000, torch.float64), (1280, 720, 16_000_000, torch.float64),
                    (1280, 720, 16_000_000, torch.float32)]:
    g = torch.Generator(device="cuda").manual_seed(5)
    t = torch.sort(torch.rand(n, device="cuda", generator=g, dtype=torch.float64)).values * 0.05 + 1.0
    ev = torch.stack([t, torch.randint(0, W, (n,), device="cuda", generator=g).double(),
                      torch.randint(0, H, (n,), device="cuda", generator=g).double(),
                      torch.randint(0, 2, (n,), device="cuda", generator=g).double()], 1).to(dt).contiguous()
    P = W * H
    es = ev.element_size()
    for _ in range(3):
        vox = ops.events_to_voxel_op(ev, nb, W, H, False)
        ops.voxel_normalize_op(vox, "std", 25. / nb)
    reps = 20
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tv = tn = 0.0
    for _ in range(reps):
        e[0].record()
        vox = ops.events_to_voxel_op(ev, nb, W, H, False)
        e[1].record()
        ops.voxel_normalize_op(vox, "std", 25. / nb)
        e[2].record()
        torch.cuda.synchronize()
        tv += e[0].elapsed_time(e[1]) / reps
        tn += e[1].elapsed_time(e[2]) / reps
    bytes_v = 4 * es * n + 4 * nb * P          # rows read + grid written (the zero fill is part of the op)
    bytes_n = 2 * 4 * nb * P + 4 * nb * P      # read for the statistics, read + written to normalise
    # CPU: the oracle port (sequential C, event order) on a bounded sample of the same list
    ns = min(n, 2_000_000)
    evc = ev[:ns].cpu().numpy()
    t0 = time.perf_counter()
    vc = orc.events_to_voxel(evc, nb, W, H)
    t1 = time.perf_counter()
    orc.voxel_normalize(vc, 'std', 25. / nb)
    t2 = time.perf_counter()
    print(json.dumps({
        "op": "events_to_voxel_grid + event_preprocess('std', hot pixel filter)", "width": W, "height": H,
        "num_bins": nb, "events": n, "event_dtype": str(dt).split(".")[-1],
        "voxelise_ms": round(tv, 4), "Mevents_per_s": round(n / tv / 1e3, 1),
        "voxelise_roofline": {"bound": "hbm", "achieved": round(bytes_v / tv / 1e6, 1), "peak": PEAK, "unit": "GB/s",
                              "frac": round(bytes_v / tv / 1e6 / PEAK, 4), "algorithmic_bytes": bytes_v},
        "normalise_ms": round(tn, 4),
        "normalise_roofline": {"bound": "hbm", "achieved": round(bytes_n / tn / 1e6, 1), "peak": PEAK, "unit": "GB/s",
                               "frac": round(bytes_n / tn / 1e6 / PEAK, 4), "algorithmic_bytes": bytes_n},
        "cpu_baseline": {"kind": "port", "cores": 1, "sample": f"{ns} events of the same list through oracle/v2e_oracle.c",
                         "Mevents_per_s": round(ns / (t1 - t0) / 1e6, 2), "normalise_ms": round(1e3 * (t2 - t1), 3)},
    }), flush=True)

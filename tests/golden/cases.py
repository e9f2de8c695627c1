"""Golden-fixture case table shared by make_golden.py (generator) and the parity tests."""
from __future__ import annotations

SD = dict(pl=1.5, ps=0.5, ql=1, qs=0, cutoff_hz=200)          # sensing-diversity (BASELINE cfg2)
NOISE = dict(leak_rate_hz=0.1, shot_noise_rate_hz=1)          # model_v2e2v.py:56-57
BASE = dict(output_mode='voxel_grid', num_bins=5, pos_thres=0.6, neg_thres=0.6, sigma_thres=0.03,
            refractory_period_s=0.001, pl=1, ps=1, ql=1, qs=1, cutoff_hz=0,
            leak_rate_hz=0, shot_noise_rate_hz=0)

# name -> dict(emu=EventEmulator kwargs, H, W, T, packs, fps, stamps=('full'|'two', 'f64'|'f32'),
#              sigma=video step, drop_first=bool (T-1 frames + T stamps after pack 0, trap T3))
CASES = {
    "cfg1_default": dict(emu={**BASE, **NOISE}, H=20, W=28, T=10, packs=4, fps=1000,
                         stamps=("full", "f64")),
    "cfg2_sd_noise": dict(emu={**BASE, **SD, **NOISE}, H=20, W=28, T=10, packs=4, fps=1000,
                          stamps=("full", "f64")),
    "cfg2_sd_noiseless": dict(emu={**BASE, **SD}, H=20, W=28, T=10, packs=4, fps=1000,
                              stamps=("full", "f64")),
    "twocol_f32": dict(emu={**BASE, **SD, **NOISE}, H=16, W=24, T=10, packs=3, fps=1000,
                       stamps=("two", "f32")),
    "twocol_f64": dict(emu={**BASE, **SD, **NOISE}, H=16, W=24, T=10, packs=3, fps=1000,
                       stamps=("two", "f64")),
    "full_f32": dict(emu={**BASE, **SD, **NOISE}, H=16, W=24, T=10, packs=3, fps=1000,
                     stamps=("full", "f32")),
    "qlqs_120fps": dict(emu={**BASE, **NOISE, **dict(pl=1.2, ps=0.8, ql=1, qs=0.5, cutoff_hz=30,
                                                     refractory_period_s=0.0005)},
                        H=16, W=24, T=8, packs=3, fps=120, stamps=("full", "f64")),
    "lowfps_gate_off": dict(emu={**BASE, **NOISE, **dict(pos_thres=0.3, neg_thres=0.3)},
                            H=16, W=24, T=10, packs=3, fps=30, stamps=("full", "f64"), sigma=12.0),
    "big_refractory": dict(emu={**BASE, **SD, **dict(refractory_period_s=0.005)},
                           H=16, W=24, T=10, packs=3, fps=1000, stamps=("full", "f64"),
                           sigma=10.0),
    "raw_mode": dict(emu={**BASE, **SD, **NOISE, **dict(output_mode='raw')}, H=16, W=24, T=10,
                     packs=3, fps=1000, stamps=("full", "f64")),
    "odd_width_nb3": dict(emu={**BASE, **SD, **NOISE, **dict(num_bins=3)}, H=18, W=26, T=6,
                          packs=3, fps=500, stamps=("full", "f64")),
    "drop_first_frame": dict(emu={**BASE, **SD, **NOISE}, H=16, W=24, T=10, packs=3, fps=1000,
                             stamps=("full", "f64"), drop_first=True),
    "ql0_pos_ne_neg": dict(emu={**BASE, **NOISE, **dict(ql=0, qs=1, cutoff_hz=100, pos_thres=0.4,
                                                        neg_thres=0.7, pl=1.0, ps=1.3)},
                           H=16, W=24, T=10, packs=3, fps=1000, stamps=("full", "f64")),
    "t2_minimal": dict(emu={**BASE, **SD, **NOISE}, H=12, W=16, T=2, packs=4, fps=1000,
                       stamps=("full", "f64")),
}

SHOT_IMAX = 40   # iterations of injected shot-noise draws available per frame

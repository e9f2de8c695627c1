#!/usr/bin/env python
"""bench.py — V2E emulate+voxelize throughput on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

A *step* is one pack (num_pack_frames = 10 HFR frames) of every sequence a rank owns, through the
fused emulate+voxelize+normalise path.  Metric: Gpixel-frames/s = sequences*H*W*T*steps / time,
summed over ranks (weak scaling: each rank owns its own sequences; no collective on the data
path).  One JSON line is printed by rank 0.

  value     device-resident inputs (a ring of distinct packs larger than L2), enqueued in
            clips of 50 packs (EventEmulator.forward_clip -> v2e_clip_forward: chained launches
            of up to 16 packs), timed with CUDA events, barrier + synchronize on both sides, max
            over ranks.
  e2e       same metric through the public EventEmulator.forward with frames in pinned HOST
            memory: per step one H2D copy of the pack and a D2H read of num_events.
  roofline  algorithmic bytes (92 B/pixel/pack; x packs per launch for a chained
            launch) / kernel duration (CUDA events around each launch, over a region of up to
            200 packs that follows the timed one: the events themselves cost ~20 us per
            launch) vs MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle port (oracle/v2e_oracle.c, OpenMP) on a bounded sample.

--impl reference times the CPU port on the host cores (the reference itself is pure Python and
cannot travel to the GPU box; BASELINE.md holds its measured numbers from the build container).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

T_FRAMES = 10
FPS = 1000.0
WORKLOADS = {
    # name: (H, W, sequences per GPU, description)
    "cfg3": (720, 1280, 1, "BASELINE configs[2]: synthetic 1280x720 HFR video @1000 fps, 10-frame "
                           "packs, fused emulate+voxelize, sensing-diversity parameters"),
    "cfg2": (180, 240, 1, "BASELINE configs[1]: sensing-diversity V2E 180x240 + 5-bin voxel grid"),
    "cfg4": (260, 346, 8, "BASELINE configs[3]: 346x260 sequences sharded by sequence, 8 per GPU"),
    # (negative: the TOTAL number of sequences, sharded contiguously over the ranks - strong scaling)
    "cfg4x64": (260, 346, -64, "BASELINE configs[3]: batch of 64 synthetic 346x260 sequences sharded "
                               "by sequence across the GPUs (64/N per GPU)"),
}
EMU = dict(output_mode="voxel_grid", num_bins=5, pos_thres=0.6, neg_thres=0.6, sigma_thres=0.03,
           refractory_period_s=0.001, leak_rate_hz=0.1, shot_noise_rate_hz=1,
           pl=1.5, ps=0.5, ql=1, qs=0, cutoff_hz=200)     # cfg2 params + model_v2e2v.py:56-57
ALGO_BYTES_PER_PIXEL_PACK = 92      # 9 f32 frames + 5 f32 bins + 6 planes in + 3 planes out
# measured DRAM bytes per PACK (ncu --set full: profiles/r02_k_fixpoint_spec_chain16_* for the
# fixed-point engine - a chained launch of 16 packs, 615.1 MB read + 308.2 MB written; profiles/r01_*
# for the others): (workload, noise, engine)
TRAFFIC_BYTES = {("cfg3", "philox", "persistent"): 55447040 + 3553024,
                 ("cfg3", "philox", "fixpoint"): (615143936 + 308164352) // 16,
                 ("cfg3", "off", "fixpoint"): 51731968 + 3537152}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def nvml_handle(index: int):
    """NVML handle of the CUDA device `index` of this process (LOCAL_RANK indexes
    CUDA_VISIBLE_DEVICES; NVML wants the physical device)."""
    import pynvml as nv
    nv.nvmlInit()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        tok = vis.split(",")[index].strip()
        return nv.nvmlDeviceGetHandleByUUID(tok) if tok.startswith("GPU-") else \
            nv.nvmlDeviceGetHandleByIndex(int(tok))
    return nv.nvmlDeviceGetHandleByIndex(index)


def bind_near_gpu(index: int):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity) before any host buffer
    exists: with one rank per GPU the pinned frame buffers of `e2e` then sit in the GPU's own NUMA
    node instead of wherever the launcher started the rank.  Returns what was done (for `config`)."""
    # opt-in (BENCH_BIND=1): on the single-NUMA-node boxes of this pool it changes nothing
    # (e2e at 8 GPUs 51.6 vs 51.1 Gpx-frames/s: the host side tops out at ~185 GB/s of aggregate
    # H2D traffic either way), and a bound rank would also confine the cpu_baseline's threads
    if not os.environ.get("BENCH_BIND"):
        return "off"
    try:
        import pynvml as nv
        nv.nvmlDeviceSetCpuAffinity(nvml_handle(index))
        return f"{len(os.sched_getaffinity(0))} cpus"
    except Exception as e:      # (restricted cpusets, no NVML: run unbound)
        return f"unavailable ({type(e).__name__})"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed regions run: NVML in-process (the
    rank's own GPU, every 10 ms - spawning nvidia-smi from every rank every 50 ms stalled other
    ranks' launches in the multi-GPU runs), nvidia-smi as the fall-back."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.source = None
        self.period = float(os.environ.get("BENCH_SAMPLER_MS", "10")) * 1e-3
        self._paused = threading.Event()
        self.ready = threading.Event()     # NVML is up (or the fall-back was chosen)
        self._nv = None          # (module, handle, max clock, reason bits) once NVML is up
        self.in_region = 0       # samples taken by sample_now() while the timed region ran

    def _nvml_loop(self) -> bool:
        try:
            import pynvml as nv
            h = nvml_handle(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = [(nv.nvmlClocksEventReasonHwSlowdown, 0), (nv.nvmlClocksEventReasonHwThermalSlowdown, 1),
                    (nv.nvmlClocksEventReasonSwThermalSlowdown, 2), (nv.nvmlClocksEventReasonSwPowerCap, 3)]
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            return False
        self.source = "nvml"
        self._nv = (nv, h, mx, bits)
        self.ready.set()
        while not self._halt.is_set():
            if not self._paused.is_set():
                self._sample()
            self._halt.wait(self.period)
        return True

    def _sample(self) -> bool:
        if self._nv is None:
            return False
        nv, h, mx, bits = self._nv
        try:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.rows.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b, _ in bits])
            return True
        except Exception:
            return False

    def pause(self):
        """No background sample (a second thread taking the GIL, an NVML call inside the driver)
        while the host enqueues the short timed region; `sample_now` covers it instead."""
        self._paused.set()

    def resume(self):
        self._paused.clear()

    def sample_now(self, still_running) -> None:
        """Samples from the calling thread while `still_running()` (the timed region's work is
        all enqueued and the device is executing it: nothing is delayed)."""
        for _ in range(8):
            if not still_running():
                break
            if self._sample():
                self.in_region += 1

    def run(self):
        if os.environ.get("BENCH_NO_SAMPLER"):     # development aid
            self.ready.set()
            return
        if not os.environ.get("BENCH_SAMPLER_SMI") and self._nvml_loop():
            return
        self.source = "nvidia-smi"
        self.ready.set()
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if r[1].isdigit()] or [0])
        reasons = sorted({n for r in self.rows for n, v in zip(self.NAMES, r[2:]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": reasons, "samples": len(self.rows),
                "samples_inside_value_region": self.in_region, "source": self.source}


def synth_pack_device(gen, base, H, W, dtype=torch.float32):
    """synth-walk-v1 on the device: integer-valued random-walk frames."""
    fr = (base + torch.cumsum(torch.randn(1, T_FRAMES, H, W, device=base.device, generator=gen) * 6,
                              dim=1)).clamp(0, 255).round()
    fr[:, 0] = base[:, 0]
    return fr.to(dtype).contiguous(), fr[:, -1:].clone()


_STAMPS = {}


def stamps(k):
    # the time stamps of pack k are an input of the run like its frames: built once, ahead of the
    # timed regions (torch.linspace costs ~10 us a call, a seventh of a pack)
    t = _STAMPS.get(k)
    if t is None:
        t = _STAMPS[k] = torch.linspace(k * (T_FRAMES - 1) / FPS, (k + 1) * (T_FRAMES - 1) / FPS,
                                        T_FRAMES, dtype=torch.float64)[None]
    return t


# -----------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    from v2e2v_b200 import EventEmulator, EventEmulatorBatch, _cabi
    binding = bind_near_gpu(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    H, W, S, desc = WORKLOADS[args.workload]
    strong = S < 0
    if strong:
        from v2e2v_b200.sharding import shard_sequences
        S = len(shard_sequences(-S, rank, world))      # this rank's contiguous block
    K, Wm = args.steps, max(args.warmup, 3)
    fbytes = 1 if args.frames == "u8" else 4
    ring = max(2, -(-160_000_000 // (H * W * T_FRAMES * fbytes * S)))   # > L2 (126 MB) of inputs
    ring = min(ring, 16 if fbytes == 4 else 20)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    torch.manual_seed(rank)
    ems, packs = [], []
    # S > 1: the sequences of a rank go out S per launch (EventEmulatorBatch: the same S
    # reference-style emulators, CTAs grouped per sequence)
    batch = None
    if S > 1:
        batch = EventEmulatorBatch(S, device=dev, noise_mode=args.noise, engine=args.engine,
                                   philox_rounds=args.philox_rounds, lazy_num_events=True, **EMU)
    for s in range(S):
        em = batch[s] if batch is not None else EventEmulator(
            device=dev, noise_mode=args.noise, engine=args.engine,
            philox_rounds=args.philox_rounds, lazy_num_events=True, **EMU)
        ems.append(em)
        base = torch.rand(1, 1, H, W, device=dev, generator=gen) * 255
        seqp = []
        for _ in range(ring):
            fr, base = synth_pack_device(gen, base, H, W)
            seqp.append(fr)
        packs.append(seqp)
    pinned = [[p.cpu().pin_memory() for p in seqp] for seqp in packs]
    if args.frames == "u8":
        # the device-resident frames as the reference's readers hold them before .float() (lossless:
        # integer-valued 0..255; bit-identical results); `e2e` keeps its f32 host frames
        packs = [[p.to(torch.uint8) for p in seqp] for seqp in packs]
    # [S, T, H, W] per ring slot for the batched launches
    bpacks = ([torch.cat([packs[s][r] for s in range(S)]) for r in range(ring)]
              if batch is not None else None)
    lib = _cabi.lib()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value") -------------------------------------------------
    k = 0
    for _ in range(Wm):
        if batch is not None:
            batch(bpacks[k % ring], stamps(k).expand(S, -1))
        else:
            ems[0](packs[0][k % ring], stamps(k))
        k += 1
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # the K packs of a sequence are enqueued in clips (EventEmulator.forward_clip: one C call,
    # scalars prepared ahead), so the device does not wait for Python between launches
    CL = 50
    G = 1 if batch is None else -(-S // EventEmulatorBatch.MAX_PER_LAUNCH)   # launches per step

    clip_ev = []

    def run_clips(k_first, n_steps, events=None):
        kk, i = k_first, 0
        while kk < k_first + n_steps:
            n = min(CL, k_first + n_steps - kk)
            if events is None:      # one event per clip: where the time goes (stderr)
                clip_ev.append((torch.cuda.Event(enable_timing=True), n, time.perf_counter()))
                clip_ev[-1][0].record()
            # (one event pair per step and launch group, group-major as forward_clip takes them)
            evs = None if events is None else [e for g in range(G) for e in events[g][i:i + 2 * n]]
            if batch is not None:
                batch.forward_clip([bpacks[(kk + j) % ring] for j in range(n)],
                                   [stamps(kk + j) for j in range(n)], events=evs)
            else:
                ems[0].forward_clip([packs[0][(kk + j) % ring] for j in range(n)],
                                    [stamps(kk + j) for j in range(n)], events=evs)
            i += 2 * n
            kk += n
        return kk

    for kk in range(0, Wm + CL + K + min(K, 200) + 8):
        stamps(kk)
    # warm the clip path too (its output block comes from cudaMalloc the first time: ~40 ms)
    import gc
    gc.collect()        # (before the warm clip: the device must not sit idle ahead of the timed region)
    k = run_clips(k, CL)
    clip_ev.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc.disable()        # no collector pause between two enqueues of the timed region
    sampler.ready.wait(5.0)
    sampler.pause()     # ... and no second thread: the region is sampled below, once it is enqueued
    barrier()
    launches0 = lib.v2e_launch_count()
    e0.record()
    k = run_clips(k, K)
    e1.record()
    sampler.sample_now(lambda: not e1.query())   # clocks / throttle reasons while the region runs
    barrier()
    sampler.resume()
    gc.enable()
    if batch is not None:
        batch.check_status()
    else:
        ems[0].check_status()
    if rank == 0 and len(clip_ev) > 1:
        per = [clip_ev[j][0].elapsed_time(clip_ev[j + 1][0]) / clip_ev[j][1] * 1e3
               for j in range(len(clip_ev) - 1)]
        host = [(clip_ev[j + 1][2] - clip_ev[j][2]) / clip_ev[j][1] * 1e6
                for j in range(len(clip_ev) - 1)]
        print("per-clip device us/step:", " ".join(f"{x:.0f}" for x in per), file=sys.stderr)
        print("per-clip host   us/step:", " ".join(f"{x:.0f}" for x in host), file=sys.stderr)
    launches = lib.v2e_launch_count() - launches0
    ms_total = reduce_max(e0.elapsed_time(e1))
    # kernel durations for the roofline: the same launches again, each bracketed by CUDA events
    # on the launching stream (the events cost ~20 us per launch, so they stay out of `value`)
    KR = min(K, 200)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2 * KR)] for _ in range(G)]
    k = run_clips(k, KR, events=ev)
    barrier()
    # One interval per pack.  When v2e_clip_forward chains packs into one launch the first pack's
    # interval is the whole launch and the others are empty: a launch is a run of intervals that
    # starts at a non-empty one.
    pack_ms = [sum(ev[g][2 * j].elapsed_time(ev[g][2 * j + 1]) for g in range(G)) for j in range(KR)]
    launch_ms, launch_packs = [], []
    for d in pack_ms:
        if d > 0.01 * G or not launch_ms:     # (an empty interval is two event records per group)
            launch_ms.append(d)
            launch_packs.append(1)
        else:
            launch_ms[-1] += d
            launch_packs[-1] += 1
    kern_pack_ms = sum(pack_ms) / len(pack_ms)              # kernel time per pack
    kern_mean_ms = sum(launch_ms) / len(launch_ms)          # ... per launch
    packs_per_launch = sum(launch_packs) / len(launch_packs)
    kern_min_pack_ms = min(m / n for m, n in zip(launch_ms, launch_packs))
    clocks = sampler.stop()     # sampled over the timed region and the kernel-timing region
    pxf_step = S * H * W * T_FRAMES
    value = world * pxf_step * K / (ms_total * 1e-3) / 1e9

    # ---- end to end through the public API with host frames ("e2e") ------------------------------
    nev_host = torch.zeros(S, dtype=torch.int64).pin_memory()

    def run_e2e(host_packs):
        """K steps of EventEmulator.forward_host on packs in pinned host memory: the H2D copy of
        the pack (double-buffered on a copy stream inside the emulator; frame 0 of a steady-state
        pack is never read and is not copied), the kernel, a D2H read of num_events."""
        for em in ems:
            em.reset()

        def step(kk):
            if batch is not None:       # one staged copy and one batched launch per group and step
                _, n = batch.forward_host(host_packs[kk % ring], stamps(kk).expand(S, -1))
                nev_host.copy_(n, non_blocking=True)                                 # D2H result
                return
            _, n = ems[0].forward_host(host_packs[0][kk % ring], stamps(kk))
            nev_host[0:1].copy_(n, non_blocking=True)                                # D2H result
        for kk in range(Wm):
            step(kk)
        barrier()
        t0 = time.perf_counter()
        for kk in range(Wm, Wm + K):
            step(kk)
        torch.cuda.synchronize(dev)
        secs = reduce_max(time.perf_counter() - t0)
        barrier()
        return world * pxf_step * K / secs / 1e9

    if batch is not None:       # [S, T, H, W] per ring slot, pinned
        pinned = [torch.cat([pinned[s][r] for s in range(S)]).pin_memory() for r in range(ring)]
    e2e_value = run_e2e(pinned)
    h2d = S * H * W * (T_FRAMES - 1) * 4      # frames 1..T-1 of a steady-state pack
    d2h = S * 8
    # same, with the frames as the reference's readers hold them on the host (uint8)
    if batch is not None:
        pinned_u8 = [p.to(torch.uint8).pin_memory() for p in pinned]
    else:
        pinned_u8 = [[p.to(torch.uint8).pin_memory() for p in seqp] for seqp in pinned]
    e2e_u8_value = run_e2e(pinned_u8)

    if rank != 0:
        return
    peak, peak_src = peaks()
    # one timed interval = one step = one pack of each of the rank's S sequences (S = 1: one
    # pack; S > 1: one batched launch)
    # (uint8 frames: 9 B of frames per pixel and pack instead of 36 - SURVEY 8d's 65 B figure)
    algo_bytes = (ALGO_BYTES_PER_PIXEL_PACK - (27 if args.frames == "u8" else 0)) * H * W * S
    achieved = algo_bytes / (kern_pack_ms * 1e-3) / 1e9
    tkey = (args.workload, args.noise, ems[0].resolved_engine) if args.frames == "f32" else None
    out = {
        "metric": "V2E+voxel Gpixel-frames/s", "value": round(value, 3), "unit": "Gpixel-frames/s",
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": round(ms_total / K, 5),
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic (synth-walk-v1 random-walk video, integer-valued f32 frames)",
        "config": {"workload": f"{args.workload}: {desc}", "height": H, "width": W,
                   "num_pack_frames": T_FRAMES, "fps": FPS, "sequences_per_gpu": S,
                   "parallelism": f"{world} x independent sequence shard(s), no collective",
                   "cpu_binding": binding,
                   "emulator": EMU, "noise_mode": args.noise, "philox_rounds": args.philox_rounds,
                   "engine": ems[0].resolved_engine, "frame_dtype": args.frames, "voxel_dtype": "f32",
                   "l2_policy": f"inputs larger than L2: ring of {ring} distinct packs "
                                f"({ring * pxf_step * fbytes / 1e6:.0f} MB) per GPU"},
        "e2e": {"value": round(e2e_value, 3), "unit": "Gpixel-frames/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "EventEmulator[Batch].forward_host on f32 packs in pinned host memory: H2D copy "
                        "of frames 1..T-1 each step (frame 0 of a steady-state pack is never read, "
                        "v2e_model.py:263,350), double-buffered on a copy stream, kernel, "
                        "num_events read back each step; PCIe-bound"},
        "e2e_uint8_frames": {"value": round(e2e_u8_value, 3), "unit": "Gpixel-frames/s",
                             "h2d_bytes_per_step": h2d // 4, "d2h_bytes_per_step": d2h,
                             "note": "as e2e, frames passed as uint8 (what the reference's "
                                     "readers hold before .float()); results are bit-identical"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4),
                     "traffic": (int(round(TRAFFIC_BYTES[tkey] * packs_per_launch))
                                 if tkey in TRAFFIC_BYTES else None),
                     "traffic_per_pack": TRAFFIC_BYTES.get(tkey),
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one "
                                       "ncu --set full capture (profiles/): bytes per pack x packs "
                                       "per launch",
                     "peak_source": peak_src, "kernel": {"fixpoint": "k_fixpoint_spec", "persistent": "k_persistent",
                                "runahead": "k_runahead"}.get(ems[0].resolved_engine,
                                                              "k_phase_a/k_phase_b"),
                     "algorithmic_bytes_per_pack": algo_bytes,
                     "packs_per_launch": round(packs_per_launch, 2),
                     "algorithmic_bytes_per_launch": int(round(algo_bytes * packs_per_launch)),
                     "kernel_ms_mean": round(kern_mean_ms, 5),
                     "kernel_ms_per_pack": round(kern_pack_ms, 5),
                     "kernel_ms_per_pack_min": round(kern_min_pack_ms, 5),
                     "kernel_launches_timed": len(launch_ms), "packs_timed": KR * S,
                     "sequences_per_launch": S},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_port(H, W, budget_s=args.cpu_budget)
    print(json.dumps(out), flush=True)


# -----------------------------------------------------------------------------------------------
def cpu_port(H, W, budget_s=12.0, threads=0, steps=None, warmup=1):
    """Time the CPU oracle port (all host threads) on a bounded sample of the workload."""
    from oracle.oracle import OracleEmulator
    from tests.gpu_common import make_planes
    cores = os.cpu_count() or 1
    nthreads = threads or cores
    em = OracleEmulator(**EMU, threads=nthreads)
    em.set_planes(*make_planes(3, H, W, EMU))
    rng = np.random.Generator(np.random.PCG64(5))
    P = H * W
    base = rng.random((1, H, W), dtype=np.float32) * 255
    imax = 24
    lk = rng.standard_normal((T_FRAMES - 1, P), dtype=np.float32)
    sh = rng.random((T_FRAMES - 1, imax, P), dtype=np.float32)

    def pack():
        nonlocal base
        fr = np.clip(base + np.cumsum(rng.standard_normal((T_FRAMES, H, W), dtype=np.float32) * 6, 0),
                     0, 255).round().astype(np.float32)
        fr[0] = base[0]
        base = fr[-1:].copy()
        return fr
    frames = [pack() for _ in range(3)]
    k = 0
    for _ in range(warmup):
        em.forward(frames[k % 3], stamps(k), leak_randn=lk, shot_rand=sh)
        k += 1
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        em.forward(frames[k % 3], stamps(k), leak_randn=lk, shot_rand=sh)
        times.append(time.perf_counter() - t0)
        k += 1
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() - t_start > budget_s or len(times) >= 200:
            break
    med = sorted(times)[len(times) // 2]
    return {"value": round(P * T_FRAMES / med / 1e9, 5), "unit": "Gpixel-frames/s",
            "cores": nthreads, "kind": "port",
            "sample": f"{len(times)} steady-state packs of {H}x{W}x{T_FRAMES} through "
                      f"oracle/v2e_oracle.c (OpenMP, {nthreads} threads, injected noise), median",
            "ms_per_step": round(med * 1e3, 3)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    H, W, S, desc = WORKLOADS[args.workload]
    K, Wm = args.steps, max(args.warmup, 1)
    K = min(K, 40)
    r = cpu_port(H, W, steps=K, warmup=Wm)
    out = {
        "impl": "reference", "metric": "V2E+voxel Gpixel-frames/s", "value": r["value"],
        "unit": "Gpixel-frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (synth-walk-v1)",
        "config": {"workload": f"{args.workload}: {desc}", "height": H, "width": W,
                   "num_pack_frames": T_FRAMES, "fps": FPS, "emulator": EMU,
                   "note": "CPU port of the reference algorithm (the reference is pure Python/"
                           "torch and is absent on the GPU box; its measured CPU numbers are in "
                           "BASELINE.md)"},
        "cpu_baseline": r,
        "e2e": {"value": r["value"], "unit": "Gpixel-frames/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--noise", default="philox", choices=["philox", "off"])
    ap.add_argument("--frames", default="f32", choices=["f32", "u8"],
                    help="dtype of the device-resident frames of `value` (f32: the reference contract)")
    ap.add_argument("--engine", default="auto", choices=["auto", "fixpoint", "persistent", "runahead", "stepwise"])
    ap.add_argument("--philox-rounds", type=int, default=10)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.noise == "off":
        EMU.update(leak_rate_hz=0, shot_noise_rate_hz=0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

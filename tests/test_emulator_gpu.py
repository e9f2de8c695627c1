"""`-m gpu` parity tests: the CUDA emulator vs the golden fixtures and vs the CPU oracle.

Bar (BASELINE.json north_star): state planes, polarity/count maps and num_events BIT-EXACT;
voxel grids within rtol 1e-5 (+atol 1e-6 for the cancellation in (v-mean)/std).
All calls go through the public EventEmulator -> v2e2v::emulate_pack -> C ABI.
"""
import numpy as np
import pytest
import torch

from oracle.oracle import OracleEmulator
from tests import gpu_common as gc
from tests.golden.cases import CASES
from tests.golden_io import GoldenCase, scalars_equal

pytestmark = pytest.mark.gpu

VOX_RTOL, VOX_ATOL = 1e-5, 1e-6


def _to_cuda_frames(frames_u8, dtype=torch.float32):
    return torch.from_numpy(frames_u8.astype(np.float32)).to("cuda").to(dtype)[None]


@pytest.mark.parametrize("engine", gc.ENGINES)
@pytest.mark.parametrize("name", list(CASES))
def test_golden_reference_outputs(name, engine):
    """CUDA path vs outputs of the unmodified reference (tests/golden)."""
    g = GoldenCase(name)
    if g.raw and engine in ("runahead", "fixpoint"):
        pytest.skip("this engine serves the voxel modes only (raw -> persistent/stepwise)")
    em = gc.cuda_emulator(g.emu, engine)
    pos, neg, nra = g.planes()
    shp = (g.H, g.W)
    em.set_planes(pos.reshape(shp), neg.reshape(shp), None if nra is None else nra.reshape(shp))
    for k in range(g.packs):
        frames, t_frames, lk, sh, sc = g.pack_inputs(k)
        if lk is not None:
            em.inject_leak_randn = torch.from_numpy(lk)
        if sh is not None:
            em.inject_shot_rand = torch.from_numpy(sh)
        ev, nev = em(_to_cuda_frames(frames), t_frames)
        assert scalars_equal(em._scalars, sc), "host scalar tables differ from the golden run"
        want = g.pack_outputs(k)
        assert nev == want["num_events"], f"pack {k}"
        got = dict(base=em.base_log_frame, lp=em.lp_log_frame0, tmem=em.timestamp_mem)
        for key in ("base", "lp", "tmem"):
            if key == "lp" and g.emu["cutoff_hz"] <= 0:
                continue
            np.testing.assert_array_equal(got[key].cpu().numpy().reshape(-1), want[key],
                                          err_msg=f"{key} pack {k}")
        if g.raw:
            np.testing.assert_array_equal(ev.cpu().numpy(), want["events"])
        else:
            np.testing.assert_allclose(ev.cpu().numpy().reshape(g.emu["num_bins"], -1),
                                       want["voxel"], rtol=VOX_RTOL, atol=VOX_ATOL)


@pytest.mark.parametrize("engine", gc.ENGINES)
@pytest.mark.parametrize("cfg,H,W,fdt", [
    (gc.CFG2, 180, 240, torch.float32),      # BASELINE configs[1]
    (gc.CFG1, 180, 240, torch.float32),      # BASELINE configs[0] parameters
    (gc.CFG2, 260, 346, torch.uint8),        # configs[3] geometry (W % 4 != 0), u8 ingest
    (gc.CFG2, 50, 70, torch.bfloat16),
])
def test_vs_oracle_injected_noise(engine, cfg, H, W, fdt):
    """Multi-pack run with injected noise: states, maps and counts bit-exact vs the oracle."""
    T, packs, imax = 10, 3, 24
    pos, neg, nra = gc.make_planes(7, H, W, cfg)
    em = gc.cuda_emulator(cfg, engine, debug_maps=True)
    orc = OracleEmulator(**cfg, threads=0)
    em.set_planes(pos, neg, nra)
    orc.set_planes(pos, neg, nra)
    for k, fr, tfr in gc.video(3, packs, T, H, W):
        lk, sh = gc.noise_arrays(100 + k, T, H * W, imax, True, True)
        em.inject_leak_randn, em.inject_shot_rand = torch.from_numpy(lk), torch.from_numpy(sh)
        ev, nev = em(fr.to("cuda").to(fdt), tfr)
        ov, onev = orc.forward(fr.numpy().astype(np.float32), tfr, leak_randn=lk, shot_rand=sh,
                               debug=True)
        assert nev == onev, f"pack {k}"
        gc.assert_state_equal(em, orc, cfg["cutoff_hz"] > 0, f"pack {k}")
        np.testing.assert_array_equal(em.last_debug["max_iters"].cpu().numpy()[1:],
                                      orc.last["max_iters"][1:])
        for key in ("pol", "cnt", "final"):
            np.testing.assert_array_equal(em.last_debug[key].cpu().numpy(), orc.last[key],
                                          err_msg=f"{key} map pack {k}")
        np.testing.assert_allclose(ev.cpu().numpy(), ov, rtol=VOX_RTOL, atol=VOX_ATOL)


@pytest.mark.parametrize("engine", gc.ENGINES)
@pytest.mark.parametrize("rounds", [7, 10])
def test_philox_mode_is_consistent_with_oracle(engine, rounds):
    """PHILOX mode: export the draws the kernel made, inject them into the oracle, and require
    the same bit-exact states (the in-kernel RNG only decides *which* draws are made)."""
    cfg, H, W, T = gc.CFG2, 96, 128, 10
    pos, neg, nra = gc.make_planes(11, H, W, cfg)
    em = gc.cuda_emulator(cfg, engine, noise_mode="philox", debug_maps=True,
                          philox_rounds=rounds)
    orc = OracleEmulator(**cfg, threads=0)
    em.set_planes(pos, neg, nra)
    orc.set_planes(pos, neg, nra)
    zs = []
    for k, fr, tfr in gc.video(5, 3, T, H, W):
        ev, nev = em(fr.to("cuda"), tfr)
        d = em.last_debug
        lk = d["leak_randn"].cpu().numpy()
        sh = gc.shot_rand_from_decisions(d["shot"].cpu().numpy(), d["pol"].cpu().numpy())
        ov, onev = orc.forward(fr.numpy(), tfr, leak_randn=lk, shot_rand=sh)
        assert nev == onev
        gc.assert_state_equal(em, orc, True, f"pack {k}")
        np.testing.assert_allclose(ev.cpu().numpy(), ov, rtol=VOX_RTOL, atol=VOX_ATOL)
        zs.append(lk)
    z = np.concatenate(zs).ravel()
    # ~330k normals: mean ~ N(0, 1/sqrt(n)), variance within a few percent
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1.0) < 0.02
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3.0) < 0.1


def test_engines_agree_philox():
    """Same seed/offset => the lock-step engines draw identical noise; the fixed-point engine
    (its own per-pixel stream) is compared with itself: run-to-run bit-identical."""
    cfg, H, W, T = gc.CFG2, 64, 96, 10
    pos, neg, nra = gc.make_planes(13, H, W, cfg)
    outs = []
    for engine in gc.LOCKSTEP_ENGINES + ("fixpoint", "fixpoint"):
        torch.manual_seed(0)
        em = gc.cuda_emulator(cfg, engine, noise_mode="philox")
        em.philox_seed = 1234
        em.set_planes(pos, neg, nra)
        res = []
        for k, fr, tfr in gc.video(9, 3, T, H, W):
            ev, nev = em(fr.to("cuda"), tfr)
            res.append((ev.cpu().numpy(), nev, em.base_log_frame.cpu().numpy().copy()))
        outs.append(res)
    nl = len(gc.LOCKSTEP_ENGINES)
    for other in outs[1:nl]:
        for a, b in zip(outs[0], other):
            assert a[1] == b[1]
            np.testing.assert_array_equal(a[2], b[2])
            np.testing.assert_allclose(a[0], b[0], rtol=VOX_RTOL, atol=VOX_ATOL)
    for a, b in zip(outs[nl], outs[nl + 1]):
        assert a[1] == b[1]
        np.testing.assert_array_equal(a[2], b[2])
        np.testing.assert_array_equal(a[0], b[0])     # integer statistics: order-free


def test_shot_noise_rate_statistics():
    """PHILOX shot events occur at ~(rate/2)*fac*(thr/nominal) per pixel-second on pixels with
    a sub-threshold brightness change (emulator_utils.py:173-188).  A strip of strongly
    flickering pixels keeps the per-frame iteration count > 0 (no iterations, no shot events:
    v2e_model.py:449)."""
    H, W, T = 128, 128, 10
    cfg = dict(gc.CFG1, shot_noise_rate_hz=20.0, leak_rate_hz=0.0)
    for engine in gc.ENGINES:
        em = gc.cuda_emulator(cfg, engine, noise_mode="philox", debug_maps=True)
        g = torch.Generator().manual_seed(0)
        base = torch.randint(60, 200, (1, 1, H, W), generator=g).float()
        total, frames_done = 0, 0
        for k in range(20):
            flick = torch.randint(0, 2, (1, T, H, W), generator=g).float() * 4 - 2
            fr = (base + flick).round()
            fr[:, 0::2, 0, :] = 10.0       # row 0 flips between 10 and 200 every frame
            fr[:, 1::2, 0, :] = 200.0
            ev, nev = em(fr.to("cuda"), gc.stamps(k, T, 1000.0))
            fin = em.last_debug["final"].view(T - 1, H, W)
            total += int(fin[:, 1:, :].sum().item())
            frames_done += T - 1
        rate = total / ((H - 1) * W * frames_done / 1000.0)
        # intensity ~130 -> fac = 1 - 0.75*(150/275) = 0.59; thr/nominal ~ 1  => ~5.9 Hz
        exp = 10.0 * (1 - 0.75 * ((130 + 20) / 275.0))
        assert 0.8 * exp < rate < 1.2 * exp, (engine, rate, exp)


def test_full_size_1280x720_properties():
    """BASELINE configs[2] geometry: oracle parity of one steady-state pack + invariants."""
    cfg, H, W, T = gc.CFG2, 720, 1280, 10
    pos, neg, nra = gc.make_planes(17, H, W, cfg)
    em = gc.cuda_emulator(cfg, "auto", noise_mode="philox", debug_maps=True)
    orc = OracleEmulator(**cfg, threads=0)
    em.set_planes(pos, neg, nra)
    orc.set_planes(pos, neg, nra)
    for k, fr, tfr in gc.video(21, 2, T, H, W):
        em.debug_maps = True
        ev, nev = em(fr.to("cuda"), tfr)
        d = em.last_debug
        # sum of final_evts_frame over pixels and frames == num_events (checksum of checksums)
        assert int(d["final"].sum().item()) == nev
        lk = d["leak_randn"].cpu().numpy()
        sh = gc.shot_rand_from_decisions(d["shot"].cpu().numpy(), d["pol"].cpu().numpy())
        ov, onev = orc.forward(fr.numpy(), tfr, leak_randn=lk, shot_rand=sh)
        assert nev == onev
        gc.assert_state_equal(em, orc, True, f"pack {k}")
        np.testing.assert_allclose(ev.cpu().numpy(), ov, rtol=VOX_RTOL, atol=VOX_ATOL)
        # normalised grid: nonzeros have mean 0 / std 1
        v = ev[ev != 0].double()
        assert abs(v.mean().item()) < 1e-4 and abs(v.std(unbiased=False).item() - 1) < 1e-4


@pytest.mark.parametrize("fdt", [torch.float32, torch.uint8, torch.bfloat16])
def test_forward_clip_equals_pack_loop(fdt):
    """forward_clip (one C call, K launches) == K calls of forward(): states, counts and voxel
    grids bit-identical (same Philox stream: the pack counter advances the same way)."""
    cfg, H, W, T, K = gc.CFG2, 96, 160, 10, 6
    pos, neg, nra = gc.make_planes(23, H, W, cfg)
    vids = [(fr.to(fdt).cuda(), tfr) for _, fr, tfr in gc.video(8, K, T, H, W)]
    res = []
    for clip in (False, True):
        em = gc.cuda_emulator(cfg, "auto", noise_mode="philox")
        em.philox_seed = 99
        em.set_planes(pos, neg, nra)
        if clip:
            vox, nev = em.forward_clip([v[0] for v in vids[:4]], [v[1] for v in vids[:4]])
            vox2, nev2 = em.forward_clip(torch.cat([v[0] for v in vids[4:]]), [v[1] for v in vids[4:]])
            em.check_status()
            vox, nev = torch.cat([vox, vox2]), torch.cat([nev, nev2]).tolist()
        else:
            outs = [em(fr, tfr) for fr, tfr in vids]
            vox, nev = torch.cat([o[0] for o in outs]), [o[1] for o in outs]
        res.append((vox.cpu().numpy(), nev, em.base_log_frame.cpu().numpy(),
                    em.timestamp_mem.cpu().numpy(), em.t_previous))
    a, b = res
    assert a[1] == b[1] and a[4] == b[4]
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[2], b[2])
    np.testing.assert_array_equal(a[3], b[3])


@pytest.mark.parametrize("fdt", [torch.float32, torch.uint8])
def test_forward_host_equals_forward(fdt):
    """forward_host (frames in pinned host memory, staged on a copy stream, frame 0 of the
    steady-state packs not copied - it is never read, v2e_model.py:263,350) == forward on device
    frames, bit for bit, over a reset in the middle."""
    cfg, H, W, T, K = gc.CFG2, 96, 160, 10, 5
    pos, neg, nra = gc.make_planes(31, H, W, cfg)
    vids = [(fr.to(fdt), tfr) for _, fr, tfr in gc.video(14, K, T, H, W)]
    res = []
    for host in (False, True):
        em = gc.cuda_emulator(cfg, "auto", noise_mode="philox")
        em.philox_seed = 5
        em.set_planes(pos, neg, nra)
        outs = []
        for rnd in range(2):
            for fr, tfr in vids:
                if host:
                    # (the staging buffers keep a stale frame 0 for every pack but the first)
                    h = fr.clone().pin_memory()
                    outs.append(em.forward_host(h, tfr))
                else:
                    outs.append(em(fr.cuda(), tfr))
            if rnd == 0:
                em.reset()
                em.set_planes(pos, neg, nra)
        torch.cuda.synchronize()
        res.append(([o[0].cpu().numpy() for o in outs], [int(o[1]) for o in outs],
                    em.base_log_frame.cpu().numpy(), em.timestamp_mem.cpu().numpy()))
    a, b = res
    assert a[1] == b[1]
    for x, y in zip(a[0], b[0]):
        np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(a[2], b[2])
    np.testing.assert_array_equal(a[3], b[3])


@pytest.mark.parametrize("noise", ["philox", "off"])
def test_forward_clip_chained_launches(noise):
    """A long clip goes out as chains (one launch of the specialised fixed-point kernel for up to
    16 consecutive packs, v2e_clip_forward): results, per-pack event counts and the carried state
    must equal the per-pack loop bit for bit, the launch count must show the chains, and the event
    intervals must add up to the clip's device time."""
    from v2e2v_b200 import _cabi
    cfg, H, W, T, K = dict(gc.CFG2), 180, 240, 10, 24
    if noise == "off":
        cfg.update(leak_rate_hz=0, shot_noise_rate_hz=0)
    pos, neg, nra = gc.make_planes(29, H, W, cfg)
    vids = [(fr.cuda(), tfr) for _, fr, tfr in gc.video(12, K, T, H, W)]
    res = []
    for clip in (False, True):
        em = gc.cuda_emulator(cfg, "fixpoint", noise_mode=noise)
        em.philox_seed = 7
        em.set_planes(pos, neg, nra)
        n0 = _cabi.lib().v2e_launch_count()
        if clip:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K)]
            vox, nev = em.forward_clip([v[0] for v in vids], [v[1] for v in vids], events=ev)
            em.check_status()
            torch.cuda.synchronize()
            nev = nev.tolist()
            launches = _cabi.lib().v2e_launch_count() - n0
            assert launches <= 3, launches      # chunks of 1, 16, 7 packs: one chain each
            spans = [ev[2 * j].elapsed_time(ev[2 * j + 1]) for j in range(K)]
            assert all(x >= 0 for x in spans)
            total = ev[0].elapsed_time(ev[2 * K - 1])
            assert sum(spans) <= total * 1.001 + 1e-3
        else:
            outs = [em(fr, tfr) for fr, tfr in vids]
            vox, nev = torch.cat([o[0] for o in outs]), [o[1] for o in outs]
        res.append((vox.cpu().numpy(), nev, em.base_log_frame.cpu().numpy(),
                    em.lp_log_frame0.cpu().numpy(), em.timestamp_mem.cpu().numpy(), em.t_previous))
    a, b = res
    assert a[1] == b[1] and a[5] == b[5]
    for i in (0, 2, 3, 4):
        np.testing.assert_array_equal(a[i], b[i])


@pytest.mark.parametrize("H,W,K", [(720, 1280, 4), (1080, 1920, 3)])
def test_forward_clip_full_size_chain_vs_oracle(H, W, K):
    """BASELINE configs[2] at full size through the bench's path: a clip of 4 packs of 1280x720
    (launches of 1 + 3 chained packs) == the per-pack loop bit for bit, and the last pack's state
    against the CPU oracle run on the kernel's own exported draws (per-pack emulator, debug maps).
    1920x1080: 14 000 pixels per tile, more active pixels than the shared-memory table holds - the
    spill path of the specialised kernel at scale."""
    cfg, T = gc.CFG2, 10
    pos, neg, nra = gc.make_planes(41, H, W, cfg)
    vids = [(fr, tfr) for _, fr, tfr in gc.video(21, K, T, H, W)]
    em_c = gc.cuda_emulator(cfg, "auto", noise_mode="philox")
    em_p = gc.cuda_emulator(cfg, "fixpoint", noise_mode="philox", debug_maps=True)
    orc = OracleEmulator(**cfg, threads=0)
    for e in (em_c, em_p, orc):
        e.set_planes(pos, neg, nra)
    em_c.philox_seed = em_p.philox_seed = 11
    vox, nev = em_c.forward_clip([v[0].cuda() for v in vids], [v[1] for v in vids])
    em_c.check_status()
    nev = nev.tolist()
    for k, (fr, tfr) in enumerate(vids):
        v1, n1 = em_p(fr.cuda(), tfr)
        d = em_p.last_debug
        sh = gc.shot_rand_from_decisions(d["shot"].cpu().numpy(), d["pol"].cpu().numpy())
        ov, onev = orc.forward(fr.numpy(), tfr, leak_randn=d["leak_randn"].cpu().numpy(), shot_rand=sh)
        assert nev[k] == n1 == onev, k
        np.testing.assert_array_equal(vox[k].cpu().numpy(), v1[0].cpu().numpy())
        np.testing.assert_allclose(v1.cpu().numpy(), ov, rtol=VOX_RTOL, atol=VOX_ATOL)
    gc.assert_state_equal(em_p, orc, True, "per-pack vs oracle")
    gc.assert_state_equal(em_c, orc, True, "clip vs oracle")


@pytest.mark.parametrize("engine", ["auto", "persistent", "stepwise"])
def test_batch_of_sequences_equals_one_by_one(engine):
    """EventEmulatorBatch (S sequences per launch, BASELINE configs[3] geometry) == S separate
    emulators: per-sequence barriers / maxima / statistics never mix sequences."""
    from v2e2v_b200 import EventEmulator, EventEmulatorBatch
    cfg, H, W, T, S, packs = gc.CFG2, 260, 346, 10, 5, 3
    planes = [gc.make_planes(40 + s, H, W, cfg) for s in range(S)]
    vids = [list(gc.video(60 + s, packs, T, H, W)) for s in range(S)]
    batch = EventEmulatorBatch(S, device="cuda", noise_mode="philox", engine=engine, **cfg)
    singles = []
    for s in range(S):
        batch[s].set_planes(*planes[s])
        em = EventEmulator(device="cuda", noise_mode="philox", engine=engine, **cfg)
        em.philox_seed = batch[s].philox_seed
        em.set_planes(*planes[s])
        singles.append(em)
    for k in range(packs):
        fr = torch.cat([vids[s][k][1] for s in range(S)]).cuda()
        tf = torch.cat([vids[s][k][2] for s in range(S)])
        vox, nev = batch(fr, tf)
        batch.check_status()
        for s in range(S):
            v1, n1 = singles[s](vids[s][k][1].cuda(), vids[s][k][2])
            assert int(nev[s].item()) == n1, (k, s)
            np.testing.assert_array_equal(batch[s].base_log_frame.cpu().numpy(),
                                          singles[s].base_log_frame.cpu().numpy())
            np.testing.assert_array_equal(batch[s].timestamp_mem.cpu().numpy(),
                                          singles[s].timestamp_mem.cpu().numpy())
            np.testing.assert_allclose(vox[s].cpu().numpy(), v1[0].cpu().numpy(),
                                       rtol=VOX_RTOL, atol=VOX_ATOL)
    assert batch.resolved_engine == ("fixpoint" if engine == "auto" else engine)


@pytest.mark.parametrize("S,shared", [(3, True), (3, False), (18, True)])
def test_batch_forward_clip_equals_step_loop(S, shared):
    """EventEmulatorBatch.forward_clip (K steps, one C call per group of <= 16 sequences) == K
    calls of EventEmulatorBatch.forward: voxel grids, counts and states bit-identical, with one
    time base for the whole batch ([1, T] stamps, scalars shared) and with one row per sequence;
    S = 18 spans two launch groups."""
    from v2e2v_b200 import EventEmulatorBatch
    cfg, T, K = gc.CFG2, 10, 5
    H, W = (64, 96) if S > 16 else (260, 346)
    planes = [gc.make_planes(70 + s, H, W, cfg) for s in range(S)]
    # (one row per sequence: every sequence on its own time base, so that a row taken from the
    # wrong sequence shows)
    vids = [list(gc.video(90 + s, K, T, H, W, fps=1000.0 if shared else 800.0 + 150.0 * s))
            for s in range(S)]
    frames = [torch.cat([vids[s][k][1] for s in range(S)]).cuda() for k in range(K)]
    stamps = [vids[0][k][2] if shared else torch.cat([vids[s][k][2] for s in range(S)])
              for k in range(K)]
    res = []
    for clip in (False, True):
        batch = EventEmulatorBatch(S, device="cuda", noise_mode="philox", **cfg)
        for s in range(S):
            batch[s].philox_seed = 1000 + s
            batch[s].set_planes(*planes[s])
        if clip:
            v1, n1 = batch.forward_clip(frames[:2], stamps[:2])
            v2, n2 = batch.forward_clip(torch.stack(frames[2:]), stamps[2:])
            vox, nev = torch.cat([v1, v2]), torch.cat([n1, n2])
        else:
            outs = [batch(frames[k], stamps[k].expand(S, -1) if shared else stamps[k])
                    for k in range(K)]
            vox, nev = torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])
        batch.check_status()
        res.append((vox.cpu().numpy(), nev.cpu().numpy(),
                    [batch[s].base_log_frame.cpu().numpy() for s in range(S)],
                    [batch[s].timestamp_mem.cpu().numpy() for s in range(S)],
                    [batch[s].t_previous for s in range(S)]))
    a, b = res
    assert a[0].shape == (K, S, cfg["num_bins"], H, W) and a[1].shape == (K, S)
    assert a[1].sum() > 0 and a[4] == b[4]
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[0], b[0])
    for s in range(S):
        np.testing.assert_array_equal(a[2][s], b[2][s])
        np.testing.assert_array_equal(a[3][s], b[3][s])


@pytest.mark.parametrize("fdt", [torch.float32, torch.uint8])
def test_batch_forward_host_equals_forward(fdt):
    """EventEmulatorBatch.forward_host (a step's frames in pinned host memory, staged on a copy
    stream, frame 0 of steady-state packs not copied) == forward on device frames, bit for bit."""
    from v2e2v_b200 import EventEmulatorBatch
    cfg, H, W, T, S, K = gc.CFG2, 64, 96, 10, 5, 4
    planes = [gc.make_planes(50 + s, H, W, cfg) for s in range(S)]
    vids = [list(gc.video(70 + s, K, T, H, W)) for s in range(S)]
    res = []
    for host in (False, True):
        batch = EventEmulatorBatch(S, device="cuda", noise_mode="philox", **cfg)
        for s in range(S):
            batch[s].philox_seed = 300 + s
            batch[s].set_planes(*planes[s])
        outs = []
        for k in range(K):
            fr = torch.cat([vids[s][k][1] for s in range(S)]).to(fdt)
            tf = torch.cat([vids[s][k][2] for s in range(S)])
            if host:
                fr = fr.pin_memory()
                fr_keep = fr.clone()
                v, n = batch.forward_host(fr, tf)
                if k:
                    fr[:, 0] = 255 - fr[:, 0]     # frame 0 of a steady-state pack is never read
                torch.cuda.synchronize()
                del fr_keep
            else:
                v, n = batch(fr.cuda(), tf)
            outs.append((v.cpu().numpy(), n.cpu().numpy()))
        batch.check_status()
        res.append((outs, [batch[s].base_log_frame.cpu().numpy() for s in range(S)],
                    [batch[s].timestamp_mem.cpu().numpy() for s in range(S)]))
    a, b = res
    for k in range(K):
        np.testing.assert_array_equal(a[0][k][1], b[0][k][1])
        np.testing.assert_array_equal(a[0][k][0], b[0][k][0])
    for s in range(S):
        np.testing.assert_array_equal(a[1][s], b[1][s])
        np.testing.assert_array_equal(a[2][s], b[2][s])


def test_batch_full_size_64_sequences():
    """BASELINE configs[3] at full size: 64 sequences of 346x260 in one EventEmulatorBatch (four
    launch groups of 16), two steps through forward_clip.  Sampled sequences == separate
    emulators bit for bit (states, counts) / within tolerance (voxel); every sequence's normalised
    grid has mean 0 / std 1 on its non-zeros (statistics never mix sequences)."""
    from v2e2v_b200 import EventEmulator, EventEmulatorBatch
    cfg, H, W, T, S, K = gc.CFG2, 260, 346, 10, 64, 2
    planes = gc.make_planes(200, H, W, cfg)
    batch = EventEmulatorBatch(S, device="cuda", noise_mode="philox", **cfg)
    vids = [list(gc.video(300 + s, K, T, H, W)) for s in range(S)]
    for s in range(S):
        batch[s].set_planes(*planes)
    frames = [torch.cat([vids[s][k][1] for s in range(S)]).to(torch.uint8).cuda() for k in range(K)]
    vox, nev = batch.forward_clip(frames, [vids[0][k][2] for k in range(K)])
    batch.check_status()
    assert vox.shape == (K, S, cfg["num_bins"], H, W) and int(nev.min().item()) > 0
    for s in (0, 15, 16, 37, 63):
        em = EventEmulator(device="cuda", noise_mode="philox", **cfg)
        em.philox_seed = batch[s].philox_seed
        em.set_planes(*planes)
        for k in range(K):
            v1, n1 = em(vids[s][k][1].to(torch.uint8).cuda(), vids[s][k][2])
            assert int(nev[k, s].item()) == n1, (k, s)
            np.testing.assert_allclose(vox[k, s].cpu().numpy(), v1[0].cpu().numpy(),
                                       rtol=VOX_RTOL, atol=VOX_ATOL)
        np.testing.assert_array_equal(batch[s].base_log_frame.cpu().numpy(), em.base_log_frame.cpu().numpy())
        np.testing.assert_array_equal(batch[s].timestamp_mem.cpu().numpy(), em.timestamp_mem.cpu().numpy())
    v = vox[K - 1].double().reshape(S, -1)
    nz = v != 0
    mean = (v * nz).sum(1) / nz.sum(1)
    var = ((v - mean[:, None]) ** 2 * nz).sum(1) / nz.sum(1)
    assert float(mean.abs().max()) < 1e-4 and float((var.sqrt() - 1).abs().max()) < 1e-4


@pytest.mark.parametrize("T,nb,vdt", [(14, 5, torch.float32), (10, 3, torch.float32),
                                      (7, 8, torch.float32), (10, 5, torch.bfloat16)])
def test_fixpoint_other_pack_shapes(T, nb, vdt):
    """Fixed-point engine outside its specialised shape: packs longer than one Philox block /
    than the compact table serves (T = 14: every active pixel takes the spill path), other bin
    counts, bf16 voxel output - against the oracle."""
    cfg = dict(gc.CFG2, num_bins=nb, leak_rate_hz=0, shot_noise_rate_hz=0)
    H, W = 48, 80
    pos, neg, _ = gc.make_planes(31, H, W, cfg)
    em = gc.cuda_emulator(cfg, "fixpoint", voxel_dtype=vdt)
    orc = OracleEmulator(**cfg)
    em.set_planes(pos, neg)
    orc.set_planes(pos, neg)
    for k, fr, tfr in gc.video(12, 3, T, H, W):
        ev, nev = em(fr.cuda(), tfr)
        ov, onev = orc.forward(fr.numpy(), tfr)
        assert nev == onev, f"pack {k}"
        gc.assert_state_equal(em, orc, True, f"pack {k}")
        if vdt == torch.bfloat16:
            np.testing.assert_allclose(ev.float().cpu().numpy(), ov, rtol=1e-2, atol=1e-2)
        else:
            np.testing.assert_allclose(ev.cpu().numpy(), ov, rtol=VOX_RTOL, atol=VOX_ATOL)
    assert em.resolved_engine == "fixpoint"


def _run_fixpoint(cfg, noise, debug, fdt, vids, planes, seed=4242):
    em = gc.cuda_emulator(cfg, "fixpoint", noise_mode=noise, debug_maps=debug)
    em.philox_seed = seed
    em.set_planes(*planes)
    out = []
    for fr, tfr in vids:
        ev, nev = em(fr.to("cuda").to(fdt), tfr)
        out.append((ev.cpu().numpy().copy(), nev, em.base_log_frame.cpu().numpy().copy(),
                    em.lp_log_frame0.cpu().numpy().copy() if cfg["cutoff_hz"] > 0 else None,
                    em.timestamp_mem.cpu().numpy().copy()))
    return out


@pytest.mark.parametrize("noise", ["philox", "off"])
@pytest.mark.parametrize("cfgname,H,W,fdt", [
    ("CFG2", 96, 128, torch.float32),
    ("CFG2", 260, 346, torch.uint8),       # configs[3] geometry, W % 4 != 0
    ("CFG2", 180, 240, torch.bfloat16),    # configs[1]
    ("CFG1", 180, 240, torch.float32),     # configs[0] parameters (no cutoff, no diversity)
])
def test_specialised_kernel_equals_generic(noise, cfgname, H, W, fdt):
    """The kernel specialised for the reference pack shape (what serves every call without debug
    maps: v2e_fixpoint_spec.cu) against the generic kernel (debug maps on, the one the oracle
    tests above pin): same seed, same video => bit-identical states, counts and voxel grids."""
    cfg = dict(getattr(gc, cfgname))
    if noise == "off":
        cfg.update(leak_rate_hz=0, shot_noise_rate_hz=0)
    planes = gc.make_planes(5, H, W, cfg)
    vids = [(fr, tfr) for _, fr, tfr in gc.video(3, 4, 10, H, W)]
    a = _run_fixpoint(cfg, noise, True, fdt, vids, planes)
    b = _run_fixpoint(cfg, noise, False, fdt, vids, planes)
    for k, (x, y) in enumerate(zip(a, b)):
        assert x[1] == y[1], f"num_events pack {k}"
        np.testing.assert_array_equal(x[2], y[2], err_msg=f"base pack {k}")
        if x[3] is not None:
            np.testing.assert_array_equal(x[3], y[3], err_msg=f"lp pack {k}")
        np.testing.assert_array_equal(x[4], y[4], err_msg=f"tmem pack {k}")
        np.testing.assert_array_equal(x[0], y[0], err_msg=f"voxel pack {k}")


@pytest.mark.parametrize("fps,sigma", [(100.0, 6.0), (30.0, 12.0), (1000.0, 25.0)])
def test_specialised_kernel_other_regimes_vs_oracle(fps, sigma):
    """Regimes away from the 1000 fps knife edge (SURVEY trap T9): at 100 / 30 fps the refractory
    gate (v2e_model.py:469) is off for most frames and a pixel emits every event it counts; at
    sigma 25 pixels cross several thresholds per frame.  Specialised kernel (no debug maps) vs the
    oracle, noise off: states and counts bit-exact."""
    cfg = dict(gc.CFG2, leak_rate_hz=0, shot_noise_rate_hz=0)
    H, W, T = 120, 160, 10
    pos, neg, _ = gc.make_planes(9, H, W, cfg)
    em = gc.cuda_emulator(cfg, "fixpoint")
    orc = OracleEmulator(**cfg, threads=0)
    em.set_planes(pos, neg)
    orc.set_planes(pos, neg)
    for k, fr, tfr in gc.video(6, 3, T, H, W, fps=fps, sigma=sigma):
        ev, nev = em(fr.cuda(), tfr)
        ov, onev = orc.forward(fr.numpy(), tfr)
        assert nev == onev, f"pack {k}"
        gc.assert_state_equal(em, orc, True, f"pack {k}")
        np.testing.assert_allclose(ev.cpu().numpy(), ov, rtol=VOX_RTOL, atol=VOX_ATOL)


def test_fixpoint_odd_width_takes_generic_kernel():
    """An odd width (pixel pairs would straddle rows) is served by the generic kernel: same
    results as the oracle."""
    cfg = dict(gc.CFG2, leak_rate_hz=0, shot_noise_rate_hz=0)
    H, W, T = 33, 47, 10
    pos, neg, _ = gc.make_planes(3, H, W, cfg)
    em = gc.cuda_emulator(cfg, "fixpoint")
    orc = OracleEmulator(**cfg)
    em.set_planes(pos, neg)
    orc.set_planes(pos, neg)
    for k, fr, tfr in gc.video(2, 3, T, H, W):
        ev, nev = em(fr.cuda(), tfr)
        ov, onev = orc.forward(fr.numpy(), tfr)
        assert nev == onev, f"pack {k}"
        gc.assert_state_equal(em, orc, True, f"pack {k}")
        np.testing.assert_allclose(ev.cpu().numpy(), ov, rtol=VOX_RTOL, atol=VOX_ATOL)


def test_specialised_kernel_table_overflow_spills():
    """More active pixels in a tile than its table holds (a flickering image): the overflow
    takes the spill path, results still match the oracle."""
    cfg = dict(gc.CFG1, leak_rate_hz=0, shot_noise_rate_hz=0)
    H, W, T = 720, 1280, 10     # 6 227 pixels per tile on 148 SMs, nearly all of them active
    pos, neg, _ = gc.make_planes(4, H, W, cfg)
    em = gc.cuda_emulator(cfg, "fixpoint")
    orc = OracleEmulator(**cfg, threads=0)
    em.set_planes(pos, neg)
    orc.set_planes(pos, neg)
    g = torch.Generator().manual_seed(1)
    base = torch.randint(40, 200, (1, 1, H, W), generator=g).float()
    for k in range(2):
        flick = (torch.randint(0, 2, (1, T, H, W), generator=g).float() * 2 - 1) * 60
        fr = (base + flick).clamp(0, 255).round()
        tfr = gc.stamps(k, T, 1000.0)
        ev, nev = em(fr.cuda(), tfr)
        ov, onev = orc.forward(fr.numpy(), tfr)
        assert nev == onev, f"pack {k}"
        gc.assert_state_equal(em, orc, False, f"pack {k}")
        np.testing.assert_allclose(ev.cpu().numpy(), ov, rtol=VOX_RTOL, atol=VOX_ATOL)


def test_time_order_error_and_reset():
    em = gc.cuda_emulator(dict(gc.CFG2, leak_rate_hz=0, shot_noise_rate_hz=0), "auto")
    fr = torch.randint(0, 255, (1, 4, 16, 16)).float().cuda()
    t = torch.tensor([[0.0, 0.001, 0.002, 0.003]], dtype=torch.float64)
    em(fr, t)
    with pytest.raises(ValueError, match="must be later than previous"):
        em(fr, t)
    em.reset()
    assert em.base_log_frame is None
    v, n = em(fr, t)          # works again after reset
    assert v.shape == (1, 5, 16, 16)


@pytest.mark.parametrize("layout", ["f64_rows", "two_col"])
def test_forward_clip_time_order_error_mid_clip(layout):
    """A clip whose 4th pack goes back in time: like the caller's loop around the reference
    (test.py:115-121, v2e_model.py:335) the three packs before it are emulated, then ValueError
    surfaces; state, time and Philox stream continue exactly where a per-pack loop would be.
    Both preparation paths: float64 stamp rows (one C call per chunk) and 2-column stamps (the
    torch-literal path, pack by pack)."""
    cfg, H, W, T, K = gc.CFG2, 64, 96, 10, 6
    pos, neg, nra = gc.make_planes(31, H, W, cfg)
    two = layout == "two_col"
    vids = [(fr.cuda(), tfr) for _, fr, tfr in gc.video(9, K, T, H, W, two_col=two)]
    bad = [v[1] for v in vids]
    bad[3] = vids[1][1]                       # pack 3 carries the stamps of pack 1
    ems = []
    for clip in (False, True):
        em = gc.cuda_emulator(cfg, "auto", noise_mode="philox")
        em.philox_seed = 5
        em.set_planes(pos, neg, nra)
        ems.append(em)
        if clip:
            with pytest.raises(ValueError, match="must be later than previous"):
                em.forward_clip([v[0] for v in vids], bad)
        else:
            for k in range(3):
                em(vids[k][0], vids[k][1])
            with pytest.raises(ValueError, match="must be later than previous"):
                em(vids[3][0], bad[3])
    a, b = ems
    torch.cuda.synchronize()
    assert a.t_previous == b.t_previous
    np.testing.assert_array_equal(a.base_log_frame.cpu().numpy(), b.base_log_frame.cpu().numpy())
    np.testing.assert_array_equal(a.timestamp_mem.cpu().numpy(), b.timestamp_mem.cpu().numpy())
    # both carry on identically with the right pack
    va, na = a(vids[3][0], vids[3][1])
    vb, nb_ = b(vids[3][0], vids[3][1])
    assert na == nb_
    np.testing.assert_array_equal(va.cpu().numpy(), vb.cpu().numpy())


def test_forward_clip_one_more_stamp_than_frames():
    """What test.py feeds after the first pack (video_readers.py:85-86,100-101): T-1 frames with T
    stamps - frame n pairs with stamp n, the last stamp only enters Tr (v2e_model.py:317).  The C
    preparation of a clip == the per-pack loop on that layout."""
    cfg, H, W, T, K = gc.CFG2, 64, 96, 10, 5
    pos, neg, nra = gc.make_planes(32, H, W, cfg)
    vids = [(fr[:, :T - 1].contiguous().cuda(), tfr) for _, fr, tfr in gc.video(10, K, T, H, W)]
    res = []
    for clip in (False, True):
        em = gc.cuda_emulator(cfg, "auto", noise_mode="philox")
        em.philox_seed = 6
        em.set_planes(pos, neg, nra)
        if clip:
            vox, nev = em.forward_clip([v[0] for v in vids], [v[1] for v in vids])
            em.check_status()
            nev = nev.tolist()
        else:
            outs = [em(fr, tfr) for fr, tfr in vids]
            vox, nev = torch.cat([o[0] for o in outs]), [o[1] for o in outs]
        res.append((vox.cpu().numpy(), nev, em.base_log_frame.cpu().numpy(),
                    em.timestamp_mem.cpu().numpy(), em.t_previous))
    a, b = res
    assert a[1] == b[1] and sum(a[1]) > 0 and a[4] == b[4]
    for i in (0, 2, 3):
        np.testing.assert_array_equal(a[i], b[i])


def test_b_gt_1_rejected_and_cpu_tensor_rejected():
    em = gc.cuda_emulator(dict(gc.CFG2, leak_rate_hz=0, shot_noise_rate_hz=0), "auto")
    t = torch.tensor([[0.0, 0.003]], dtype=torch.float64)
    with pytest.raises(NotImplementedError):
        em(torch.zeros(2, 4, 16, 16, device="cuda"), t)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        em(torch.zeros(1, 4, 16, 16), t)


def test_non_integer_frames_slow_path():
    """Non-integer frame values take the fp64 lin_log path and still match the oracle."""
    cfg = dict(gc.CFG2, leak_rate_hz=0, shot_noise_rate_hz=0)
    H, W, T = 32, 48, 6
    pos, neg, _ = gc.make_planes(5, H, W, cfg)
    for engine in gc.ENGINES:
        em = gc.cuda_emulator(cfg, engine)
        orc = OracleEmulator(**cfg)
        em.set_planes(pos, neg)
        orc.set_planes(pos, neg)
        g = torch.Generator().manual_seed(3)
        for k in range(2):
            fr = torch.rand(1, T, H, W, generator=g) * 255     # arbitrary float values
            tfr = gc.stamps(k, T, 1000.0)
            ev, nev = em(fr.cuda(), tfr)
            ov, onev = orc.forward(fr.numpy(), tfr)
            assert nev == onev
            gc.assert_state_equal(em, orc, True)


def test_runahead_integer_frames_contract():
    """integer_frames=True selects the run-ahead variant without the fp64 lin_log path: same
    results on integer-valued frames, and a violation is detected and raised, never served."""
    cfg = dict(gc.CFG2, leak_rate_hz=0, shot_noise_rate_hz=0)
    H, W, T = 64, 96, 10
    pos, neg, _ = gc.make_planes(7, H, W, cfg)
    em = gc.cuda_emulator(cfg, "runahead", integer_frames=True)
    orc = OracleEmulator(**cfg)
    em.set_planes(pos, neg)
    orc.set_planes(pos, neg)
    last = None
    for k, fr, tfr in gc.video(4, 3, T, H, W):
        ev, nev = em(fr.cuda(), tfr)
        ov, onev = orc.forward(fr.numpy(), tfr)
        assert nev == onev
        gc.assert_state_equal(em, orc, True, f"pack {k}")
        np.testing.assert_allclose(ev.cpu().numpy(), ov, rtol=VOX_RTOL, atol=VOX_ATOL)
        last = fr
    bad = last.clone()
    bad[0, 3, 5, 7] += 0.25
    with pytest.raises(ValueError, match="not an integer"):
        em(bad.cuda(), gc.stamps(3, T, 1000.0))


def test_auto_engine_is_fastest_known():
    em = gc.cuda_emulator(dict(gc.CFG2, leak_rate_hz=0, shot_noise_rate_hz=0), "auto")
    em(torch.zeros(1, 4, 64, 64, device="cuda"), torch.tensor([[0.0, 0.003]], dtype=torch.float64))
    assert em.resolved_engine == "fixpoint"
    raw = gc.cuda_emulator(dict(gc.CFG2, leak_rate_hz=0, shot_noise_rate_hz=0, output_mode="raw"),
                           "auto")
    raw(torch.zeros(1, 4, 64, 64, device="cuda"), torch.tensor([[0.0, 0.003]], dtype=torch.float64))
    assert raw.resolved_engine == "persistent"

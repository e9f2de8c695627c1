"""Pins the CPU oracle (oracle/v2e_oracle.c) to outputs of the unmodified reference.

The fixtures in tests/golden/ were produced by running the reference `EventEmulator`
(v2e/v2e_model.py:290-536) itself (tests/golden/make_golden.py).  The bar: state planes
(base_log_frame, lp_log_frame0, timestamp_mem) and num_events bit-exact after every pack, raw
event rows exact as a multiset, voxel grid rtol 1e-5 / atol 1e-6 (the reference's f32 sums in
event_preprocess_pytorch have unspecified order).
"""
import numpy as np
import pytest

from oracle.oracle import OracleEmulator
from tests.golden.cases import CASES
from tests.golden_io import GoldenCase

VOX_RTOL, VOX_ATOL = 1e-5, 1e-6


def make_oracle(g: GoldenCase, **kw) -> OracleEmulator:
    emu = {k: v for k, v in g.emu.items()}
    em = OracleEmulator(**emu, **kw)
    em.set_planes(*g.planes())
    return em


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("use_stored_scalars", [False, True])
def test_oracle_matches_reference(name, use_stored_scalars, oracle_lib):
    g = GoldenCase(name)
    em = make_oracle(g)
    for k in range(g.packs):
        frames, t_frames, lk, sh, sc = g.pack_inputs(k)
        ev, nev = em.forward(frames, t_frames, leak_randn=lk, shot_rand=sh,
                             scalars=sc if use_stored_scalars else None)
        want = g.pack_outputs(k)
        assert nev == want["num_events"], f"pack {k}"
        np.testing.assert_array_equal(em.base_log_frame, want["base"], err_msg=f"base pack {k}")
        if g.emu["cutoff_hz"] > 0:
            np.testing.assert_array_equal(em.lp_log_frame0, want["lp"], err_msg=f"lp pack {k}")
        np.testing.assert_array_equal(em.timestamp_mem, want["tmem"], err_msg=f"tmem pack {k}")
        if g.raw:
            np.testing.assert_array_equal(ev, want["events"])
        else:
            np.testing.assert_allclose(ev.reshape(g.emu["num_bins"], -1), want["voxel"],
                                       rtol=VOX_RTOL, atol=VOX_ATOL)


def test_oracle_threads_invariant(oracle_lib):
    """OpenMP and serial runs of the oracle are bit-identical (per-pixel independence)."""
    g = GoldenCase("cfg2_sd_noise")
    a, b = make_oracle(g, threads=1), make_oracle(g, threads=4)
    for k in range(g.packs):
        frames, t_frames, lk, sh, sc = g.pack_inputs(k)
        va, na = a.forward(frames, t_frames, leak_randn=lk, shot_rand=sh)
        vb, nb_ = b.forward(frames, t_frames, leak_randn=lk, shot_rand=sh)
        assert na == nb_
        np.testing.assert_array_equal(va, vb)
        np.testing.assert_array_equal(a.base_log_frame, b.base_log_frame)
        np.testing.assert_array_equal(a.timestamp_mem, b.timestamp_mem)


def test_oracle_frame_dtypes_agree(oracle_lib):
    """u8 / f32 / bf16 frame inputs give identical results on integer-valued frames."""
    g = GoldenCase("cfg2_sd_noiseless")
    ems = [make_oracle(g) for _ in range(3)]
    for k in range(g.packs):
        frames, t_frames, lk, sh, sc = g.pack_inputs(k)
        f32 = frames.astype(np.float32)
        bf16 = (f32.view(np.uint32) >> 16).astype(np.uint16)   # exact: integers <= 255
        outs = [em.forward(x, t_frames, scalars=sc) for em, x in zip(ems, (frames, f32, bf16))]
        for o in outs[1:]:
            assert o[1] == outs[0][1]
            np.testing.assert_array_equal(o[0], outs[0][0])

"""Host scalar preparation (csrc/host_scalars.c) == the reference's torch expressions, bitwise.

The torch-literal restatement lives in oracle/host_scalars_torch.py (itself pinned to the live
reference by the golden fixtures: the oracle reproduces the reference only if these scalars are
right).  Covers the knife-edge configuration of SURVEY.md trap T9 explicitly.
"""
import math

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle.host_scalars_torch import pack_scalars_torch
from tests.golden.cases import CASES
from tests.golden_io import GoldenCase, scalars_equal
from v2e2v_b200 import build
from v2e2v_b200.host_scalars import TimeState, pack_scalars


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build()


def run_both(stamp_list, T, nb, refr, cutoff, ql, qs, rate):
    """Feed a sequence of packs through both implementations; compare every pack."""
    tp_c = TimeState()
    tp_t = None
    for k, tfr in enumerate(stamp_list):
        first = k == 0
        a = pack_scalars(tfr, T, nb, refr, cutoff, ql, qs, rate, tp_c, first)
        b, tp_t = pack_scalars_torch(tfr, T, nb, refr, cutoff, ql, qs, rate, tp_t, first)
        assert scalars_equal(a, b), f"pack {k}"
        assert tp_c.value == float(tp_t)
        assert tp_c.is_f64 == (tp_t.dtype == torch.float64)


@pytest.mark.parametrize("name", list(CASES))
def test_matches_golden_scalars(name):
    """The C implementation reproduces the tables stored with the reference runs."""
    g = GoldenCase(name)
    tp = TimeState()
    for k in range(g.packs):
        frames, t_frames, _, _, sc = g.pack_inputs(k)
        e = g.emu
        got = pack_scalars(t_frames, frames.shape[0], e["num_bins"], e["refractory_period_s"],
                           e["cutoff_hz"], e["ql"], e["qs"], e["shot_noise_rate_hz"], tp, k == 0)
        assert scalars_equal(got, sc), f"{name} pack {k}"


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("two_col", [False, True])
def test_knife_edge_1000fps(dtype, two_col):
    """1000 fps, T=10, 1 ms refractory: Tr and duration/num_iters are mathematically equal."""
    T, nb = 10, 5
    packs = []
    for k in range(12):
        t = torch.linspace(k * 9 / 1000, (k + 1) * 9 / 1000, T, dtype=torch.float64).to(dtype)
        packs.append((t[[0, -1]] if two_col else t)[None])
    run_both(packs, T, nb, 0.001, 200.0, 1, 0, 1.0)


@settings(max_examples=150, deadline=None)
@given(T=st.integers(2, 16), nb=st.integers(2, 8), fps=st.floats(5.0, 5000.0),
       t0=st.floats(0.0, 1000.0), f64=st.booleans(), two_col=st.booleans(),
       refr=st.floats(0.0, 0.02), cutoff=st.sampled_from([0.0, 15.0, 200.0]),
       ql=st.sampled_from([0.0, 0.5, 1.0]), qs=st.sampled_from([0.0, 0.3, 1.0]),
       rate=st.sampled_from([0.0, 1.0, 5.5]), jitter=st.floats(0.0, 0.3), seed=st.integers(0, 99))
def test_random_stamps(T, nb, fps, t0, f64, two_col, refr, cutoff, ql, qs, rate, jitter, seed):
    rng = np.random.default_rng(seed)
    dtype = torch.float64 if f64 else torch.float32
    packs = []
    t = t0
    for k in range(3):
        gaps = (1.0 + jitter * (rng.random(T - 1) - 0.5)) / fps
        stamps = t + np.concatenate([[0.0], np.cumsum(gaps)])
        t = stamps[-1]
        tt = torch.tensor(stamps, dtype=torch.float64).to(dtype)
        if not f64 and not bool((tt[1:] > tt[:-1]).all()):
            return  # f32 resolution exhausted at this t0/fps: not a valid input
        packs.append((tt[[0, -1]] if two_col else tt)[None])
    try:
        run_both(packs, T, nb, refr, cutoff, ql, qs, rate)
    except ValueError:
        # both sides must agree on the time-order error as well
        tp = TimeState()
        with pytest.raises(ValueError):
            for k, tfr in enumerate(packs):
                pack_scalars(tfr, T, nb, refr, cutoff, ql, qs, rate, tp, k == 0)


def test_time_order_error():
    tp = TimeState()
    t = torch.tensor([[0.0, 0.001, 0.002]], dtype=torch.float64)
    pack_scalars(t, 3, 5, 0.001, 0.0, 1, 1, 0.0, tp, True)
    with pytest.raises(ValueError, match="must be later than previous"):
        pack_scalars(t, 3, 5, 0.001, 0.0, 1, 1, 0.0, tp, False)   # same stamps again


def test_ts_step_formula():
    """duration/num_iters on an int32 tensor == rcp(f32(ni)) * f32(duration)  (trap T6)."""
    for T, nb in [(10, 5), (7, 5), (16, 3), (2, 5)]:
        duration = (nb - 1) / (T - 1)
        for ni in range(1, 300):
            want = float((duration / torch.tensor([ni], dtype=torch.int32))[0])
            got = np.float32(np.float32(1.0) / np.float32(ni)) * np.float32(duration)
            assert np.float32(want) == got


@pytest.mark.parametrize("first", [True, False])
def test_clip_prepare_equals_pack_loop(first):
    """v2e_host_clip_prepare (K packs of a sequence in one host call) == K calls of
    v2e_host_pack_scalars with t_previous carried along, plus the per-pack pointers."""
    import ctypes as C
    from v2e2v_b200 import _cabi
    from v2e2v_b200.host_scalars import PackScalarPrep
    T, nb, K, n_st = 10, 5, 7, 11
    prep = PackScalarPrep(T, nb, 0.001, 200.0, 1.0, 0.0, 1.0)
    rows = [torch.linspace(k * 0.009 + 0.5, (k + 1) * 0.009 + 0.5, n_st, dtype=torch.float64)[None]
            for k in range(K)]
    tp = TimeState(0.25, False)
    want = []
    tp_loop = TimeState(tp.value, tp.is_f64)
    for k in range(K):
        sc = _cabi.PackScalars()
        prep.fill(rows[k], tp_loop, first and k == 0, sc)
        want.append(bytes(sc))
    tmpl = _cabi.Pack()
    tmpl.base_log_frame, tmpl.voxel, tmpl.num_events, tmpl.status = 0x1000, 0x200000, 0x3000, 0x3008
    tmpl.philox_seed, tmpl.philox_offset = 77, 5
    hp = prep.hp
    hp.n_stamps, hp.first_pack, hp.t_prev, hp.t_prev_f64 = n_st, int(first), tp.value, int(tp.is_f64)
    out = (_cabi.Pack * K)()
    fptrs = (C.c_void_p * K)(*[0x900000 + 64 * k for k in range(K)])
    rptrs = (C.c_void_p * K)(*[r.data_ptr() for r in rows])
    tprev, bad = C.c_double(0), C.c_int32(-1)
    rc = _cabi.lib().v2e_host_clip_prepare(C.byref(hp), rptrs, 0, prep._tfr_ptr, C.byref(tmpl), fptrs, 0,
                                           4096, 16, K, 1, C.byref(tprev), C.byref(bad), out)
    assert rc == _cabi.OK and tprev.value == tp_loop.value
    for k in range(K):
        assert bytes(out[k].sc) == want[k], f"pack {k}"
        assert out[k].frames == 0x900000 + 64 * k and out[k].voxel == 0x200000 + 4096 * k
        assert out[k].num_events == 0x3000 + 16 * k and out[k].status == 0x3008 + 16 * k
        assert out[k].philox_offset == 5 + k and out[k].philox_seed == 77
        assert out[k].base_log_frame == 0x1000
    # one sequence of a batch: records interleaved (out_stride), frames / rows / meta by offset
    out2 = (_cabi.Pack * (3 * K))()
    big = torch.stack([torch.cat([r, r + 100.0]) for r in rows])      # [K, 2, n_st]: row 1 = +100 s
    rp2 = (C.c_void_p * K)(*[big[k].data_ptr() for k in range(K)])
    hp.t_prev = 50.0
    rc = _cabi.lib().v2e_host_clip_prepare(C.byref(hp), rp2, n_st * 8, prep._tfr_ptr, C.byref(tmpl), fptrs, 32,
                                           4096, 48, K, 3, C.byref(tprev), C.byref(bad),
                                           C.cast(C.addressof(out2) + C.sizeof(_cabi.Pack), C.POINTER(_cabi.Pack)))
    assert rc == _cabi.OK and tprev.value == float(big[K - 1, 1, T - 1])
    for k in range(K):
        o = out2[3 * k + 1]
        assert o.frames == 0x900000 + 64 * k + 32 and o.num_events == 0x3000 + 48 * k
        assert o.sc.num_frames == T and out2[3 * k].sc.num_frames == 0
    hp.t_prev = tp.value
    # a clip that goes back in time is refused at the offending pack, nothing is reported done
    rptrs[3] = rows[1].data_ptr()
    tprev.value = -1.0
    rc = _cabi.lib().v2e_host_clip_prepare(C.byref(hp), rptrs, 0, prep._tfr_ptr, C.byref(tmpl), fptrs, 0,
                                           4096, 16, K, 1, C.byref(tprev), C.byref(bad), out)
    assert rc == _cabi.ERR_TIME_ORDER and bad.value == 3 and tprev.value == -1.0

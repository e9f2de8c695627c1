"""Standalone voxeliser / normaliser on the GPU (v2e_events_to_voxel, v2e_voxel_normalize through
v2e2v_b200.event_process, the drop-ins for utils/event_process.py:15-177) against the reference's
own outputs (tests/golden/voxel_cases.npz) and against the CPU oracle at data-reader sizes.

The CUDA scatter accumulates with L2 reductions in no particular order, the reference in event
order: grids agree to f32 rounding of the sums (rtol 1e-5, atol 2e-6), not bit for bit.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "voxel_cases.npz"))
TAGS = sorted({k.split("_")[0] for k in G.files})
TOL = dict(rtol=1e-5, atol=2e-6)


def seeded_events(seed, n, W, H, dtype=np.float64, t0=3.0, span=0.04):
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.sort(rng.random(n)) * span + t0
    return np.stack([t, rng.integers(0, W, n).astype(np.float64), rng.integers(0, H, n).astype(np.float64),
                     rng.integers(0, 2, n).astype(np.float64)], 1).astype(dtype).reshape(n, 4)


@pytest.mark.parametrize("tag", TAGS)
def test_events_to_voxel_matches_reference_outputs(tag):
    from v2e2v_b200 import event_process as ep
    ev = G[f"{tag}_events"]
    nb, W, H = (int(v) for v in G[f"{tag}_dims"])
    before = ev.copy()
    v_np = ep.events_to_voxel_grid(ev, nb, W, H)                       # numpy in -> numpy out
    np.testing.assert_array_equal(ev, before)                          # (not mutated, unlike :40,46)
    assert isinstance(v_np, np.ndarray) and v_np.shape == (nb, H, W) and v_np.dtype == np.float32
    np.testing.assert_allclose(v_np, G[f"{tag}_voxel_np"], **TOL)
    v_t = ep.events_to_voxel_grid_pytorch(torch.from_numpy(ev.copy()), nb, W, H)
    assert v_t.is_cuda
    np.testing.assert_allclose(v_t.cpu().numpy(), G[f"{tag}_voxel_torch"], **TOL)


@pytest.mark.parametrize("tag", TAGS)
def test_event_preprocess_matches_reference_outputs(tag):
    from v2e2v_b200 import event_process as ep
    v_np, v_t = G[f"{tag}_voxel_np"], G[f"{tag}_voxel_torch"]
    keep = v_np.copy()
    np.testing.assert_allclose(ep.event_preprocess(v_np, 'std', False), G[f"{tag}_norm_np"], **TOL)
    np.testing.assert_allclose(ep.event_preprocess(v_np, 'std', True), G[f"{tag}_norm_np_hot"], **TOL)
    np.testing.assert_allclose(ep.event_preprocess(v_np, 'maxmin', False), G[f"{tag}_maxmin_np"], **TOL)
    np.testing.assert_array_equal(v_np, keep)
    vt = torch.from_numpy(v_t.copy()).cuda()
    np.testing.assert_allclose(ep.event_preprocess_pytorch(vt.clone(), 'std', False).cpu().numpy(),
                               G[f"{tag}_norm_torch"], **TOL)
    np.testing.assert_allclose(ep.event_preprocess_pytorch(vt.clone(), 'std', True).cpu().numpy(),
                               G[f"{tag}_norm_torch_hot"], **TOL)


@pytest.mark.parametrize("n,W,H,dt", [(1_000_000, 346, 260, np.float64), (300_000, 240, 180, np.float32),
                                      (2_000_000, 1280, 720, np.float64)])
@pytest.mark.parametrize("torch_sem", [False, True])
def test_events_to_voxel_vs_oracle_reader_sizes(n, W, H, dt, torch_sem):
    """The sizes the data readers voxelise (video_readers.py:159-179): CUDA == oracle within the
    tolerance; event conservation: every event adds its polarity, split over two bins."""
    from v2e2v_b200 import event_process as ep
    nb = 5
    ev = seeded_events(11 + n % 7, n, W, H, dt)
    want = orc.events_to_voxel(ev, nb, W, H, torch_semantics=torch_sem)
    if torch_sem:
        got = ep.events_to_voxel_grid_pytorch(torch.from_numpy(ev), nb, W, H).cpu().numpy()
    else:
        got = ep.events_to_voxel_grid(ev, nb, W, H)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    # (ts <= nb-1, so no event has a right-hand share beyond the last bin: the grid sums to Σ pol)
    pol = np.where(ev[:, 3] == 0, -1.0, 1.0)
    assert abs(float(got.astype(np.float64).sum()) - float(pol.sum())) < 1e-3 * n ** 0.5
    # normalise on the GPU == oracle
    gn = ep.event_preprocess(got, 'std', True)
    wn = orc.voxel_normalize(got, 'std', 25. / nb)
    np.testing.assert_allclose(gn, wn, rtol=1e-5, atol=1e-5)
    nz = gn[gn != 0].astype(np.float64)
    assert abs(nz.mean()) < 1e-3 and abs(nz.std() - 1) < 1e-3


def test_events_to_voxel_edge_cases():
    from v2e2v_b200 import event_process as ep
    nb, W, H = 5, 16, 12
    # empty list -> zero grid (:29-30)
    z = ep.events_to_voxel_grid(np.zeros((0, 4)), nb, W, H)
    assert z.shape == (nb, H, W) and not z.any()
    # one event, and all events at one stamp: deltaT == 0 -> 1 (:36-37), everything lands in bin 0
    one = np.array([[7.5, 3, 4, 1.0]])
    v = ep.events_to_voxel_grid(one, nb, W, H)
    assert v[0, 4, 3] == 1.0 and v.sum() == 1.0
    same = np.array([[2.0, 1, 1, 0.0], [2.0, 1, 1, 0.0], [2.0, 5, 2, 1.0]])
    v = ep.events_to_voxel_grid(same, nb, W, H)
    assert v[0, 1, 1] == -2.0 and v[0, 2, 5] == 1.0            # polarity 0 counts as -1 (:46)
    np.testing.assert_array_equal(v, orc.events_to_voxel(same, nb, W, H))
    # an all-zero grid is returned unchanged by the normaliser (:143 / :168)
    assert not ep.event_preprocess(np.zeros((nb, H, W), np.float32), 'std', True).any()
    # CPU tensors are moved, never computed on the host: the op itself rejects them
    from v2e2v_b200 import ops
    with pytest.raises(RuntimeError):
        ops.events_to_voxel_op(torch.zeros(4, 4), nb, W, H, False)

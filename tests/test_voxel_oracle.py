"""Pins the oracle's standalone voxel functions (oracle/v2e_oracle.c: events -> voxel grid,
normalisation) to outputs of the unmodified reference (utils/event_process.py:15-177), stored in
tests/golden/voxel_cases.npz by tests/golden/make_golden.py.

np.add.at / index_add_ (CPU) accumulate in event order in f32 - left bins first, then right bins -
as the oracle does, so the grids are compared bit for bit; the normalised grids within rtol 1e-5 (f32 sums of unspecified order).
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "voxel_cases.npz"))
TAGS = sorted({k.split("_")[0] for k in G.files})


@pytest.mark.parametrize("tag", TAGS)
def test_events_to_voxel_matches_reference(tag, oracle_lib):
    ev = G[f"{tag}_events"]
    nb, W, H = (int(v) for v in G[f"{tag}_dims"])
    before = ev.copy()
    v_np = orc.events_to_voxel(ev, nb, W, H, torch_semantics=False)
    v_t = orc.events_to_voxel(ev, nb, W, H, torch_semantics=True)
    np.testing.assert_array_equal(ev, before)            # (the reference mutates; the oracle does not)
    np.testing.assert_array_equal(v_np, G[f"{tag}_voxel_np"])
    np.testing.assert_array_equal(v_t, G[f"{tag}_voxel_torch"])
    assert v_np.shape == (nb, H, W) and v_np.dtype == np.float32


@pytest.mark.parametrize("tag", TAGS)
def test_voxel_normalize_matches_reference(tag, oracle_lib):
    nb = int(G[f"{tag}_dims"][0])
    v_np, v_t = G[f"{tag}_voxel_np"], G[f"{tag}_voxel_torch"]
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(orc.voxel_normalize(v_np, 'std', 0.0), G[f"{tag}_norm_np"], **tol)
    np.testing.assert_allclose(orc.voxel_normalize(v_np, 'std', 25. / nb), G[f"{tag}_norm_np_hot"], **tol)
    np.testing.assert_allclose(orc.voxel_normalize(v_np, 'maxmin', 0.0), G[f"{tag}_maxmin_np"], **tol)
    np.testing.assert_allclose(orc.voxel_normalize(v_t, 'std', 0.0), G[f"{tag}_norm_torch"], **tol)
    np.testing.assert_allclose(orc.voxel_normalize(v_t, 'std', 20. / nb), G[f"{tag}_norm_torch_hot"], **tol)

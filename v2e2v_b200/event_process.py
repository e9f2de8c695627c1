"""Drop-ins for utils/event_process.py of the reference, backed by the CUDA library.

  events_to_voxel_grid          utils/event_process.py:15-63   (numpy semantics)
  events_to_voxel_grid_pytorch  utils/event_process.py:66-129  (torch semantics)
  event_preprocess              utils/event_process.py:132-154 (hot pixel threshold 25/num_bins)
  event_preprocess_pytorch      utils/event_process.py:157-177 (hot pixel threshold 20/num_bins)

Inputs may be numpy arrays or torch tensors on any device; the work runs on `device` (default
"cuda") and there is no CPU fallback.  Unlike the reference, the caller's `events` array is NOT
mutated (the reference rescales column 0 and rewrites polarity 0 -> -1 in place).
Accumulation order differs from np.add.at / index_add_ (L2 reductions), so voxel values agree
to f32 rounding (rtol 1e-5), not bit-for-bit.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _to_dev(a, device, dtype=None):
    t = torch.as_tensor(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device)


def events_to_voxel_grid(events, num_bins, width, height, device="cuda"):
    """[N,4] rows (t,x,y,p) -> float32 voxel grid [num_bins,H,W]; numpy in -> numpy out."""
    assert num_bins > 0 and width > 0 and height > 0
    was_numpy = isinstance(events, np.ndarray)
    ev = _to_dev(events, device)
    assert ev.dim() == 2 and ev.shape[1] == 4
    vox = ops.events_to_voxel_op(ev, num_bins, width, height, False)
    return vox.cpu().numpy() if was_numpy else vox


def events_to_voxel_grid_pytorch(events, num_bins, width, height):
    """torch variant: events is a tensor (moved to CUDA if needed); returns a CUDA tensor."""
    assert num_bins > 0 and width > 0 and height > 0
    ev = events if events.is_cuda else events.cuda()
    assert ev.dim() == 2 and ev.shape[1] == 4
    with torch.no_grad():
        return ops.events_to_voxel_op(ev, num_bins, width, height, True)


def _preprocess(voxel, mode, thresh, device):
    was_numpy = isinstance(voxel, np.ndarray)
    v = _to_dev(voxel, device, torch.float32).contiguous()
    if not was_numpy and v.data_ptr() == voxel.data_ptr():
        v = v.clone()                      # the reference returns a new tensor in 'std' mode
    ops.voxel_normalize_op(v, mode, thresh)
    return v.cpu().numpy() if was_numpy else v


def event_preprocess(event_voxel_grid, mode='std', filter_hot_pixel=False, device="cuda"):
    nb = event_voxel_grid.shape[0]
    return _preprocess(event_voxel_grid, mode, 25. / nb if filter_hot_pixel else 0.0, device)


def event_preprocess_pytorch(event_voxel_grid, mode='std', filter_hot_pixel=True):
    nb = event_voxel_grid.shape[0]
    dev = event_voxel_grid.device if event_voxel_grid.is_cuda else "cuda"
    return _preprocess(event_voxel_grid, mode, 20. / nb if filter_hot_pixel else 0.0, dev)

"""torch custom-op layer over the C ABI (include/v2e2v_b200.h).

The ops registered here (namespace ``v2e2v``) are thin: they pull ``data_ptr()`` and the current
CUDA stream out of the tensors and call the in-tree CUDA library through ctypes.  There is no
CPU implementation and no fallback: a CPU tensor reaching an op raises.

  v2e2v::emulate_pack      -> v2e_pack_forward        (EventEmulator.forward, v2e/v2e_model.py:290)
  v2e2v::events_to_voxel   -> v2e_events_to_voxel     (utils/event_process.py:15-129)
  v2e2v::voxel_normalize   -> v2e_voxel_normalize     (utils/event_process.py:132-177)
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _cabi

_FRAME_DT = {torch.float32: _cabi.DT_F32, torch.uint8: _cabi.DT_U8, torch.bfloat16: _cabi.DT_BF16}
_VOX_DT = {torch.float32: _cabi.DT_F32, torch.bfloat16: _cabi.DT_BF16}


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("v2e2v_b200 ops run on CUDA tensors only (there is no CPU fallback)")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# ---------------------------------------------------------------------------------------------
# emulate_pack: the launch record is a plain Python object (ctypes structs); the registered op
# below takes the tensors it reads / mutates so that torch sees the data dependencies.
# ---------------------------------------------------------------------------------------------
class PackLaunch:
    """Reusable launch record for up to 16 sequences sharing one v2e_config."""

    def __init__(self, cfg: _cabi.Config, n_seq: int):
        self.cfg = cfg
        self.n_seq = n_seq
        self.packs = (_cabi.Pack * n_seq)()
        self._lib = _cabi.lib()
        self.workspace_bytes = int(self._lib.v2e_workspace_bytes(C.byref(cfg), n_seq))
        if self.workspace_bytes == 0:
            raise _cabi.V2EError(_cabi.ERR_BAD_ARG, "v2e_workspace_bytes", "invalid config")

    def resolve_engine(self) -> int:
        return int(self._lib.v2e_resolve_engine(C.byref(self.cfg), self.n_seq))

    def launch(self, workspace: torch.Tensor, stream_ptr: int) -> None:
        rc = self._lib.v2e_pack_forward(C.byref(self.cfg), self.packs, self.n_seq,
                                        workspace.data_ptr(), workspace.numel(), stream_ptr)
        _cabi.check(rc, "v2e_pack_forward")


_LAUNCHES = {}
_next_handle = [1]


def register_launch(pl: PackLaunch) -> int:
    h = _next_handle[0]
    _next_handle[0] += 1
    _LAUNCHES[h] = pl
    return h


def release_launch(h: int) -> None:
    _LAUNCHES.pop(h, None)


_lib_def = torch.library.Library("v2e2v", "DEF")
_lib_def.define(
    "emulate_pack(int handle, Tensor frames, Tensor(a!) base, Tensor(b!) lp, Tensor(c!) tmem, "
    "Tensor(d!) out, Tensor(e!) num_events, Tensor(f!) workspace) -> ()")
_lib_def.define(
    "events_to_voxel(Tensor events, int num_bins, int width, int height, bool torch_semantics) "
    "-> Tensor")
_lib_def.define(
    "voxel_normalize(Tensor(a!) voxel, str mode, float hot_pixel_thresh) -> ()")


def _emulate_pack_cuda(handle, frames, base, lp, tmem, out, num_events, workspace):
    _require_cuda(frames, base, tmem, out, num_events, workspace)
    _LAUNCHES[handle].launch(workspace, _stream_ptr(frames.device))


def _events_to_voxel_cuda(events, num_bins, width, height, torch_semantics):
    _require_cuda(events)
    if events.numel() and (events.dim() != 2 or events.size(1) != 4):
        raise ValueError("events must be [N,4] rows (t,x,y,p)")
    if events.dtype not in (torch.float32, torch.float64):
        events = events.to(torch.float64)
    events = events.contiguous()
    vox = torch.empty((num_bins, height, width), dtype=torch.float32, device=events.device)
    rc = _cabi.lib().v2e_events_to_voxel(
        _ptr(events) if events.numel() else None, int(events.dtype == torch.float64),
        events.size(0) if events.dim() == 2 else 0, num_bins, width, height, 0, 0.0, 0.0,
        int(torch_semantics), vox.data_ptr(), _stream_ptr(events.device))
    _cabi.check(rc, "v2e_events_to_voxel")
    return vox


_NORM_WS = {}


def _voxel_normalize_cuda(voxel, mode, hot_pixel_thresh):
    _require_cuda(voxel)
    if voxel.dtype != torch.float32 or not voxel.is_contiguous():
        raise ValueError("voxel must be a contiguous float32 tensor")
    key = (voxel.device, torch.cuda.current_stream(voxel.device).cuda_stream)
    ws = _NORM_WS.get(key)
    if ws is None:
        ws = _NORM_WS[key] = torch.empty(40960, dtype=torch.uint8, device=voxel.device)
    rc = _cabi.lib().v2e_voxel_normalize(voxel.data_ptr(), voxel.numel(),
                                         {"std": 0, "maxmin": 1}[mode], float(hot_pixel_thresh),
                                         ws.data_ptr(), _stream_ptr(voxel.device))
    _cabi.check(rc, "v2e_voxel_normalize")


_lib_def.impl("emulate_pack", _emulate_pack_cuda, "CUDA")
_lib_def.impl("events_to_voxel", _events_to_voxel_cuda, "CUDA")
_lib_def.impl("voxel_normalize", _voxel_normalize_cuda, "CUDA")


def _no_cpu(*a, **k):
    raise RuntimeError("v2e2v_b200 has no CPU implementation: move the tensors to a B200")


for _name in ("emulate_pack", "events_to_voxel", "voxel_normalize"):
    _lib_def.impl(_name, _no_cpu, "CPU")


@torch.library.register_fake("v2e2v::events_to_voxel")
def _events_to_voxel_fake(events, num_bins, width, height, torch_semantics):
    return events.new_empty((num_bins, height, width), dtype=torch.float32)


@torch.library.register_fake("v2e2v::emulate_pack")
def _emulate_pack_fake(handle, frames, base, lp, tmem, out, num_events, workspace):
    return None


@torch.library.register_fake("v2e2v::voxel_normalize")
def _voxel_normalize_fake(voxel, mode, hot_pixel_thresh):
    return None


emulate_pack = torch.ops.v2e2v.emulate_pack
events_to_voxel_op = torch.ops.v2e2v.events_to_voxel
voxel_normalize_op = torch.ops.v2e2v.voxel_normalize

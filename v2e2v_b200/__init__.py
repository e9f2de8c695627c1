"""v2e2v_b200 — B200-native V2E (video -> events) emulator and voxel-grid path of V2E2V.

Public surface (mirrors the reference's):
  EventEmulator            drop-in for v2e.v2e_model.EventEmulator
  EventEmulatorBatch       S independent sequences (one EventEmulator each) in shared launches
  events_to_voxel_grid*    drop-ins for utils.event_process.events_to_voxel_grid[_pytorch]
  event_preprocess*        drop-ins for utils.event_process.event_preprocess[_pytorch]
"""
from . import _cabi  # noqa: F401

__all__ = ["EventEmulator", "EventEmulatorBatch", "events_to_voxel_grid", "events_to_voxel_grid_pytorch",
           "event_preprocess", "event_preprocess_pytorch"]


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import v2e2v_b200` stays cheap
    if name in ("EventEmulator", "EventEmulatorBatch"):
        from . import emulator
        return getattr(emulator, name)
    if name in ("events_to_voxel_grid", "events_to_voxel_grid_pytorch", "event_preprocess",
                "event_preprocess_pytorch"):
        from . import event_process
        return getattr(event_process, name)
    raise AttributeError(name)

"""Per-object point-cloud preprocessing on the device.

Replaces the CPU loop of `MSR3DDataset.preprocess_pcd` (data/datasets/msr3d.py:181-215) and the padding of
`dataset_wrapper.py:151-158`: the raw scan points of the objects of a scene go to the GPU once, back to back,
and ONE kernel produces what the hot path consumes — `obj_fts (O_max, P, 6)` (sampled, centred, scaled to the
unit ball, padded with ones), `obj_locs (O_max, 6)` (centre + size, padded with zeros) and `obj_masks`.
The random draws stay where the reference makes them (host): pass the `np.random.choice` indices and the
`build_rotate_mat` matrix.  No CPU path: CPU tensors raise."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _C

c_int, c_vp = ctypes.c_int, ctypes.c_void_p
_C.register_protos({
    "msr3d_preprocess_objects_f32": [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
})


def pack_objects(obj_pcds: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """List of (N_i, 6) tensors -> (points (sum N_i, 6) f32, offsets (O+1,) int32), same device as the inputs."""
    sizes = [int(o.shape[0]) for o in obj_pcds]
    offsets = torch.tensor([0] + sizes, dtype=torch.int64).cumsum(0).to(torch.int32)
    points = torch.cat([o.reshape(-1, 6).float() for o in obj_pcds], 0) if obj_pcds else torch.zeros(0, 6)
    return points, offsets


def preprocess_objects(points: torch.Tensor, offsets: torch.Tensor, sample_idx: torch.Tensor,
                       rot: Optional[torch.Tensor], o_max: int):
    """points (total, 6) f32 CUDA; offsets (O+1,) int32; sample_idx (O, P) int32 indices inside each object;
    rot (3, 3) f32 or None.  Returns (obj_fts (o_max, P, 6) f32, obj_locs (o_max, 6) f32, obj_masks (o_max,) bool)."""
    if not points.is_cuda:
        raise RuntimeError("msr3d_b200.data: CUDA tensors only (no CPU path)")
    dev = points.device
    assert points.dtype == torch.float32 and points.dim() == 2 and points.shape[1] == 6 and points.is_contiguous()
    O, P = sample_idx.shape
    assert offsets.numel() == O + 1 and o_max >= O
    offsets = offsets.to(device=dev, dtype=torch.int32).contiguous()
    sample_idx = sample_idx.to(device=dev, dtype=torch.int32).contiguous()
    rot_d = rot.to(device=dev, dtype=torch.float32).contiguous() if rot is not None else None
    fts = torch.empty(o_max, P, 6, dtype=torch.float32, device=dev)
    locs = torch.empty(o_max, 6, dtype=torch.float32, device=dev)
    masks = torch.empty(o_max, dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        _C.call("msr3d_preprocess_objects_f32", points.data_ptr(), offsets.data_ptr(), sample_idx.data_ptr(),
                rot_d.data_ptr() if rot_d is not None else None, O, o_max, P, fts.data_ptr(), locs.data_ptr(),
                masks.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return fts, locs, masks

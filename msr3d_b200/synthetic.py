"""Seeded synthetic scene batches with the shapes and edge cases of the MSR3D data
pipeline.  Host-side (CPU tensors); used by bench.py and tests.

Mirrors what the reference's loaders produce, not their code:
  * per-object clouds are centred and scaled into the unit ball
    (data/datasets/msr3d.py:205-209), rgb in [-1, 1];
  * objects with fewer than P points are sampled WITH replacement
    (data/datasets/msr3d.py:200-201) -> duplicated points, FPS ties;
  * scenes are padded to O objects with all-ones points and obj_masks False
    (data/datasets/dataset_wrapper.py:155-158);
  * obj_locs = centre(3) + size(3); the agent pose is a position plus a z-yaw
    quaternion (data/data_utils.py:544-552).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

#: BASELINE.json `configs`, per-GPU shapes: name -> (B, O, P)
CONFIGS = {
    "cfg1_8x256": (1, 8, 256),
    "cfg2_60x1024": (4, 60, 1024),
    "cfg4_80x2048": (2, 80, 2048),
    "cfg5_msnn_60x1024": (8, 60, 1024),
}


@dataclass
class SceneBatch:
    obj_fts: torch.Tensor            # (B,O,P,6) f32
    obj_locs: torch.Tensor           # (B,O,6)  f32
    obj_masks: torch.Tensor          # (B,O)    bool, True = real object
    anchor_locs: torch.Tensor        # (B,3)    f32
    anchor_orientation: torch.Tensor  # (B,4)   f32 quaternion (x,y,z,w)

    def as_dict(self) -> dict:
        return {"obj_fts": self.obj_fts, "obj_locs": self.obj_locs, "obj_masks": self.obj_masks,
                "anchor_locs": self.anchor_locs, "anchor_orientation": self.anchor_orientation}

    def to(self, device, non_blocking: bool = False) -> "SceneBatch":
        return SceneBatch(*(t.to(device, non_blocking=non_blocking) for t in
                            (self.obj_fts, self.obj_locs, self.obj_masks, self.anchor_locs,
                             self.anchor_orientation)))

    def pin(self) -> "SceneBatch":
        return SceneBatch(*(t.pin_memory() for t in
                            (self.obj_fts, self.obj_locs, self.obj_masks, self.anchor_locs,
                             self.anchor_orientation)))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.obj_fts, self.obj_locs, self.obj_masks, self.anchor_locs,
                    self.anchor_orientation))


def make_batch(B: int, O: int, P: int, seed: int = 0, pad_frac: float = 0.1,
               dup_frac: float = 0.2) -> SceneBatch:
    g = torch.Generator().manual_seed(seed)
    xyz = torch.randn(B, O, P, 3, generator=g)
    # anisotropic boxes look more like furniture than spheres do
    xyz = xyz * (0.25 + 0.75 * torch.rand(B, O, 1, 3, generator=g))
    n_pad = int(round(O * pad_frac)) if O >= 4 else 0
    n_valid = O - n_pad
    n_dup = int(round(n_valid * dup_frac))
    for b in range(B):
        for o in range(n_dup):  # the first n_dup objects: 256 (or P/4) unique points, resampled
            uniq = max(4, min(256, P // 4))
            sel = torch.randint(0, uniq, (P,), generator=g)
            xyz[b, o] = xyz[b, o, :uniq][sel]
    xyz = xyz - xyz.mean(dim=2, keepdim=True)
    xyz = xyz / xyz.norm(dim=3).amax(dim=2).clamp_min(1e-6)[..., None, None]
    rgb = torch.rand(B, O, P, 3, generator=g) * 2 - 1
    fts = torch.cat([xyz, rgb], dim=3).contiguous()
    masks = torch.ones(B, O, dtype=torch.bool)
    centers = torch.rand(B, O, 3, generator=g) * torch.tensor([8.0, 8.0, 2.5]) - torch.tensor([4.0, 4.0, 0.0])
    sizes = torch.rand(B, O, 3, generator=g) * 1.9 + 0.1
    locs = torch.cat([centers, sizes], dim=2)
    if n_pad:
        fts[:, n_valid:] = 1.0     # dataset_wrapper.py:156 pads with ones
        locs[:, n_valid:] = 0.0
        masks[:, n_valid:] = False
    anchor = torch.rand(B, 3, generator=g) * torch.tensor([8.0, 8.0, 2.5]) - torch.tensor([4.0, 4.0, 0.0])
    theta = torch.rand(B, generator=g) * 2 * math.pi
    quat = torch.stack([torch.zeros(B), torch.zeros(B), torch.sin(theta / 2), torch.cos(theta / 2)], 1)
    return SceneBatch(fts, locs.contiguous(), masks, anchor, quat)


def make_config_batch(name: str, seed: int = 0) -> SceneBatch:
    B, O, P = CONFIGS[name]
    return make_batch(B, O, P, seed=seed)


# ---------------------------------------------------------------------------
# Deterministic weights keyed by the reference's state-dict names
# ---------------------------------------------------------------------------
def _key_seed(name: str, seed: int) -> int:
    h = 1469598103934665603
    for ch in name.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ (seed * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF


def make_state_dict(spec, seed: int = 0) -> dict:
    """Random but architecture-plausible tensors for every (name, shape, dtype) of
    ``spec`` (tests/golden/state_dict_spec.json lists the reference's OSE3DSituation
    keys).  Independent of module construction order, so the reference model (in the
    authoring container), the oracle and the CUDA modules all see identical weights."""
    sd = {}
    for name, shape, dtype in spec:
        g = torch.Generator().manual_seed(_key_seed(name, seed))
        shape = tuple(shape)
        leaf = name.rsplit(".", 1)[-1]
        if dtype == "int64":  # BatchNorm num_batches_tracked
            t = torch.zeros(shape, dtype=torch.int64)
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "weight" and len(shape) == 1:      # LayerNorm / BatchNorm scale
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif name == "anchor_size":
            t = torch.ones(shape)
        else:                                            # free parameters (anchor_feat, ...)
            t = 0.05 * torch.randn(shape, generator=g)
        sd[name] = t
    return sd

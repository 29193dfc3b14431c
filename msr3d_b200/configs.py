"""Shipped configuration of the prompter (configs/msr3d.yaml:172-217 of the reference) as plain attribute
dicts — what hydra hands `OSE3DSituation` — and the state-dict spec (names, shapes, dtypes) of the shipped
model, so that benches, tools and `smoke()` can build the flagship configuration without the test tree."""
from __future__ import annotations

import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(o):
    if isinstance(o, dict):
        return AttrDict({k: to_attr(v) for k, v in o.items()})
    if isinstance(o, (list, tuple)):
        return [to_attr(v) for v in o]
    return o


def prompter_cfg(situation_type="as_transform_for_objects", fusion="cond", freeze=True, **spatial_over):
    """configs/msr3d.yaml:174-217 as a plain attribute dict (what hydra hands the reference)."""
    se = {"dim_loc": 6, "num_attention_heads": 8, "dim_feedforward": 2048, "dropout": 0.1,
          "activation": "gelu", "spatial_multihead": True, "spatial_dim": 5, "spatial_dist_norm": True,
          "spatial_attn_fusion": fusion, "num_layers": 3, "obj_loc_encoding": "same_all",
          "pairwise_rel_type": "center"}
    se.update(spatial_over)
    return to_attr({"model": {
        "name": "OSE3DSituation", "situation_type": situation_type, "scene_token_len": 60,
        "loc_fourier_dim": 63, "hidden_size": 256, "label_size": 300, "vision_backbone_name": "gtpcd",
        "use_spatial_attn": True, "use_anchor": True, "use_orientation": True, "fourier_size": 84,
        "attn_flat": {"use_attn_flat": False, "mcan_flat_mlp_size": 512, "mcan_flat_glimpses": 1,
                      "mcan_flat_out_size": 1024},
        "vision": {"name": "PcdObjEncoder", "args": {
            "sa_n_points": [32, 16, None], "sa_n_samples": [32, 32, None], "sa_radii": [0.2, 0.4, None],
            "sa_mlps": [[3, 64, 64, 128], [128, 128, 128, 256], [256, 256, 512, 768]],
            "dropout": 0.1, "freeze": bool(freeze), "path": None}},
        "spatial_encoder": se}})


def load_spec():
    """[(name, shape, dtype)] of the shipped model's state dict (6 018 290 parameters)."""
    with open(os.path.join(_HERE, "specs", "state_dict_spec.json")) as fh:
        return [tuple(x) for x in json.load(fh)]


def build_prompter(seed: int, situation_type="as_transform_for_objects", device="cpu", fusion="cond", freeze=True,
                   **spatial_over):
    """Drop-in OSE3DSituation of the shipped configuration with seeded synthetic weights (eval mode).
    Other fusions / situation modes own a different set of keys; values are keyed by NAME
    (synthetic.make_state_dict), so every key shared with the shipped model gets the same tensor."""
    import msr3d_b200
    model = msr3d_b200.build_model(prompter_cfg(situation_type, fusion, freeze, **spatial_over))
    model.load_state_dict(model_state_dict(model, seed), strict=True)
    return model.to(device).eval()


def model_state_dict(model, seed: int) -> dict:
    """Seeded synthetic tensors for exactly the keys of `model` (what tests/golden/gen_golden.py loads into the
    reference model of the same configuration)."""
    from . import synthetic
    spec = [(k, list(v.shape), "int64" if str(v.dtype) == "torch.int64" else "float32")
            for k, v in model.state_dict().items()]
    return synthetic.make_state_dict(spec, seed)

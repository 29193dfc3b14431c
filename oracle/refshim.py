"""ORACLE — TEST INFRASTRUCTURE ONLY.  Authoring-container only: needs /root/reference.

Imports the *unmodified* MSR3D reference hot-path modules (PcdObjEncoder,
PointNetPP, TransformerSpatialEncoderLayer, OSE3DSituation, ...) on CPU so that
golden vectors can be produced from the reference's own Python
(tests/golden/gen_golden.py).  Nothing here copies reference source: it only
arranges ``sys.modules`` so that the reference files import where they lie.

What is stubbed and why:
  * packages ``model``, ``modules``, ``modules.vision``, ``optim`` are
    registered with ``__path__`` only, skipping their heavy ``__init__`` files
    (clip / hydra / peft / timm / data stack are not installed here);
  * ``fvcore.common.registry.Registry`` -> a dict-backed register()/get();
  * ``easydict``, ``omegaconf.OmegaConf.to_container``, ``timm.models.*``;
  * three names that left ``transformers.modeling_utils`` in transformers 5.x
    (modules/layers/transformers.py:14-16 imports them but the hot path never
    calls them);
  * ``pointnet2._ext`` -> oracle/ext_oracle.py (the reference extension is
    CUDA-only: ball_query.cpp:27-29 asserts "CPU not supported").
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MSR3D_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "modules", "third_party", "pointnet2"))


class _Registry:
    """Minimal stand-in for fvcore.common.registry.Registry."""

    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(cls):
                assert cls.__name__ not in self._obj_map, f"{cls.__name__} already registered in {self._name}"
                self._obj_map[cls.__name__] = cls
                return cls
            return deco
        self._obj_map[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._obj_map[name]


class AttrDict(dict):
    """Attribute-style config (what hydra/omegaconf gives the reference)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_attr(v) for v in obj]
    return obj


def _to_container(cfg):
    if isinstance(cfg, dict):
        return {k: _to_container(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return [_to_container(v) for v in cfg]
    return cfg


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_installed = False


def install(ext_module=None):
    """Arrange sys.modules/sys.path; idempotent.  Returns nothing."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    for pkg, rel in (("model", "model"), ("modules", "modules"),
                     ("modules.vision", "modules/vision"), ("optim", "optim")):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
        sys.modules[pkg] = m

    # transformers probes `timm` with find_spec at import time, so import it
    # (and the submodules the reference touches) before any stub goes in.
    import transformers.activations  # noqa: F401
    import transformers.modeling_outputs  # noqa: F401
    import transformers.modeling_utils as mu
    if not hasattr(mu, "Conv1D"):
        from transformers.pytorch_utils import Conv1D
        mu.Conv1D = Conv1D
    for missing in ("find_pruneable_heads_and_indices", "prune_conv1d_layer"):
        if not hasattr(mu, missing):
            setattr(mu, missing, None)

    if "fvcore" not in sys.modules:
        _stub_module("fvcore")
        _stub_module("fvcore.common")
        _stub_module("fvcore.common.registry", Registry=_Registry)
    if "easydict" not in sys.modules:
        _stub_module("easydict", EasyDict=AttrDict)
    if "omegaconf" not in sys.modules:
        oc = type("OmegaConf", (), {"to_container": staticmethod(_to_container)})
        _stub_module("omegaconf", OmegaConf=oc)
    if "timm" not in sys.modules:
        _stub_module("timm")
        _stub_module("timm.models")
        _stub_module("timm.models.vision_transformer", PatchEmbed=object, Attention=object, Mlp=object)
        _stub_module("timm.models.layers", DropPath=object, trunc_normal_=None)

    if ext_module is None:
        from oracle import ext_oracle as ext_module
    pn2 = types.ModuleType("pointnet2")
    pn2.__path__ = []
    pn2._ext = ext_module
    sys.modules["pointnet2"] = pn2
    sys.modules["pointnet2._ext"] = ext_module
    _installed = True


def load():
    """Import the reference hot-path modules; returns a namespace of them."""
    install()
    ns = types.SimpleNamespace()
    ns.pointnet2_utils = importlib.import_module("modules.third_party.pointnet2.pointnet2_utils")
    ns.pointnet2_modules = importlib.import_module("modules.third_party.pointnet2.pointnet2_modules")
    ns.pointnet = importlib.import_module("modules.layers.pointnet")
    ns.transformers = importlib.import_module("modules.layers.transformers")
    ns.utils = importlib.import_module("modules.utils")
    ns.pcd_pointnet_encoder = importlib.import_module("modules.vision.pcd_pointnet_encoder")
    ns.ose3d_situation = importlib.import_module("model.ose3d_situation")
    ns.build = importlib.import_module("modules.build")
    return ns


def msr3d_prompter_cfg(**over):
    """The `model.prompter` section of configs/msr3d.yaml:174-217 as an attr-dict."""
    cfg = {
        "model": {
            "name": "OSE3DSituation",
            "situation_type": "as_transform_for_objects",
            "scene_token_len": 60,
            "loc_fourier_dim": 63,
            "hidden_size": 256,
            "label_size": 300,
            "vision_backbone_name": "gtpcd",
            "use_spatial_attn": True,
            "use_anchor": True,
            "use_orientation": True,
            "fourier_size": 84,
            "attn_flat": {"use_attn_flat": False, "mcan_flat_mlp_size": 512,
                          "mcan_flat_glimpses": 1, "mcan_flat_out_size": 1024},
            "vision": {
                "name": "PcdObjEncoder",
                "args": {
                    "sa_n_points": [32, 16, None],
                    "sa_n_samples": [32, 32, None],
                    "sa_radii": [0.2, 0.4, None],
                    "sa_mlps": [[3, 64, 64, 128], [128, 128, 128, 256], [256, 256, 512, 768]],
                    "dropout": 0.1,
                    "freeze": True,
                    "path": None,
                },
            },
            "spatial_encoder": {
                "dim_loc": 6, "num_attention_heads": 8, "dim_feedforward": 2048,
                "dropout": 0.1, "activation": "gelu", "spatial_multihead": True,
                "spatial_dim": 5, "spatial_dist_norm": True,
                "spatial_attn_fusion": "cond", "num_layers": 3,
                "obj_loc_encoding": "same_all", "pairwise_rel_type": "center",
            },
        }
    }
    cfg["model"].update(over)
    return to_attr(cfg)

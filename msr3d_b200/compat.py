"""Makes this package answer to the reference's import paths, so MSR3D code that does
`import pointnet2._ext`, `from modules.vision.pcd_pointnet_encoder import PcdObjEncoder`,
`from model.ose3d_situation import OSE3DSituation` or goes through `modules.build` /
`model.build` registries gets the B200-native implementation.

`install_ext()` alone is the minimal drop-in: the reference's own Python
(pointnet2_utils.py:22-30) then runs on these kernels."""
from __future__ import annotations

import sys
import types


def install_ext() -> None:
    """Publish msr3d_b200.ext as `pointnet2._ext` (the name pointnet2_utils.py hard-codes)."""
    from . import ext
    pkg = sys.modules.get("pointnet2")
    if pkg is None:
        pkg = types.ModuleType("pointnet2")
        pkg.__path__ = []
        sys.modules["pointnet2"] = pkg
    pkg._ext = ext
    sys.modules["pointnet2._ext"] = ext


def register_plugins() -> None:
    """Register PcdObjEncoder / OSE3DSituation in this package's registries (idempotent)."""
    from . import encoder, prompter, registry  # noqa: F401  (prompter self-registers)
    if "PcdObjEncoder" not in getattr(registry.VISION_REGISTRY, "_obj_map", {}):
        registry.VISION_REGISTRY.register(encoder.PcdObjEncoder)


def install_modules() -> None:
    """Alias the reference module paths of the hot path to this package's modules."""
    from . import encoder, prompter, registry, spatial
    register_plugins()

    def alias(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    alias("modules.build", VISION_REGISTRY=registry.VISION_REGISTRY, LANGUAGE_REGISTRY=registry.LANGUAGE_REGISTRY,
          GROUNDING_REGISTRY=registry.GROUNDING_REGISTRY, HEADS_REGISTRY=registry.HEADS_REGISTRY,
          build_module=registry.build_module)
    alias("model.build", MODEL_REGISTRY=registry.MODEL_REGISTRY, BaseModel=registry.BaseModel,
          build_model=registry.build_model)
    alias("modules.vision.pcd_pointnet_encoder", PcdObjEncoder=encoder.PcdObjEncoder)
    alias("modules.layers.pointnet", PointNetPP=encoder.PointNetPP)
    alias("modules.layers.transformers", **{k: getattr(spatial, k) for k in (
        "MultiHeadAttentionSpatial", "TransformerSpatialEncoderLayer", "TransformerEncoderLayer",
        "CrossAttentionLayer")})
    alias("model.ose3d_situation", OSE3DSituation=prompter.OSE3DSituation,
          generate_fourier_features=spatial.generate_fourier_features)
    install_ext()

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


def _patch(name: str, **attrs):
    """Put `attrs` on module `name`.  If the reference module itself can be imported (the reference tree is on
    sys.path) it is imported and ONLY these attributes are replaced — everything else it defines (GPT2Model,
    TransformerDecoderLayer, ... which model/msr3d/msr3d.py:27 imports from modules.layers.transformers) stays the
    reference's.  Otherwise a stand-in module is registered whose missing attributes fail with an explanation instead of
    a bare ImportError."""
    import importlib
    mod = sys.modules.get(name)
    if mod is None:
        try:
            mod = importlib.import_module(name)
        except Exception:  # noqa: BLE001  (reference tree absent, or one of its heavy dependencies is)
            mod = None
    if mod is None:
        parts = name.split(".")
        for i in range(1, len(parts)):                 # parent packages
            pkg = ".".join(parts[:i])
            if pkg not in sys.modules:
                pm = types.ModuleType(pkg)
                pm.__path__ = []
                sys.modules[pkg] = pm
        mod = types.ModuleType(name)
        mod.__msr3d_stub__ = True

        def _missing(attr, _name=name):
            raise AttributeError(
                f"{_name}.{attr}: this is msr3d_b200's stand-in for the reference module (it only carries the hot-path "
                f"classes).  Put the MSR3D tree on sys.path before msr3d_b200.compat.install_modules() so that the real "
                f"module is imported and patched instead.")
        mod.__getattr__ = _missing
        sys.modules[name] = mod
        parent = sys.modules.get(name.rpartition(".")[0])
        if parent is not None:
            setattr(parent, name.rpartition(".")[2], mod)
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def install_modules() -> None:
    """Replace the hot-path classes inside the reference's modules (or stand-ins for them when the reference tree is not
    importable) by this package's implementations."""
    from . import encoder, prompter, registry, spatial
    register_plugins()
    _patch("modules.build", VISION_REGISTRY=registry.VISION_REGISTRY, LANGUAGE_REGISTRY=registry.LANGUAGE_REGISTRY,
           GROUNDING_REGISTRY=registry.GROUNDING_REGISTRY, HEADS_REGISTRY=registry.HEADS_REGISTRY,
           build_module=registry.build_module)
    _patch("model.build", MODEL_REGISTRY=registry.MODEL_REGISTRY, BaseModel=registry.BaseModel,
           build_model=registry.build_model)
    _patch("modules.vision.pcd_pointnet_encoder", PcdObjEncoder=encoder.PcdObjEncoder)
    _patch("modules.layers.pointnet", PointNetPP=encoder.PointNetPP)
    _patch("modules.layers.transformers", **{k: getattr(spatial, k) for k in (
        "MultiHeadAttentionSpatial", "TransformerSpatialEncoderLayer", "TransformerEncoderLayer",
        "CrossAttentionLayer")})
    _patch("model.ose3d_situation", OSE3DSituation=prompter.OSE3DSituation,
           generate_fourier_features=spatial.generate_fourier_features)
    install_ext()

"""Plugin registries with the reference's API (modules/build.py:6-21, model/build.py:6-19).

fvcore's Registry raises on duplicate names, so a replacement registers *instead of*
the reference class.  When fvcore is importable its Registry is used; otherwise a
behaviour-compatible minimal one."""
from __future__ import annotations

import torch.nn as nn

try:  # pragma: no cover - fvcore is not in the authoring image
    from fvcore.common.registry import Registry
except Exception:  # noqa: BLE001
    class Registry:
        def __init__(self, name: str):
            self._name = name
            self._obj_map = {}

        def _do_register(self, name, obj):
            assert name not in self._obj_map, \
                f"An object named '{name}' was already registered in '{self._name}' registry!"
            self._obj_map[name] = obj

        def register(self, obj=None):
            if obj is None:
                def deco(func_or_class):
                    self._do_register(func_or_class.__name__, func_or_class)
                    return func_or_class
                return deco
            self._do_register(obj.__name__, obj)
            return obj

        def get(self, name):
            ret = self._obj_map.get(name)
            if ret is None:
                raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
            return ret

        def __contains__(self, name):
            return name in self._obj_map


VISION_REGISTRY = Registry("vision")
LANGUAGE_REGISTRY = Registry("language")
GROUNDING_REGISTRY = Registry("grounding")
HEADS_REGISTRY = Registry("heads")
MODEL_REGISTRY = Registry("model")


def cfg2dict(cfg):
    """common/type_utils.py:9-10 (OmegaConf.to_container) for omegaconf nodes and plain mappings."""
    try:
        from omegaconf import OmegaConf, DictConfig, ListConfig
        if isinstance(cfg, (DictConfig, ListConfig)):
            return OmegaConf.to_container(cfg)
    except ImportError:
        pass
    if isinstance(cfg, dict):
        return {k: cfg2dict(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return [cfg2dict(v) for v in cfg]
    return cfg


_REGS = {"vision": VISION_REGISTRY, "language": LANGUAGE_REGISTRY, "grounding": GROUNDING_REGISTRY,
         "heads": HEADS_REGISTRY}


def build_module(module_type, cfg):
    if module_type not in _REGS:
        raise NotImplementedError(f"module type {module_type} not implemented")
    return _REGS[module_type].get(cfg.name)(cfg, **cfg2dict(cfg.args))


class BaseModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()

    def get_opt_params(self):
        raise NotImplementedError("Function to obtain all default parameters for optimization")


def build_model(cfg):
    return MODEL_REGISTRY.get(cfg.model.name)(cfg)

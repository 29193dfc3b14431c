import json
import os

import numpy as np
import torch

from msr3d_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


from msr3d_b200.configs import AttrDict, to_attr, prompter_cfg, build_prompter  # noqa: E402,F401  (moved into the package)


def load_spec():
    with open(os.path.join(GOLDEN, "state_dict_spec.json")) as fh:
        return [tuple(x) for x in json.load(fh)]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

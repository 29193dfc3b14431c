"""Generates tests/golden/preprocess.npz from the UNMODIFIED reference `MSR3DDataset.preprocess_pcd`
(data/datasets/msr3d.py:181-215: rotate, centre / size, random subsample, centre + scale to the unit ball)
and `build_rotate_mat` (data/data_utils.py:175-189), plus the padding of `dataset_wrapper.py:151-158`.

Authoring-container only (needs /root/reference):   python tests/golden/gen_golden_preprocess.py
The dataset module cannot be imported here (jsonlines, nltk, accelerate, open3d ... are not installed), so the
two FUNCTION DEFINITIONS are located in the reference files with `ast`, compiled from the reference source
where it lies and executed unchanged in a namespace that holds only what they touch (numpy, torch, random,
scipy Rotation).  `np.random.choice` / `random.choice` are recorded so that the device op, which takes the
sample indices and the rotation as inputs, can be fed the same draws.
"""
import ast
import os
import random
import sys
import types

import numpy as np
import torch
from scipy.spatial.transform import Rotation as R

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("MSR3D_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def extract(path, name):
    """Compile the (unmodified) source of function / method `name` found in `path`."""
    src = open(path).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise KeyError(name)


def load():
    ns = {"np": np, "torch": torch, "random": random, "R": R}
    exec(compile(extract(os.path.join(REF, "data", "data_utils.py"), "build_rotate_mat"), "data_utils.py", "exec"), ns)
    import textwrap
    exec(compile(textwrap.dedent(extract(os.path.join(REF, "data", "datasets", "msr3d.py"), "preprocess_pcd")),
                 "msr3d.py", "exec"), ns)
    return ns


def case(ns, seed, sizes, P, split):
    rng = np.random.RandomState(seed)
    objs = []
    for n in sizes:
        xyz = (rng.randn(n, 3) * rng.uniform(0.1, 1.5, size=3) + rng.uniform(-3, 3, size=3)).astype(np.float32)
        if n == 7:
            xyz[:] = xyz[0]                                   # a degenerate object: max_dist < 1e-6 -> 1
        rgb = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
        objs.append(np.concatenate([xyz, rgb], 1))
    raw = [o.copy() for o in objs]
    draws = []
    real_choice = np.random.choice

    def rec_choice(a, size=None, replace=True):
        out = real_choice(a, size=size, replace=replace)
        draws.append(np.asarray(out))
        return out
    np.random.seed(seed)
    random.seed(seed)
    np.random.choice = rec_choice
    try:
        self = types.SimpleNamespace(split=split, num_points=P)
        out = ns["preprocess_pcd"](self, objs, return_anchor=False, rot_aug=True)
    finally:
        np.random.choice = real_choice
    # the rotation the reference drew: recover it from the first object (the function rotates in place)
    random.seed(seed)
    rot = ns["build_rotate_mat"](split, rot_aug=True)
    O_max = len(sizes) + 2
    fts = out["obj_fts"]
    pad = torch.ones((O_max - fts.shape[0],) + tuple(fts.shape[1:]), dtype=fts.dtype)
    return {
        "points": np.concatenate(raw, 0), "offsets": np.cumsum([0] + list(sizes)).astype(np.int32),
        "sample_idx": np.stack(draws, 0).astype(np.int32),
        "rot": (rot if rot is not None else np.eye(3, dtype=np.float32)).astype(np.float32),
        "has_rot": np.array(rot is not None),
        "obj_fts": torch.cat([fts, pad], 0).float().numpy(),                        # dataset_wrapper.py:155
        "obj_locs": torch.cat([out["obj_locs"], torch.zeros(O_max - fts.shape[0], 6, dtype=out["obj_locs"].dtype)], 0).float().numpy(),
        "obj_masks": (np.arange(O_max) < len(sizes)),
    }


def main():
    ns = load()
    out = {}
    for name, kw in {
        "train": dict(seed=3, sizes=[500, 64, 7, 1300, 32], P=128, split="train"),   # rotation drawn, up- and down-sampling
        "val": dict(seed=4, sizes=[300, 2000, 50], P=256, split="val"),              # no rotation
    }.items():
        for k, v in case(ns, **kw).items():
            out[f"{name}:{k}"] = v
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()}, bool(out["train:has_rot"]))


if __name__ == "__main__":
    main()

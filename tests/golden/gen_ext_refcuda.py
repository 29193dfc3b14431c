"""Runs the reference's OWN CUDA kernels (oracle/_ref, compiled from the reference .cu
files for sm_100) on a B200 and stores their outputs as golden vectors:
    gpurun -- python tests/golden/gen_ext_refcuda.py      # writes gpurun_out/ext_refcuda.npz
then copy gpurun_out/ext_refcuda.npz to tests/golden/.  Inputs are regenerated from
seeds by tests/golden/ext_cases.py, so the fixture holds outputs only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refcuda  # noqa: E402
from tests.golden import ext_cases  # noqa: E402


def main():
    ref = refcuda.load()
    ref_nofmad = refcuda.load("pointnet2_ref_ext_nofmad")
    assert ref is not None, "oracle/_ref/pointnet2_ref_ext.so missing (python oracle/build_ref.py)"
    out = {}
    for name, x, m in ext_cases.fps_cases():
        out[f"fps/{name}"] = ref.furthest_point_sampling(x.cuda(), m).cpu().numpy()
        if ref_nofmad is not None:
            out[f"fps_nofmad/{name}"] = ref_nofmad.furthest_point_sampling(x.cuda(), m).cpu().numpy()
    for name, c, x, r, ns in ext_cases.ball_cases():
        out[f"ball/{name}"] = ref.ball_query(c.cuda(), x.cuda(), r, ns).cpu().numpy()
        if ref_nofmad is not None:
            out[f"ball_nofmad/{name}"] = ref_nofmad.ball_query(c.cuda(), x.cuda(), r, ns).cpu().numpy()
    for name, u, k in ext_cases.nn_cases():
        d, i = ref.three_nn(u.cuda(), k.cuda())
        out[f"nn_idx/{name}"] = i.cpu().numpy()
        out[f"nn_dist/{name}"] = d.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "ext_refcuda.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", torch.cuda.get_device_name(0))


if __name__ == "__main__":
    main()

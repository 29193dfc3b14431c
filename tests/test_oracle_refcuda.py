"""CPU: the C oracle against golden vectors produced on a B200 by the reference's OWN CUDA
kernels (reference .cu files compiled for sm_100 into oracle/_ref; generator:
tests/golden/gen_ext_refcuda.py).  This is what pins the oracle's int32 semantics —
FPS tie-break order, FMA contraction, strict comparisons — to the reference binary."""
import os

import numpy as np
import pytest

from oracle import ext_oracle as eo
from tests.golden import ext_cases


@pytest.fixture(scope="module")
def fx(golden_dir):
    return np.load(os.path.join(golden_dir, "ext_refcuda.npz"))


def test_fps_matches_reference_cuda(fx):
    eo.set_threads(4)
    for name, x, m in ext_cases.fps_cases():
        np.testing.assert_array_equal(eo.furthest_point_sampling(x, m).numpy(), fx[f"fps/{name}"], err_msg=name)


def test_ball_query_matches_reference_cuda(fx):
    for name, c, x, r, ns in ext_cases.ball_cases():
        np.testing.assert_array_equal(eo.ball_query(c, x, r, ns).numpy(), fx[f"ball/{name}"], err_msg=name)


def test_three_nn_matches_reference_cuda(fx):
    for name, u, k in ext_cases.nn_cases():
        d, i = eo.three_nn(u, k)
        np.testing.assert_array_equal(i.numpy(), fx[f"nn_idx/{name}"])
        np.testing.assert_array_equal(d.numpy(), fx[f"nn_dist/{name}"])


def test_fmad_false_build_is_reproduced_by_the_uncontracted_oracle(fx):
    """-fmad=false reference build == oracle with fmad=0: contraction is the only difference."""
    old = eo.FMAD
    eo.FMAD = 0
    try:
        n_diff = 0
        for name, x, m in ext_cases.fps_cases():
            got = eo.furthest_point_sampling(x, m).numpy()
            np.testing.assert_array_equal(got, fx[f"fps_nofmad/{name}"], err_msg=name)
            n_diff += int((got != fx[f"fps/{name}"]).sum())
        for name, c, x, r, ns in ext_cases.ball_cases():
            got = eo.ball_query(c, x, r, ns).numpy()
            np.testing.assert_array_equal(got, fx[f"ball_nofmad/{name}"], err_msg=name)
            n_diff += int((got != fx[f"ball/{name}"]).sum())
    finally:
        eo.FMAD = old
    print("index entries that differ between -fmad=true and -fmad=false reference builds:", n_diff)

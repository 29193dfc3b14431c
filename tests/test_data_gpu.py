"""GPU: device-side object preprocessing (msr3d_b200/data.py) against the fixture produced by the
unmodified reference `preprocess_pcd` (tests/golden/preprocess.npz) and against the oracle at scene size."""
import numpy as np
import pytest
import torch

from oracle import data_oracle
from tests import helpers

pytestmark = pytest.mark.gpu


def _run(points, offsets, sample_idx, rot, o_max):
    from msr3d_b200 import data
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    fts, locs, masks = data.preprocess_objects(t(points).cuda(), t(offsets), t(sample_idx),
                                               t(rot) if rot is not None else None, o_max)
    return fts.cpu().numpy(), locs.cpu().numpy(), masks.cpu().numpy()


def _check(got, want):
    fts, locs, masks = got
    wf, wl, wm = want
    np.testing.assert_array_equal(masks, wm)
    np.testing.assert_array_equal(fts[..., 3:], wf[..., 3:])                       # colours are copied: exact
    np.testing.assert_array_equal(fts[~wm], wf[~wm])                               # padding: exact
    np.testing.assert_allclose(locs, wl, rtol=1e-5, atol=2e-6)
    # An object whose points all coincide (size 0) is ill-conditioned in the reference itself: the sampled points
    # minus their float32 mean are pure rounding noise (~1e-7), which is either kept (max_dist < 1e-6 -> 1) or
    # blown up to O(1) (max_dist just above 1e-6), depending on the summation order.  There only finiteness and the
    # unit-ball bound are comparable; everywhere else coordinates agree to float32 rounding.
    degenerate = wm & (np.abs(wl[:, 3:]).max(-1) < 1e-6)
    ok = ~degenerate
    np.testing.assert_allclose(fts[ok][..., :3], wf[ok][..., :3], rtol=1e-5, atol=2e-6)
    assert np.isfinite(fts).all() and np.abs(fts[degenerate][..., :3]).max(initial=0.0) <= 1.0 + 1e-5


@pytest.mark.parametrize("case", ["train", "val"])
def test_preprocess_matches_reference_fixture(case):
    fx = helpers.golden("preprocess")
    g = lambda k: fx[f"{case}:{k}"]  # noqa: E731
    rot = g("rot") if bool(g("has_rot")) else None
    got = _run(g("points"), g("offsets"), g("sample_idx"), rot, g("obj_fts").shape[0])
    _check(got, (g("obj_fts"), g("obj_locs"), g("obj_masks")))
    fts = got[0]
    valid = g("obj_masks")
    norms = np.sqrt((fts[valid][..., :3] ** 2).sum(-1)).max(-1)
    assert np.all((np.abs(norms - 1.0) < 1e-5) | (norms < 1e-5))                   # unit ball (or the degenerate object)


def test_preprocess_scene_size_matches_oracle():
    """A scene like config 2: 54 objects of 30 .. 20000 points -> 60 x 1024, rotation by pi/2."""
    rng = np.random.RandomState(0)
    sizes = list(rng.randint(30, 20000, size=54))
    pts = [np.concatenate([(rng.randn(n, 3) * rng.uniform(0.05, 2.0, 3) + rng.uniform(-5, 5, 3)),
                           rng.uniform(-1, 1, (n, 3))], 1).astype(np.float32) for n in sizes]
    points = np.concatenate(pts, 0)
    offsets = np.cumsum([0] + sizes).astype(np.int32)
    idx = np.stack([rng.choice(n, size=1024, replace=n < 1024) for n in sizes]).astype(np.int32)
    th = np.pi / 2
    rot = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], dtype=np.float32)
    got = _run(points, offsets, idx, rot, 60)
    _check(got, data_oracle.preprocess_objects(points, offsets, idx, rot, 60))
    # deterministic
    again = _run(points, offsets, idx, rot, 60)
    assert all(np.array_equal(a, b) for a, b in zip(got, again))


def test_preprocess_empty_scene_is_all_padding():
    from msr3d_b200 import data
    fts, locs, masks = data.preprocess_objects(torch.zeros(0, 6, device="cuda"), torch.zeros(1, dtype=torch.int32),
                                               torch.zeros(0, 16, dtype=torch.int32), None, 3)
    assert fts.shape == (3, 16, 6) and bool((fts == 1).all()) and bool((locs == 0).all()) and not bool(masks.any())

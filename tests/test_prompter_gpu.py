"""GPU: the whole hot path (PointNet++ on tensor cores + spatial transformer) against the
reference fixtures and the CPU oracle.  fp tolerance: 1e-3 relative (north star)."""
import numpy as np
import pytest
import torch

from msr3d_b200 import synthetic
from tests import helpers

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.mark.parametrize("name,stype", [("prompter_cfg1", "as_transform_for_objects"),
                                        ("prompter_mid", "as_transform_for_objects"),
                                        ("prompter_as_object", "as_object")])
def test_prompter_matches_reference_fixture(name, stype):
    fx = helpers.golden(name)
    seed = int(fx["seed"])
    model = helpers.build_prompter(seed, stype, "cuda")
    batch = synthetic.make_batch(int(fx["B"]), int(fx["O"]), int(fx["P"]), seed=seed).to("cuda")
    tr = {}
    flat = batch.obj_fts.reshape(-1, batch.obj_fts.shape[2], 6)
    emb = model.obj_encoder.pcd_net.forward_fused(flat, trace=tr)
    np.testing.assert_array_equal(tr["fps1"].cpu().numpy(), fx["fps_idx_0"])
    np.testing.assert_array_equal(tr["ball1"].cpu().numpy(), fx["ball_idx_0"])
    np.testing.assert_array_equal(tr["fps2"].cpu().numpy(), fx["fps_idx_1"])
    np.testing.assert_array_equal(tr["ball2"].cpu().numpy(), fx["ball_idx_1"])
    e = helpers.rel_err(emb.cpu().view(fx["obj_embeds"].shape), torch.from_numpy(fx["obj_embeds"]))
    with torch.no_grad():
        out = model(batch.as_dict())
    t = helpers.rel_err(out["obj_tokens"].cpu(), torch.from_numpy(fx["obj_tokens"]))
    print(f"{name}: encoder rel err {e:.2e}, obj_tokens rel err {t:.2e}")
    assert e < TOL and t < TOL
    np.testing.assert_array_equal(out["obj_masks"].cpu().numpy(), fx["obj_masks_out"])


def test_encoder_full_size_matches_oracle(sd_spec):
    """config-2 shapes (240 clouds x 1024 points) against the CPU oracle."""
    from oracle import ext_oracle as eo, modules_oracle as mo
    sd = synthetic.make_state_dict(sd_spec, 21)
    batch = synthetic.make_batch(4, 60, 1024, seed=21)
    eo.set_threads(16)
    torch.set_num_threads(16)
    with torch.no_grad():
        want, _ = mo.pcd_obj_encoder(sd, "obj_encoder", batch.obj_fts)
    model = helpers.build_prompter(21, device="cuda")
    got = model.obj_encoder.encode(batch.obj_fts.cuda())
    e = helpers.rel_err(got.cpu(), want)
    print("encoder 240x1024 rel err", e)
    assert e < TOL
    # padded objects are all-ones clouds: identical embeddings, still finite
    assert torch.isfinite(got).all()


def test_layerwise_encoder_path_matches_fused():
    model = helpers.build_prompter(5, device="cuda")
    batch = synthetic.make_batch(1, 6, 512, seed=5).to("cuda")
    flat = batch.obj_fts.reshape(-1, 512, 6)
    with torch.no_grad():
        a = model.obj_encoder.pcd_net.forward_fused(flat)
        b = model.obj_encoder.pcd_net.forward_layerwise(flat)
    assert helpers.rel_err(a, b) < TOL


def test_encoder_linearity_free_properties_at_full_size():
    """Size-independent checks at BASELINE config 2: determinism, permutation of clouds, padding."""
    model = helpers.build_prompter(9, device="cuda")
    batch = synthetic.make_batch(4, 60, 1024, seed=9).to("cuda")
    e1 = model.obj_encoder.encode(batch.obj_fts)
    e2 = model.obj_encoder.encode(batch.obj_fts)
    assert torch.equal(e1, e2)                                     # run-to-run bit-identical
    perm = torch.randperm(60, device="cuda")
    e3 = model.obj_encoder.encode(batch.obj_fts[:, perm])
    assert torch.equal(e3, e1[:, perm])                            # clouds are independent
    pad = e1[:, -6:]                                               # 10% padded all-ones objects
    assert torch.equal(pad, pad[:, :1].expand_as(pad))

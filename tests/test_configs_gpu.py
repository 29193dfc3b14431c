"""GPU: the other BASELINE.json configurations as parity cases (not bench lines):
config 4 shapes (80 objects x 2048 points, bs 2) and config 5 shapes (msnn, bs 8, 60 x 1024)."""
import numpy as np
import pytest
import torch

from msr3d_b200 import synthetic
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cfg4_80x2048", "cfg5_msnn_60x1024"])
def test_config_shapes_forward_matches_oracle(sd_spec, name):
    from msr3d_b200.fused import FusedPrompter
    from oracle import ext_oracle as eo, modules_oracle as mo
    B, O, P = synthetic.CONFIGS[name]
    eo.set_threads(16); torch.set_num_threads(16)
    sd = synthetic.make_state_dict(sd_spec, 41)
    batch = synthetic.make_batch(B, O, P, seed=41)
    # the oracle runs scene by scene (scenes are independent): scene 0 with the index trace, the others for the tokens
    one = synthetic.SceneBatch(*(t[:1] for t in (batch.obj_fts, batch.obj_locs, batch.obj_masks,
                                                 batch.anchor_locs, batch.anchor_orientation)))
    tr = {}
    with torch.no_grad():
        want, want_mask = mo.ose3d_situation(sd, one.as_dict(), trace=tr)
        want_all = [want]
        for sc in range(1, B):
            scene = synthetic.SceneBatch(*(t[sc:sc + 1] for t in (batch.obj_fts, batch.obj_locs, batch.obj_masks,
                                                                  batch.anchor_locs, batch.anchor_orientation)))
            want_all.append(mo.ose3d_situation(sd, scene.as_dict())[0])
        want_all = torch.cat(want_all)
    model = helpers.build_prompter(41, device="cuda")
    eng = FusedPrompter(model)
    eng.prepare(B, O, P, training=False)
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(3))
    eng.load_inputs(batch.to("cuda"), up.cuda())
    tokens = eng.forward().clone()          # the engine's output buffer is reused by the next forward
    e = helpers.rel_err(tokens.cpu(), want_all)
    print(f"{name}: tokens rel err (all {B} scenes) {e:.2e}")
    assert e < 1e-3
    # FPS over 2048 points exercises the 8-points-per-thread instance and the 512-wide reference rank
    enc_tr = {}
    flat = batch.obj_fts[:1].reshape(-1, P, 6).cuda()
    model.obj_encoder.pcd_net.forward_fused(flat, trace=enc_tr)
    np.testing.assert_array_equal(enc_tr["fps1"].cpu().numpy(), tr["obj_encoder.pcd_net.encoder.0.fps_idx"].numpy())
    np.testing.assert_array_equal(enc_tr["ball1"].cpu().numpy(), tr["obj_encoder.pcd_net.encoder.0.ball_idx"].numpy())
    # size-independent properties on the whole batch: scenes are independent and deterministic
    eng.backward()
    g1 = eng.flat_g.clone()
    perm = torch.arange(B - 1, -1, -1)
    rb = synthetic.SceneBatch(*(t[perm] for t in (batch.obj_fts, batch.obj_locs, batch.obj_masks,
                                                  batch.anchor_locs, batch.anchor_orientation)))
    eng.load_inputs(rb.to("cuda"), up[perm].cuda())
    t2 = eng.forward()
    assert helpers.rel_err(t2[perm.cuda()], tokens) < 1e-5        # reversing the batch reverses the tokens
    assert torch.isfinite(g1).all()
    # gradients against the oracle's autograd (transformer half, fed the engine's own encoder output)
    from tests.test_fused_gpu import _oracle_grads, _check_grads
    eng.load_inputs(batch.to("cuda"), up.cuda())
    eng.flat_g.zero_()
    eng.forward()
    eng.backward()
    torch.cuda.synchronize()
    emb = eng.ws["emb"].cpu().view(B, O, -1)
    _, want_g = _oracle_grads(sd, batch, up, obj_embeds=emb)
    worst = _check_grads(model, want_g, tol=5e-4)
    print(f"{name}: worst gradient {worst[0]} {worst[1]:.2e}")

"""GPU: the attention fusions (mul / bias / ctx / add, modules/layers/transformers.py:207-244) and the situation modes
(as_object, as_object_add_loc, model/ose3d_situation.py:334-395) that no shipped config uses, ON THE FUSED ENGINE
(SURVEY 8f-4): fusion-templated attention kernels + spatial-weight kernels (csrc/spatial.cu), anchor-token input
assembly (csrc/modes.cu).  Checked against fixtures of the unmodified reference (tests/golden/gen_golden.py) and,
gradient by gradient, against the fp64 oracle."""
import numpy as np
import pytest
import torch

from msr3d_b200 import configs, synthetic
from tests import helpers
from tests.test_fused_gpu import _check_grads, _check_grads_own_scale, _oracle_grads

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _run_engine(model, batch, up, training=False):
    from msr3d_b200.fused import FusedPrompter
    B, O, P = batch.obj_fts.shape[:3]
    eng = FusedPrompter(model)
    eng.prepare(B, O, P, training=training)
    eng.load_inputs(batch.to("cuda"), up.cuda())
    eng.flat_g.zero_()
    tokens = eng.forward()
    eng.backward()
    torch.cuda.synchronize()
    return eng, tokens


def _check_fixture(model, tokens, fx):
    e = helpers.rel_err(tokens.cpu(), torch.from_numpy(fx["obj_tokens"]))
    assert e < TOL, e
    grads = dict(model.named_parameters())
    worst = 0.0
    for key in fx.files:
        if key.startswith("grad:"):
            ref = fx[key]
            got = grads[key[5:]].grad.detach().cpu().numpy()
            err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-3)
            worst = max(worst, err)
            assert err < 2e-3, (key, err)
    return e, worst


@pytest.mark.parametrize("fusion", ["mul", "bias", "ctx", "add"])
def test_engine_other_fusions_match_reference_and_oracle(fusion):
    fx = helpers.golden(f"prompter_fusion_{fusion}")
    seed, B, O, P = int(fx["seed"]), int(fx["B"]), int(fx["O"]), int(fx["P"])
    model = configs.build_prompter(seed, device="cuda", fusion=fusion)
    sd = configs.model_state_dict(model, seed)
    batch = synthetic.make_batch(B, O, P, seed=seed)
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(seed + 77))
    eng, tokens = _run_engine(model, batch, up)
    assert not eng.attn_block_active
    e, worst_fx = _check_fixture(model, tokens, fx)
    # every gradient against the fp64 oracle fed with the engine's own encoder output
    emb = eng.ws["emb"].cpu().view(B, O, -1)
    want_tokens, want = _oracle_grads(sd, batch, up, obj_embeds=emb, fusion=fusion)
    e64 = helpers.rel_err(tokens.cpu(), want_tokens)
    worst = _check_grads(model, want, tol=3e-4)
    own = _check_grads_own_scale(model, want, tol=2e-3)
    print(f"fusion {fusion}: tokens vs reference {e:.2e} (grads {worst_fx:.2e}), vs fp64 oracle {e64:.2e}; "
          f"worst grad {worst[0]} {worst[1]:.2e}, own scale {own[0]} {own[1]:.2e}")
    assert e64 < 1e-4


@pytest.mark.parametrize("fusion", ["bias", "add"])
def test_other_fusions_larger_scene_with_padding_and_single_spatial_head(fusion):
    """config-2-sized scenes (60 objects, 10 % padded keys), spatial_multihead False (one pairwise_loc_fc row shared
    by the heads) against the fp64 oracle."""
    seed, B, O, P = 61, 2, 60, 256
    model = configs.build_prompter(seed, device="cuda", fusion=fusion, spatial_multihead=False)
    sd = configs.model_state_dict(model, seed)
    batch = synthetic.make_batch(B, O, P, seed=seed)
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(seed + 77))
    eng, tokens = _run_engine(model, batch, up)
    emb = eng.ws["emb"].cpu().view(B, O, -1)
    want_tokens, want = _oracle_grads(sd, batch, up, obj_embeds=emb, fusion=fusion, spatial_multihead=False)
    e64 = helpers.rel_err(tokens.cpu(), want_tokens)
    worst = _check_grads(model, want, tol=3e-4)
    print(f"fusion {fusion} single spatial head: tokens {e64:.2e}, worst grad {worst[0]} {worst[1]:.2e}")
    assert e64 < 1e-4


def test_module_call_with_other_fusion_runs_on_the_kernels():
    """The drop-in module call (autograd.py) dispatches the non-shipped fusions to the engine too."""
    from msr3d_b200 import _C, autograd as ag
    seed = 5006
    model = configs.build_prompter(seed, device="cuda", fusion="bias")
    sd = configs.model_state_dict(model, seed)
    batch = synthetic.make_batch(1, 8, 256, seed=seed)
    assert ag.eligible(model, batch.to("cuda").as_dict()) is None
    n0 = _C.launch_count()
    out = model(batch.to("cuda").as_dict())
    assert _C.launch_count() > n0
    up = torch.randn(1, 8, 256, generator=torch.Generator().manual_seed(seed + 77))
    (out["obj_tokens"] * up.cuda()).sum().backward()
    fx = helpers.golden("prompter_fusion_bias")
    _check_fixture(model, out["obj_tokens"].detach(), fx)

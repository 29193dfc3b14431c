"""GPU: the one-launch attention block (csrc/attn_block.cu: cluster of 8 CTAs per scene) against the four-launch
path it replaces (QKV+cond GEMM, attention core, out-projection GEMM, LayerNorm tail) on the same inputs: every
tensor the backward pass reads, the layer outputs and the gradients that follow."""
import os

import pytest
import torch

from msr3d_b200 import synthetic
from tests import helpers

pytestmark = pytest.mark.gpu

SAVED = ("qkvc", "P", "LOC", "ctx", "xhat_a", "rstd_a", "xhat_1", "rstd_1", "y2")


def _engine(seed, B, O, P, training, attn_block):
    from msr3d_b200.fused import FusedPrompter
    old = os.environ.get("MSR3D_ATTN_BLOCK")
    os.environ["MSR3D_ATTN_BLOCK"] = "1" if attn_block else "0"
    try:
        model = helpers.build_prompter(seed, device="cuda")
        model.train(training)
        eng = FusedPrompter(model)
        eng.prepare(B, O, P, training=training)
    finally:
        if old is None:
            os.environ.pop("MSR3D_ATTN_BLOCK", None)
        else:
            os.environ["MSR3D_ATTN_BLOCK"] = old
    return model, eng


@pytest.mark.parametrize("training", [False, True])
@pytest.mark.parametrize("B,O,P,seed", [(1, 8, 256, 11), (2, 12, 256, 12), (4, 60, 1024, 13), (3, 64, 256, 14), (2, 33, 256, 15)])
def test_attn_block_equals_four_launch_path(B, O, P, seed, training):
    batch = synthetic.make_batch(B, O, P, seed=seed)
    if B > 1:    # padded keys in the last scene
        batch.obj_masks[-1, O - max(1, O // 5):] = False
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(seed + 1))
    res = []
    for ab in (False, True):
        model, eng = _engine(seed, B, O, P, training, ab)
        assert eng.attn_block_active == ab
        eng.load_inputs(batch.to("cuda"), up.cuda())
        eng.flat_g.zero_()
        tokens = eng.forward().clone()
        eng.backward()
        torch.cuda.synchronize()
        saved = {k: [t.clone() for t in eng.ws[k]] for k in SAVED}
        res.append((tokens, saved, eng.flat_g.clone()))
    (t0, s0, g0), (t1, s1, g1) = res
    for k in SAVED:
        for i, (a, b) in enumerate(zip(s0[k], s1[k])):
            if k in ("P", "LOC"):
                err = float((a - b).abs().max())
            else:
                err = helpers.rel_err(b, a)
            assert err < 1e-4, (k, i, err)
    assert helpers.rel_err(t1, t0) < 1e-4
    assert helpers.rel_err(g1, g0) < 5e-4
    assert torch.isfinite(t1).all()


def test_attn_block_deterministic():
    """The block has no atomics: two runs give bit-identical saved tensors."""
    B, O, P, seed = 4, 60, 256, 5
    batch = synthetic.make_batch(B, O, P, seed=seed)
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(seed + 1))
    model, eng = _engine(seed, B, O, P, True, True)
    eng.load_inputs(batch.to("cuda"), up.cuda())
    x0 = eng.ws["x_in"][0]
    from msr3d_b200 import _C  # noqa: F401
    eng.forward()
    torch.cuda.synchronize()
    x_ref = x0.clone()
    a = {k: eng.ws[k][0].clone() for k in SAVED}
    # rerun layer 0's block on the same layer input
    name, fn, args, tag, side = next(e for e in eng.fwd_prog if e[0].startswith("msr3d_attn_block_fwd"))
    for k in SAVED:
        eng.ws[k][0].zero_()
    x0.copy_(x_ref)
    assert fn(*args, torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    for k in SAVED:
        assert torch.equal(a[k], eng.ws[k][0]), k


@pytest.mark.parametrize("B,O,P,seed", [(1, 8, 256, 21), (4, 60, 256, 22), (2, 64, 256, 23), (3, 33, 256, 24)])
def test_attention_backward_on_tensor_cores_equals_cuda_core_kernel(B, O, P, seed):
    """msr3d_attn_bwd_tc_f32 against msr3d_attn_bwd_f32 on the tensors of a real backward pass (same contract)."""
    from msr3d_b200 import _C
    batch = synthetic.make_batch(B, O, P, seed=seed)
    if B > 1:
        batch.obj_masks[-1, O - max(1, O // 5):] = False
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(seed + 1))
    model, eng = _engine(seed, B, O, P, True, True)
    eng.load_inputs(batch.to("cuda"), up.cuda())
    eng.flat_g.zero_()
    eng.forward()
    eng.backward()
    torch.cuda.synchronize()
    w = eng.ws
    st = torch.cuda.current_stream().cuda_stream
    for i in range(eng.nl):
        outs = []
        for fn in ("msr3d_attn_bwd_f32", "msr3d_attn_bwd_tc_f32"):
            o = torch.full_like(w["dqkvc"][i], float("nan"))
            _C.call(fn, w["dctx"][i].data_ptr(), 256, w["qkvc"][i].data_ptr(), 816, w["pw"].data_ptr(), w["P"][i].data_ptr(),
                    w["LOC"][i].data_ptr(), B, O, 8, o.data_ptr(), st)
            torch.cuda.synchronize()
            outs.append(o)
        ref, got = outs
        assert torch.isfinite(got).all(), i                      # every column written
        for name, sl in (("dq", slice(0, 256)), ("dk", slice(256, 512)), ("dv", slice(512, 768)), ("dcond", slice(768, 816))):
            e = helpers.rel_err(got[:, sl], ref[:, sl])
            assert e < 2e-4, (i, name, e)

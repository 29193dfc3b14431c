"""GPU: scene-token splice (msr3d_b200/splice.py) against the fixture produced by the
unmodified reference `MSR3D.build_embeds` (tests/golden/splice.npz) and against the oracle at full size."""
import numpy as np
import pytest
import torch

from oracle import modules_oracle as mo
from tests import helpers

pytestmark = pytest.mark.gpu

# relative tolerance of one rounding to the embedding dtype (the GEMM itself is fp32-grade, split 6)
ULP = {torch.float32: 1e-5, torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}


def _run(case_tensors, dtype, D_check=None):
    from msr3d_b200 import splice
    ids, am, tok, masks, W, b, emb, G = case_tensors
    dev = "cuda"
    W, b, tok = (t.to(dev).clone().requires_grad_(True) for t in (W, b, tok))
    emb_d = emb.to(dev).to(dtype).clone()
    out, am_out = splice.splice_scene_tokens(emb_d, am.to(dev), ids.to(dev), tok, masks.to(dev), W, b)
    assert out.data_ptr() == emb_d.data_ptr()                      # in place, like the reference
    (out.float() * G.to(dev)).sum().backward()
    torch.cuda.synchronize()
    return out.detach().cpu(), am_out.cpu(), W.grad.cpu(), b.grad.cpu(), tok.grad.cpu()


@pytest.mark.parametrize("case", ["f32", "f16", "bf16"])
def test_splice_matches_reference_fixture(case):
    fx = helpers.golden("splice")
    get = lambda k: torch.from_numpy(fx[f"{case}:{k}"])  # noqa: E731
    dtype = getattr(torch, str(fx[f"{case}:dtype"]))
    out, am, dW, db, dtok = _run((get("ids"), get("attention_mask"), get("tokens"), get("masks"), get("W"),
                                  get("b"), get("emb_rows"), get("G")), dtype)
    ref = get("out")
    ids = get("ids")
    keep = ids != 31495
    assert torch.equal(out.float()[keep], ref[keep])               # untouched rows: bit-identical
    err = (out.float()[~keep] - ref[~keep]).abs().max() / ref[~keep].abs().max()
    assert err <= ULP[dtype], f"spliced rows differ: {err}"
    np.testing.assert_array_equal(am.numpy().astype(np.int64), fx[f"{case}:am_out"])
    for name, got in (("dW", dW), ("db", db), ("dtokens", dtok)):
        want = get(name)
        e = helpers.rel_err(got, want)
        assert e < 1e-4, f"{name}: rel err {e}"


@pytest.mark.parametrize("B,T,L,D,dtype", [(4, 204, 60, 4096, torch.float16), (2, 77, 9, 100, torch.bfloat16),
                                           (1, 40, 8, 256, torch.float32)])
def test_splice_matches_oracle(B, T, L, D, dtype):
    g = torch.Generator().manual_seed(B * 1000 + D)
    ids = torch.randint(0, 30000, (B, T), generator=g)
    for bb in range(B):
        s = int(torch.randint(0, T - L + 1, (1,), generator=g))
        ids[bb, s:s + L] = 31495
    am = (torch.rand(B, T, generator=g) > 0.1).long()
    tok = torch.randn(B, L, 256, generator=g)
    masks = torch.rand(B, L, generator=g) > 0.2
    W = torch.randn(D, 256, generator=g) * 0.02
    b = torch.randn(D, generator=g) * 0.02
    emb = (torch.randn(B, T, D, generator=g) * 0.02).to(dtype).float()
    G = torch.randn(B, T, D, generator=g).to(dtype).float()
    out, am_out, dW, db, dtok = _run((ids, am, tok, masks, W, b, emb, G), dtype)
    Wr, br, tr = (t.clone().requires_grad_(True) for t in (W, b, tok))
    want, want_am = mo.scene_splice(Wr, br, emb.to(dtype), am, ids, tr, masks)
    (want.float() * G).sum().backward()
    keep = ids != 31495
    assert torch.equal(out[keep], want.detach()[keep])
    w = want.detach().float()[~keep]
    err = (out.float()[~keep] - w).abs().max() / w.abs().max()
    assert err <= ULP[dtype], f"spliced rows differ from the oracle: {err}"
    assert torch.equal(am_out, want_am)
    for name, got, ref in (("dW", dW, Wr.grad), ("db", db, br.grad), ("dtokens", dtok, tr.grad)):
        e = helpers.rel_err(got, ref)
        assert e < 2e-4, f"{name}: rel err {e}"


def test_splice_placeholder_count_mismatch_raises():
    from msr3d_b200 import splice
    ids = torch.zeros(2, 16, dtype=torch.int64, device="cuda")
    ids[0, 3:6] = 31495                                            # 3 placeholders for 2 x 2 tokens
    emb = torch.zeros(2, 16, 64, device="cuda", dtype=torch.float16)
    with pytest.raises(RuntimeError, match="shape mismatch"):
        splice.splice_scene_tokens(emb, torch.ones(2, 16, dtype=torch.long, device="cuda"), ids,
                                   torch.zeros(2, 2, 256, device="cuda"), torch.ones(2, 2, dtype=torch.bool, device="cuda"),
                                   torch.zeros(64, 256, device="cuda"), None)

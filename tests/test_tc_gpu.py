"""GPU: tensor-core kernels (tcgen05 split-bf16) against fp64 / oracle references."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from msr3d_b200 import ops as o
    return o


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


LIN_CASES = [(240, 768, 768, "none"), (3840, 259, 256, "relu"), (300, 63, 256, "none"),
             (128, 64, 128, "gelu"), (1, 16, 8, "none"), (515, 130, 200, "relu"), (240, 256, 2048, "gelu")]


@pytest.mark.parametrize("M,K,N,act", LIN_CASES)
def test_linear_tc_split3_is_fp32_grade(ops, M, K, N, act):
    g = torch.Generator().manual_seed(M + K + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    want = A.double() @ W.double().t() + b.double()
    want = {"none": lambda x: x, "relu": torch.relu, "gelu": torch.nn.functional.gelu}[act](want)
    pw = ops.pack_weight(W.cuda())
    got = ops.linear_tc(A.cuda(), pw, b.cuda(), act=act, split=3).cpu().double()
    assert _rel(got, want) < 2e-5, _rel(got, want)
    got1 = ops.linear_tc(A.cuda(), pw, b.cuda(), act=act, split=1).cpu().double()
    assert _rel(got1, want) < 3e-2           # single bf16 pass: sanity only
    print(f"linear {M}x{K}x{N} {act}: split3 rel {_rel(got, want):.2e}  split1 rel {_rel(got1, want):.2e}")


def test_linear_tc_strided_pool_accumulate(ops):
    g = torch.Generator().manual_seed(3)
    A = torch.randn(96, 320, generator=g)
    W = torch.randn(70, 96, generator=g) * 0.1
    # transposed A view (stride_ak != 1), as used by the weight-gradient GEMMs
    want = A.t().double() @ W.double().t()
    pw = ops.pack_weight(W.cuda())
    got = ops.linear_tc(A.cuda().t(), pw, None).cpu().double()
    assert _rel(got, want) < 2e-5
    acc = torch.ones(320, 70).cuda()
    ops.linear_tc(A.cuda().t(), pw, None, out=acc, accumulate=True)
    assert _rel(acc.cpu().double(), want + 1) < 2e-5
    # transposed weight view (stride_k != 1), as used by the data-gradient GEMMs
    pwt = ops.pack_weight(W.cuda().t())          # (96 x 70) logical
    X = torch.randn(50, 70, generator=g)
    assert _rel(ops.linear_tc(X.cuda(), pwt, None).cpu().double(), X.double() @ W.double()) < 2e-5
    # max-pool over groups of 16 rows (SA3 / GroupAll)
    A2 = torch.randn(64, 40, generator=g)
    W2 = torch.randn(33, 40, generator=g)
    y = torch.relu(A2.double() @ W2.double().t()).view(4, 16, 33).max(1).values
    got = ops.linear_tc(A2.cuda(), ops.pack_weight(W2.cuda()), None, act="relu", pool=16).cpu().double()
    assert got.shape == (4, 33) and _rel(got, y) < 2e-5


def test_pack_weight_scale_and_permutation(ops):
    g = torch.Generator().manual_seed(4)
    W = torch.randn(64, 6, generator=g)
    s = torch.rand(64, generator=g) + 0.5
    perm = torch.full((64,), -1, dtype=torch.int32)
    perm[:6] = torch.tensor([3, 4, 5, 0, 1, 2], dtype=torch.int32)   # [rgb | xyz]
    pw = ops.pack_weight(W.cuda(), row_scale=s.cuda(), col_perm=perm.cuda(), K=6)
    X = torch.randn(20, 6, generator=g)
    want = X[:, [3, 4, 5, 0, 1, 2]].double() @ (W.double() * s.double()[:, None]).t()
    Xp = X.cuda()
    got = ops.linear_tc(Xp, pw, None).cpu().double()
    # packed column j multiplies A column j, whose source weight column is perm[j]
    want = X.double() @ (W.double() * s.double()[:, None])[:, [3, 4, 5, 0, 1, 2]].t()
    assert _rel(got, want) < 2e-5


@pytest.mark.parametrize("level", ["sa1", "sa2"])
def test_sa_level_fused_matches_oracle(ops, level, sd_spec):
    from msr3d_b200 import synthetic
    from oracle import ext_oracle as eo, modules_oracle as mo
    sd = synthetic.make_state_dict(sd_spec, 5)
    g = torch.Generator().manual_seed(8)
    if level == "sa1":
        B, N, npoint, radius, nfeat, widths, pre = 6, 1024, 32, 0.2, 3, (64, 64, 128), "obj_encoder.pcd_net.encoder.0"
        batch = synthetic.make_batch(1, B, N, seed=12)
        rec = batch.obj_fts[0].contiguous()                     # (B, N, 6) xyz|rgb
        xyz, feat_cn = rec[..., :3].contiguous(), rec[..., 3:].transpose(1, 2).contiguous()
    else:
        B, N, npoint, radius, nfeat, widths, pre = 5, 32, 16, 0.4, 128, (128, 128, 256), "obj_encoder.pcd_net.encoder.1"
        xyz = (torch.randn(B, N, 3, generator=g) * 0.3).contiguous()
        feat_cn = torch.relu(torch.randn(B, nfeat, N, generator=g)).contiguous()
    eo.set_threads(4)
    tr = {}
    with torch.no_grad():
        want_xyz, want = mo.sa_module(sd, pre, xyz, feat_cn, npoint, radius, 32, tr)   # (B,C3,npoint)
    # --- CUDA path
    dev = "cuda"
    if level == "sa1":
        rec_d = rec.to(dev)
        xyz_v, feat_v = rec_d, rec_d[..., 3:]
    else:
        xyz_v = xyz.to(dev)
        feat_v = feat_cn.transpose(1, 2).contiguous().to(dev)     # point-major (B,N,128)
    fps_idx, new_xyz, ball = ops.sa_group(xyz_v, npoint, radius, 32)
    np.testing.assert_array_equal(fps_idx.cpu().numpy(), tr[pre + ".fps_idx"].numpy())
    np.testing.assert_array_equal(ball.cpu().numpy(), tr[pre + ".ball_idx"].numpy())
    assert torch.equal(new_xyz.cpu(), want_xyz)
    from msr3d_b200.encoder import fold_sa_level
    weights, biases = fold_sa_level({k: v.to(dev) for k, v in sd.items() if k.startswith(pre)}, pre + ".mlps.0", nfeat, fused=True)
    out = torch.empty(B, npoint, widths[2], device=dev)
    ops.sa_mlp_fused(feat_v, nfeat, xyz_v, new_xyz, ball, weights, biases, widths, out, split=3)
    got = out.cpu().transpose(1, 2)
    rel = _rel(got.double(), want.double())
    print(level, "fused rel err", rel)
    assert rel < 1e-4


def test_gemm_tc_strided_operands_and_three_way_split():
    """fp32 operands on both sides through strides; split 6 is fp32-grade on a cancelling sum."""
    from msr3d_b200.fused import gemm_tc
    g = torch.Generator().manual_seed(17)
    dY = torch.randn(240, 816, generator=g)
    X = torch.randn(240, 256, generator=g)
    X = X + 50.0                                  # large common offset: sum over tokens cancels heavily
    dY = dY - dY.mean(0, keepdim=True)
    want = dY.double().t() @ X.double()           # dW = dY^T X   (816 x 256), contraction over tokens
    a, b = dY.cuda().t(), X.cuda().t()            # (816 x 240) and (256 x 240) views, k-stride != 1
    e3 = _rel(gemm_tc(a, b, split=3).cpu().double(), want)
    e6 = _rel(gemm_tc(a, b, split=6).cpu().double(), want)
    e32 = _rel((dY.t() @ X).double(), want)
    print(f"cancelling dW: split3 {e3:.2e}  split6 {e6:.2e}  torch-fp32-cpu {e32:.2e}")
    assert e6 < e3 / 3 and e6 < 5 * e32 + 1e-6     # within a small factor of fp32 accumulation itself
    # split-K with atomics + bias/residual only once
    W = torch.randn(256, 2048, generator=g) * 0.02
    A = torch.randn(240, 2048, generator=g)
    bias, res = torch.randn(256, generator=g), torch.randn(240, 256, generator=g)
    out = torch.zeros(240, 256).cuda()
    gemm_tc(A.cuda(), W.cuda(), bias=bias.cuda(), residual=res.cuda(), out=out, accumulate=2, ksplit=8)
    assert _rel(out.cpu().double(), A.double() @ W.double().t() + bias.double() + res.double()) < 2e-5

"""GPU: TMA-fed 3xTF32 GEMM (csrc/gemm3x.cu) against fp64 references — every operand orientation the
transformer uses (weights, their transposes, token-contracting weight gradients), tails, split-K, epilogues."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _mk(shape, seed, transposed):
    """fp32 CUDA matrix of logical `shape`, stored row-major or as the transpose of a row-major tensor."""
    g = torch.Generator().manual_seed(seed)
    if transposed:
        return torch.randn(shape[1], shape[0], generator=g).cuda().t()
    return torch.randn(*shape, generator=g).cuda()


# (M, N, K, A transposed, B transposed, ksplit, tile_n)
CASES = [
    (240, 816, 256, False, False, 1, 0),     # qkvc
    (240, 816, 256, False, False, 4, 64),    # qkvc, split-K
    (240, 256, 256, False, False, 2, 64),    # out_proj
    (240, 2048, 256, False, False, 1, 64),   # ffn1
    (240, 256, 2048, False, False, 16, 64),  # ffn2
    (240, 2048, 256, False, True, 1, 64),    # du: B = W2^T (stored [k][n])
    (240, 256, 816, False, True, 13, 64),    # dx: B = Wqkvc^T
    (256, 2048, 240, True, True, 1, 128),    # dW2 = df^T h (both operands token-major)
    (816, 256, 240, True, True, 2, 128),     # dWqkvc
    (240, 256, 768, False, False, 6, 64),    # obj_proj
    (60, 64, 40, False, False, 1, 64),       # tiny, K tail (40 = 32 + 8)
    (300, 200, 100, True, False, 1, 128),    # A transposed only, ragged everything (N % 16 != 0)
    (129, 72, 36, False, True, 1, 0),        # B transposed only
]


@pytest.mark.parametrize("M,N,K,ta,tb,ksplit,tile_n", CASES)
def test_gemm3x_matches_fp64(M, N, K, ta, tb, ksplit, tile_n):
    from msr3d_b200.fused import gemm_tc
    A = _mk((M, K), 1000 + M + N + K, ta)
    Bm = _mk((N, K), 2000 + M + N + K, tb) / K ** 0.5
    # pitches must be multiples of 4 floats: pad by slicing a wider buffer when the contiguous extent is not
    def pitch_ok(t):
        ld = t.stride(0) if t.stride(1) == 1 else t.stride(1)
        return ld % 4 == 0
    if not pitch_ok(A) or not pitch_ok(Bm):
        pytest.skip("pitch not a multiple of 4 floats (the engine routes such operands to linear_tc)")
    bias = torch.randn(N, generator=torch.Generator().manual_seed(5)).cuda()
    want = A.double() @ Bm.double().t() + bias.double()
    out = torch.zeros(M, N, device="cuda") if ksplit > 1 else None
    got = gemm_tc(A, Bm, bias, out=out, accumulate=2 if ksplit > 1 else 0, ksplit=ksplit, impl="tf32x3", tile_n=tile_n)
    torch.cuda.synchronize()
    e = _rel(got, want)
    print(f"gemm3x {M}x{N}x{K} ta={ta} tb={tb} ksplit={ksplit} bn={tile_n}: rel {e:.2e}")
    assert e < 5e-6, e


def test_gemm3x_is_fp32_grade_where_bf16_split_is_not():
    """Badly conditioned sum: 3xTF32 must sit at the fp32 level (compare with torch fp32 matmul)."""
    from msr3d_b200.fused import gemm_tc
    g = torch.Generator().manual_seed(7)
    A = torch.randn(240, 256, generator=g).cuda() + 3.0
    Bm = torch.randn(256, 256, generator=g).cuda()
    want = A.double() @ Bm.double().t()
    got = gemm_tc(A, Bm, impl="tf32x3", tile_n=64)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref32 = A @ Bm.t()
    e, e32 = _rel(got, want), _rel(ref32, want)
    print(f"3xTF32 rel {e:.2e}; torch fp32 rel {e32:.2e}")
    assert e < 4 * max(e32, 2e-7)


def test_gemm3x_epilogues():
    from msr3d_b200.fused import gemm_tc
    g = torch.Generator().manual_seed(11)
    A = torch.randn(240, 256, generator=g).cuda()
    W = (torch.randn(512, 256, generator=g) / 16).cuda()
    aux = torch.randn(240, 512, generator=g).cuda()
    res = torch.randn(240, 512, generator=g).cuda()
    want = (A.double() @ W.double().t()) * aux.double() + res.double()
    got = gemm_tc(A, W, impl="tf32x3", aux=aux, residual=res, tile_n=64)
    assert _rel(got, want) < 5e-6
    # red.add accumulation on top of existing contents
    base = torch.ones(240, 512, device="cuda")
    got = gemm_tc(A, W, impl="tf32x3", out=base.clone(), accumulate=2, ksplit=4, tile_n=128)
    assert _rel(got, A.double() @ W.double().t() + 1) < 5e-6


def _dropout_mask(seed, p, site, n):
    import ctypes
    from msr3d_b200 import _C, fused  # noqa: F401  (fused registers the prototype)
    out = torch.empty(n, device="cuda")
    _C.call("msr3d_dropout_mask_f32", ctypes.c_ulonglong(seed), None, ctypes.c_float(p), site, n, out.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
    return out


@pytest.mark.parametrize("M,F,p,fs", [(240, 2048, 0.0, 64), (240, 2048, 0.1, 64), (240, 2048, 0.1, 32), (60, 256, 0.0, 64),
                                      (300, 512, 0.1, 32)])
def test_ffn_pair_forward_and_backward(M, F, p, fs):
    """Fused FFN pair (csrc/gemm3x.cu) against fp64: f, h, gelu'*mask forward; du, dy2 backward."""
    import ctypes
    from msr3d_b200 import _C, fused  # noqa: F401
    g = torch.Generator().manual_seed(M + F + fs)
    y2 = torch.randn(M, 256, generator=g).cuda()
    W1 = (torch.randn(F, 256, generator=g) / 16).cuda()
    b1 = (torch.randn(F, generator=g) * 0.1).cuda()
    W2 = (torch.randn(256, F, generator=g) / F ** 0.5).cuda()
    b2 = (torch.randn(256, generator=g) * 0.1).cuda()
    seed, site = 1234, 5
    h = torch.empty(M, F, device="cuda")
    gp = torch.empty(M, F, device="cuda")
    f = torch.zeros(M, 256, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _C.call("msr3d_ffn_pair_fwd_f32", y2.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), M, F,
            ctypes.c_ulonglong(seed), None, ctypes.c_float(p), site, h.data_ptr(), gp.data_ptr(), f.data_ptr(), fs, st)
    torch.cuda.synchronize()
    mask = _dropout_mask(seed, p, site, M * F).view(M, F).double() if p > 0 else torch.ones(M, F, dtype=torch.float64, device="cuda")
    u = y2.double() @ W1.double().t() + b1.double()
    cdf = 0.5 * (1 + torch.erf(u / 2 ** 0.5))
    pdf = torch.exp(-0.5 * u * u) / (2 * torch.pi) ** 0.5
    h_want = u * cdf * mask
    gp_want = (cdf + u * pdf) * mask
    f_want = h_want @ W2.double().t() + b2.double()
    eh, eg, ef = _rel(h, h_want), _rel(gp, gp_want), _rel(f, f_want)
    print(f"ffn_pair fwd M={M} F={F} p={p} fs={fs}: h {eh:.2e} gp {eg:.2e} f {ef:.2e}")
    assert eh < 5e-6 and eg < 1e-5 and ef < 5e-6
    # backward
    df = torch.randn(M, 256, generator=g).cuda()
    dy2 = torch.randn(M, 256, generator=g).cuda()       # residual-path gradient already in the buffer
    dy2_0 = dy2.clone()
    du = torch.empty(M, F, device="cuda")
    _C.call("msr3d_ffn_pair_bwd_f32", df.data_ptr(), W1.data_ptr(), W2.data_ptr(), gp.data_ptr(), M, F, du.data_ptr(),
            dy2.data_ptr(), fs, st)
    torch.cuda.synchronize()
    du_want = (df.double() @ W2.double()) * gp.double()
    dy2_want = dy2_0.double() + du_want @ W1.double()
    ed, ey = _rel(du, du_want), _rel(dy2, dy2_want)
    print(f"ffn_pair bwd: du {ed:.2e} dy2 {ey:.2e}")
    assert ed < 5e-6 and ey < 5e-6

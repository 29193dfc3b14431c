"""Per-phase timelines (SM clocks written by the kernels themselves) and in-chain latencies of the TMA-fed
kernels (csrc/gemm3x.cu).   python tools/trace_kernels.py   (on the GPU box)
In-chain = CUDA graph of 40 back-to-back launches of the same kernel (PDL between them), warm L2."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from msr3d_b200 import _C, fused  # noqa: E402

dev = torch.device("cuda")
lib = _C.lib()
lib.msr3d_set_trace_buffer.argtypes = [ctypes.c_void_p]
MHZ = 1965.0
SLOTS = ["entry", "setup", "pdl_wait", "chunk0_landed", "conv_done", "mma1_issued", "acc1_seen", "epi1_done",
         "b2_landed", "b2_conv", "acc2_seen", "end"]


def chain_us(launch, n=40, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            launch(s.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(n):
            launch(s.cuda_stream)
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (n * reps)


def timeline(launch, n_cta):
    buf = torch.zeros(n_cta * 16, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        launch(st)                      # warm
    torch.cuda.synchronize()
    lib.msr3d_set_trace_buffer(buf.data_ptr())
    launch(st)
    torch.cuda.synchronize()
    lib.msr3d_set_trace_buffer(None)
    t = buf.view(n_cta, 16).cpu().double()
    rel = (t - t[:, :1]) / MHZ          # us since this CTA's entry
    out = []
    for i, name in enumerate(SLOTS):
        col = rel[:, i][t[:, i] > 0]
        if len(col):
            out.append(f"{name}={col.mean():.2f}(max {col.max():.2f})")
    return "  ".join(out)


def gemm_launcher(M, N, K, ta=False, tb=False, ksplit=1, tile_n=64, b_static=1):
    A = torch.randn(K, M, device=dev).t() if ta else torch.randn(M, K, device=dev)
    Bm = torch.randn(K, N, device=dev).t() if tb else torch.randn(N, K, device=dev)
    out = torch.zeros(M, N, device=dev)
    d = fused.GemmDesc()
    d.A, d.stride_am, d.stride_ak = A.data_ptr(), A.stride(0), A.stride(1)
    d.B, d.stride_bn, d.stride_bk = Bm.data_ptr(), Bm.stride(0), Bm.stride(1)
    d.out, d.ld_out, d.M, d.N, d.K = out.data_ptr(), N, M, N, K
    d.accumulate, d.ksplit, d.tile_n = (2 if ksplit > 1 else 0), ksplit, tile_n
    keep = (A, Bm, out, d)

    def launch(st):
        rc = lib.msr3d_gemm_tf32x3_f32(ctypes.byref(d), b_static, st)
        assert rc == 0, lib.msr3d_last_error()
    launch.keep = keep
    bn = tile_n
    ctas = -(-M // 128) * -(-N // bn) * ksplit
    return launch, ctas


def ffn_launcher(M, F, fs, bwd, p=0.1):
    y2 = torch.randn(M, 256, device=dev)
    W1 = torch.randn(F, 256, device=dev) / 16
    b1 = torch.randn(F, device=dev)
    W2 = torch.randn(256, F, device=dev) / 45
    b2 = torch.randn(256, device=dev)
    h, gp, du = (torch.empty(M, F, device=dev) for _ in range(3))
    f = torch.zeros(M, 256, device=dev)
    keep = (y2, W1, b1, W2, b2, h, gp, du, f)

    def launch(st):
        if not bwd:
            rc = lib.msr3d_ffn_pair_fwd_f32(y2.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), M, F,
                                            1234, None, p, 3, h.data_ptr(), gp.data_ptr(), f.data_ptr(), fs, st)
        else:
            rc = lib.msr3d_ffn_pair_bwd_f32(y2.data_ptr(), W1.data_ptr(), W2.data_ptr(), gp.data_ptr(), M, F, du.data_ptr(),
                                            f.data_ptr(), fs, st)
        assert rc == 0, lib.msr3d_last_error()
    launch.keep = keep
    return launch, -(-M // 128) * (F // fs)


if __name__ == "__main__":
    cases = [("qkvc 240x816x256 bn64 ks4", gemm_launcher(240, 816, 256, ksplit=4)),
             ("qkvc 240x816x256 bn64 ks1", gemm_launcher(240, 816, 256, ksplit=1)),
             ("out_proj 240x256x256 bn64 ks4", gemm_launcher(240, 256, 256, ksplit=4)),
             ("out_proj 240x256x256 bn64 ks8", gemm_launcher(240, 256, 256, ksplit=8)),
             ("dx 240x256x816 tb bn64 ks13", gemm_launcher(240, 256, 816, tb=True, ksplit=13)),
             ("dW2 256x2048x240 ta tb bn128", gemm_launcher(256, 2048, 240, ta=True, tb=True, tile_n=128, b_static=0)),
             ("dW2 256x2048x240 ta tb bn64", gemm_launcher(256, 2048, 240, ta=True, tb=True, tile_n=64, b_static=0)),
             ("ffn_pair fwd fs64", ffn_launcher(240, 2048, 64, False)),
             ("ffn_pair fwd fs32", ffn_launcher(240, 2048, 32, False)),
             ("ffn_pair bwd fs64", ffn_launcher(240, 2048, 64, True)),
             ("ffn_pair bwd fs32", ffn_launcher(240, 2048, 32, True))]
    for name, (la, ctas) in cases:
        us = chain_us(la)
        print(f"{us:7.2f} us in-chain  {name}  [{ctas} CTAs]")
        print("         " + timeline(la, ctas))

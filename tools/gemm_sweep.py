"""In-chain latency of the transformer's kernels under different launch heuristics (on the GPU box).

    python tools/gemm_sweep.py [--json gpurun_out/sweep.json]

Every variant is captured as a CUDA graph of CHAIN back-to-back launches of the same kernel on one
stream (the dependency pattern of the real step: programmatic dependent launch between consecutive
small kernels) and replayed; the figure reported is microseconds per launch inside the chain.  The
launch heuristics in csrc/linear_tc.cu / csrc/spatial.cu are chosen from this table.
"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from msr3d_b200 import _C, fused  # noqa: E402
from msr3d_b200.fused import GemmDesc  # noqa: E402

_C.register_protos({"msr3d_set_tunable": [ctypes.c_int, ctypes.c_int]})
TUNE = {"bn": 0, "ksplit": 1, "attn_fwd_warps": 2, "attn_bwd_warps": 3, "attn_fwd_z": 4, "attn_bwd_cluster": 5}
CHAIN, REPLAYS = 40, 25
dev = torch.device("cuda")


def tune(**kw):
    for k in TUNE:
        _C.call("msr3d_set_tunable", TUNE[k], int(kw.get(k, 0)))


def chain_us(launch):
    """launch(stream) enqueues ONE kernel; returns us per launch inside a graph-replayed chain."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            launch(s.cuda_stream)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(CHAIN):
            launch(st)
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REPLAYS):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (CHAIN * REPLAYS)


def gemm_launcher(M, N, K, split, accumulate=0, ksplit=1, transposed_a=False, transposed_b=False, bias=True,
                  epilogue=0):
    A = torch.randn(K, M, device=dev).t() if transposed_a else torch.randn(M, K, device=dev)
    Bm = torch.randn(K, N, device=dev).t() if transposed_b else torch.randn(N, K, device=dev)
    out = torch.zeros(M, N, device=dev)
    aux = torch.zeros(M, N, device=dev)
    bv = torch.zeros(N, device=dev)
    seed = torch.zeros(1, dtype=torch.int64, device=dev)
    d = GemmDesc()
    d.A, d.stride_am, d.stride_ak = A.data_ptr(), A.stride(0), A.stride(1)
    d.B, d.stride_bn, d.stride_bk = Bm.data_ptr(), Bm.stride(0), Bm.stride(1)
    d.bias = bv.data_ptr() if bias else None
    d.out, d.ld_out, d.M, d.N, d.K = out.data_ptr(), N, M, N, K
    d.split, d.accumulate, d.ksplit, d.epilogue = split, accumulate, ksplit, epilogue
    if epilogue == 2:
        d.aux_out, d.ld_aux, d.dropout_p, d.seed_ptr = aux.data_ptr(), N, 0.1, seed.data_ptr()
    if epilogue == 3:
        d.aux, d.ld_aux = aux.data_ptr(), N
    keep = (A, Bm, out, aux, bv, seed, d)
    fn = _C.lib().msr3d_gemm_tc_f32

    def launch(stream):
        rc = fn(ctypes.byref(d), stream)
        assert rc == 0, _C.lib().msr3d_last_error()
    launch.keep = keep
    return launch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--tokens", type=int, default=240)
    args = ap.parse_args()
    M, H, F, Q = args.tokens, 256, 2048, fused.QKVC
    res = {}

    # floor: an (almost) empty kernel in the same chain
    tiny = torch.zeros(32, device=dev)
    fn = _C.lib().msr3d_dropout_mask_f32
    fn.argtypes = [ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_float, ctypes.c_uint, ctypes.c_longlong,
                   ctypes.c_void_p, ctypes.c_void_p]
    for pdl in (1, 0):
        _C.lib().msr3d_set_pdl(pdl)
        res[f"empty_kernel pdl={pdl}"] = round(chain_us(lambda st: fn(1, None, 0.0, 0, 32, tiny.data_ptr(), st)), 2)
    _C.lib().msr3d_set_pdl(1)

    # (tag, M, N, K, kwargs, split-K candidates)
    shapes = [
        ("qkvc    240x816x256", M, Q, H, dict(), [1, 2, 4]),
        ("out_proj 240x256x256", M, H, H, dict(), [1, 2, 4]),
        ("ffn1    240x2048x256 (gelu epilogue)", M, F, H, dict(epilogue=2), [1]),
        ("ffn2    240x256x2048", M, H, F, dict(), [4, 8, 16, 32]),
        ("du      240x2048x256 (W^T, mul epilogue)", M, F, H, dict(transposed_b=True, epilogue=3, bias=False), [1]),
        ("dy2     240x256x2048 (W^T)", M, H, F, dict(transposed_b=True, bias=False), [4, 8, 16, 32]),
        ("dx      240x256x816 (W^T)", M, H, Q, dict(transposed_b=True, bias=False), [2, 4, 13]),
        ("dW2     256x2048x240 (A^T,B^T)", H, F, M, dict(transposed_a=True, transposed_b=True, bias=False), [1, 2, 4]),
        ("dWqkvc  816x256x240 (A^T,B^T)", Q, H, M, dict(transposed_a=True, transposed_b=True, bias=False), [1, 2, 4]),
        ("obj_proj 240x256x768", M, H, 768, dict(), [1, 3, 6, 12]),
        ("loc_fc  240x256x63", M, H, 63, dict(), [1]),
        ("splice  240x4096x256", M, 4096, H, dict(), [1]),
    ]
    for split in (6,):
        for tag, m, n, k, kw, kss in shapes:
            for bn in (128, 64):
                for ks in kss:
                    acc = 2 if ks > 1 else 0
                    if ks > 1 and kw.get("epilogue", 0) != 0:
                        continue
                    tune(bn=bn)
                    la = gemm_launcher(m, n, k, split, accumulate=acc, ksplit=ks, **kw)
                    res[f"gemm split={split} {tag} bn={bn} ksplit={ks}"] = round(chain_us(la), 2)
    tune()

    # attention core
    B, L = M // 60, 60
    qkvc = torch.randn(M, Q, device=dev)
    pw = torch.rand(B, L, L, 5, device=dev)
    pad = torch.zeros(B, L, dtype=torch.uint8, device=dev)
    P = torch.zeros(B, 8, L, L, device=dev)
    LOC = torch.zeros(B, 8, L, L, device=dev)
    ctx = torch.zeros(M, H, device=dev)
    dqkvc = torch.zeros(M, Q, device=dev)
    lib = _C.lib()
    for warps in (8, 16):
        for z in (2, 4, 8):
            tune(attn_fwd_warps=warps, attn_fwd_z=z)
            res[f"attn_fwd warps={warps} z={z}"] = round(chain_us(lambda st: lib.msr3d_attn_fwd_f32(
                qkvc.data_ptr(), Q, pw.data_ptr(), pad.data_ptr(), B, L, 8, P.data_ptr(), LOC.data_ptr(),
                ctx.data_ptr(), H, st)), 2)
    lib.msr3d_attn_fwd_f32(qkvc.data_ptr(), Q, pw.data_ptr(), pad.data_ptr(), B, L, 8, P.data_ptr(), LOC.data_ptr(),
                           ctx.data_ptr(), H, torch.cuda.current_stream().cuda_stream)
    for warps in (8, 16):
        for cl in (1, 2, 4, 8):
            tune(attn_bwd_warps=warps, attn_bwd_cluster=cl)
            res[f"attn_bwd warps={warps} cluster={cl}"] = round(chain_us(lambda st: lib.msr3d_attn_bwd_f32(
                ctx.data_ptr(), H, qkvc.data_ptr(), Q, pw.data_ptr(), P.data_ptr(), LOC.data_ptr(), B, L, 8,
                dqkvc.data_ptr(), st)), 2)
    tune()
    for k, v in res.items():
        print(f"{v:9.2f} us  {k}")
    if args.json:
        os.makedirs(os.path.dirname(args.json) or ".", exist_ok=True)
        with open(args.json, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()

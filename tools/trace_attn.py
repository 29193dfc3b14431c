"""attn_fwd inside a realistic chain (qkvc GEMM -> attention, repeated): in-chain latency and the kernel's own
timeline (SM clocks).   python tools/trace_attn.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from msr3d_b200 import _C, fused  # noqa: E402
from tools.trace_kernels import gemm_launcher, chain_us, lib, MHZ  # noqa: E402

dev = torch.device("cuda")
B, L, H, Q = 4, 60, 8, fused.QKVC
M = B * L
x = torch.randn(M, 256, device=dev)
W = torch.randn(Q, 256, device=dev) / 16
qkvc = torch.zeros(M, Q, device=dev)
pw = torch.rand(B, L, L, 5, device=dev)
pad = torch.zeros(B, L, dtype=torch.uint8, device=dev)
P = torch.zeros(B, H, L, L, device=dev)
LOC = torch.zeros(B, H, L, L, device=dev)
ctx = torch.zeros(M, 256, device=dev)
d = fused.GemmDesc()
d.A, d.stride_am, d.stride_ak = x.data_ptr(), 256, 1
d.B, d.stride_bn, d.stride_bk = W.data_ptr(), 256, 1
d.out, d.ld_out, d.M, d.N, d.K = qkvc.data_ptr(), Q, M, Q, 256
d.accumulate, d.ksplit, d.tile_n = 2, 4, 64


def gemm(st):
    assert lib.msr3d_gemm_tf32x3_f32(ctypes.byref(d), 1, st) == 0


def attn(st):
    assert lib.msr3d_attn_fwd_f32(qkvc.data_ptr(), Q, pw.data_ptr(), pad.data_ptr(), B, L, H, P.data_ptr(),
                                  LOC.data_ptr(), ctx.data_ptr(), 256, st) == 0


def both(st):
    gemm(st)
    attn(st)


print(f"gemm alone {chain_us(gemm):.2f} us, attn alone {chain_us(attn):.2f} us, gemm+attn {chain_us(both):.2f} us")
n_cta = 8 * 4 * 4
buf = torch.zeros(256 * 16, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    both(st)
torch.cuda.synchronize()
lib.msr3d_set_trace_buffer(buf.data_ptr())
both(st)
torch.cuda.synchronize()
lib.msr3d_set_trace_buffer(None)
t = buf.view(256, 16)[:n_cta].cpu().double()
for name, slot in (("pdl_wait_done", 13), ("staged", 14), ("softmax", 15), ("end", 8)):
    rel = (t[:, slot] - t[:, 12]) / MHZ
    print(f"{name}: mean {rel.mean():.2f} us  max {rel.max():.2f} us")

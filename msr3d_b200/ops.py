"""Tensor-level wrappers over the C ABI (groups B-D).  Everything here launches
hand-written kernels from libmsr3d_b200.so on torch's current CUDA stream; torch is
used for allocation only."""
from __future__ import annotations

from typing import Optional

import torch

from . import _C

ACT = {"none": 0, None: 0, "relu": 1, "gelu": 2}
#: 3 = split-bf16 (fp32-grade), 1 = single bf16 pass.  Module default; tests flip it.
DEFAULT_SPLIT = 3
NUM_SMS = 148   # B200
#: 192-wide N tiles for the last SA3 layer when 128-wide ones spill into a second wave: 10 us faster for
#: the encoder alone, 5 us SLOWER per pipelined step (120 CTAs that each own an SM for the whole kernel
#: starve the concurrent transformer chain) -> off unless MSR3D_SA3_WIDE=1
#: N tile of the packed weights of GEMMs that run on few CTAs (final fc: M = 240 rows; first SA3 layer):
#: 64 doubles their CTA count.  A/B: MSR3D_ENC_BN64=0/1
ENC_SMALL_TILES = __import__("os").environ.get("MSR3D_ENC_BN64", "0") == "1"
SA3_WIDE_TILES = __import__("os").environ.get("MSR3D_SA3_WIDE", "0") == "1"
SA3_CHAIN = __import__("os").environ.get("MSR3D_SA3_CHAIN", "1") == "1"     # SA3 layers hand their outputs on pre-split (bf16 hi / lo)
#: roofline.KernelTimer or None; when set, launches are bracketed by CUDA events on the current stream
TIMER = None


class _NoRange:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _range(name):
    return TIMER.range(name) if (TIMER is not None and name) else _NoRange()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


class PackedWeight:
    """Shared-memory image of an (N x K) weight for tcgen05 (see csrc/linear_tc.cu)."""

    __slots__ = ("data", "N", "K", "BN", "BK")

    def __init__(self, data, N, K, BN, BK):
        self.data, self.N, self.K, self.BN, self.BK = data, N, K, BN, BK


def pack_weight(W: torch.Tensor, row_scale: Optional[torch.Tensor] = None,
                col_perm: Optional[torch.Tensor] = None, BN: int = 128, BK: int = 64,
                out: Optional[torch.Tensor] = None, K: Optional[int] = None) -> PackedWeight:
    """W: 2-D float32 CUDA tensor (N, K_src), any strides.  K = packed logical width
    (defaults to K_src, or len(col_perm))."""
    assert W.is_cuda and W.dtype == torch.float32 and W.dim() == 2
    N, Ksrc = W.shape
    Kl = int(K if K is not None else (col_perm.numel() if col_perm is not None else Ksrc))
    nbytes = _C.lib().msr3d_packed_weight_bytes(N, Kl, BN, BK)
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    assert out.numel() >= nbytes
    if col_perm is not None:
        Kp = -(-Kl // BK) * BK
        assert col_perm.dtype == torch.int32 and col_perm.numel() == Kp and col_perm.is_cuda
    # with a permutation the kernel bounds-checks source columns against K_src
    with torch.cuda.device(W.device):
        _C.call("msr3d_pack_weight_f32", W.data_ptr(), N, Ksrc if col_perm is not None else Kl,
                W.stride(0), W.stride(1), _ptr(row_scale), _ptr(col_perm), BN, BK,
                out.data_ptr(), _stream())
    return PackedWeight(out, N, Kl, BN, BK)


def linear_tc(A: torch.Tensor, W: PackedWeight, bias: Optional[torch.Tensor] = None,
              act: str = "none", pool: int = 1, split: Optional[int] = None,
              out: Optional[torch.Tensor] = None, accumulate: bool = False,
              tag: Optional[str] = None) -> torch.Tensor:
    """out[M,N] = act(A[M,K] @ W^T + bias) on the tensor cores; A may be any 2-D view."""
    assert A.is_cuda and A.dtype == torch.float32 and A.dim() == 2
    assert W.BN in (64, 128, 192) and W.BK == 64
    M, K = A.shape
    assert K == W.K, (K, W.K)
    rows = M // pool if pool > 1 else M
    if out is None:
        out = torch.empty(rows, W.N, dtype=torch.float32, device=A.device)
    assert out.stride(1) == 1 and out.shape[0] == rows and out.shape[1] >= W.N
    with torch.cuda.device(A.device), _range(tag):
        _C.call("msr3d_linear_tc_f32", A.data_ptr(), A.stride(0), A.stride(1), M, K,
                W.data.data_ptr(), _ptr(bias), out.data_ptr(), out.stride(0), W.N, ACT[act],
                int(pool), int(split or DEFAULT_SPLIT), int(accumulate), int(W.BN), _stream())
    return out


def linear_chain(A, W: PackedWeight, bias: Optional[torch.Tensor], act: str = "none", pool: int = 1,
                 out: Optional[torch.Tensor] = None, out_split=None, tag: Optional[str] = None):
    """One dense layer of a chain (csrc/linear_tma.cu, split-bf16 hh + hl + lh): `A` is an fp32 (M, K) tensor or a
    (hi, lo) pair of bf16 (M, K) tensors written by the previous layer; the result goes to `out` (fp32, optionally
    max-pooled over `pool` rows) or, with out_split=(hi, lo) bf16 (M, N) tensors, leaves PRE-SPLIT for the next layer."""
    assert W.BN == 128 and W.BK == 64
    if isinstance(A, tuple):
        a_hi, a_lo = A
        assert a_hi.dtype == torch.bfloat16 and a_lo.dtype == torch.bfloat16 and a_hi.stride(1) == 1 and a_hi.shape == a_lo.shape
        M, K = a_hi.shape
        a_args = (None, a_hi.data_ptr(), a_lo.data_ptr(), a_hi.stride(0))
        dev = a_hi.device
    else:
        assert A.is_cuda and A.dtype == torch.float32 and A.dim() == 2 and A.stride(1) == 1
        M, K = A.shape
        a_args = (A.data_ptr(), None, None, A.stride(0))
        dev = A.device
    assert K == W.K, (K, W.K)
    if out_split is not None:
        o_hi, o_lo = out_split
        assert o_hi.dtype == torch.bfloat16 and o_hi.shape == (M, W.N) and o_lo.shape == (M, W.N) and o_hi.stride(0) == o_lo.stride(0)
        o_args = (None, o_hi.data_ptr(), o_lo.data_ptr(), o_hi.stride(0))
    else:
        rows = M // pool if pool > 1 else M
        if out is None:
            out = torch.empty(rows, W.N, dtype=torch.float32, device=dev)
        o_args = (out.data_ptr(), None, None, out.stride(0))
    with torch.cuda.device(dev), _range(tag):
        _C.call("msr3d_linear_bf16x3_f32", a_args[0], a_args[1], a_args[2], a_args[3], M, K, W.data.data_ptr(), _ptr(bias),
                o_args[0], o_args[1], o_args[2], o_args[3], W.N, ACT[act], int(pool), _stream())
    return out_split if out_split is not None else out


def sa_group(xyz_records: torch.Tensor, npoint: int, radius: float, nsample: int, tag: Optional[str] = None):
    """FPS + centre gather + ball query.  xyz_records: (B, N, ld) f32, xyz = first 3 floats."""
    assert xyz_records.is_cuda and xyz_records.dtype == torch.float32 and xyz_records.is_contiguous()
    B, N, ld = xyz_records.shape
    dev = xyz_records.device
    fps_idx = torch.empty(B, npoint, dtype=torch.int32, device=dev)
    new_xyz = torch.empty(B, npoint, 3, dtype=torch.float32, device=dev)
    ball = torch.empty(B, npoint, nsample, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _range(tag):
        _C.call("msr3d_sa_group_f32", xyz_records.data_ptr(), ld, B, N, int(npoint), float(radius),
                int(nsample), fps_idx.data_ptr(), new_xyz.data_ptr(), ball.data_ptr(), _stream())
    return fps_idx, new_xyz, ball


def sa_mlp_fused(feat: torch.Tensor, nfeat: int, xyz: torch.Tensor, new_xyz: torch.Tensor,
                 ball: torch.Tensor, weights, biases, widths, out: torch.Tensor,
                 split: Optional[int] = None, tag: Optional[str] = None) -> torch.Tensor:
    """Gather + 3 x (conv+BN+ReLU) + max-pool; feat (B,Nsrc,ld_feat) and xyz (B,Nsrc,ld_xyz) may
    be views into the same record tensor (last-dim stride 1)."""
    B, npoint, nsample = ball.shape
    Nsrc = feat.shape[1]
    assert feat.stride(2) == 1 and xyz.stride(2) == 1 and feat.stride(0) == Nsrc * feat.stride(1)
    assert out.stride(2) == 1 and out.shape[:2] == (B, npoint)
    with torch.cuda.device(feat.device), _range(tag):
        _C.call("msr3d_sa_mlp_fused_f32", feat.data_ptr(), feat.stride(1), int(nfeat), xyz.data_ptr(),
                xyz.stride(1), new_xyz.data_ptr(), ball.data_ptr(), B, Nsrc, npoint, nsample,
                weights[0].data.data_ptr(), weights[1].data.data_ptr(), weights[2].data.data_ptr(),
                biases[0].data_ptr(), biases[1].data_ptr(), biases[2].data_ptr(),
                int(widths[0]), int(widths[1]), int(widths[2]), out.data_ptr(), out.stride(1),
                int(split or DEFAULT_SPLIT), _stream())
    return out

"""Unfrozen PointNet++ encoder on this repo's kernels.

`PcdObjEncoder(freeze=False)` (modules/vision/pcd_pointnet_encoder.py:58-62) runs `pcd_net` with autograd and
whatever mode its BatchNorm layers are in.  The reference does that with ~70 ATen launches per direction plus the
atomicAdd scatter kernels of pointnet2 (`_ext_src/src/group_points_gpu.cu:43-75`).  Here the whole encoder is ONE
`torch.autograd.Function`:

  forward   per SA level: FPS + centre gather + ball query (bit-exact index kernels, csrc/ext_ops.cu), the grouped
            operand X built row-major in one pass, then per layer  Z = X W^T on the tensor cores
            (msr3d_gemm_tc_f32, split-bf16 tcgen05) -> batch statistics (two-stage deterministic reduction, running
            statistics updated as nn.BatchNorm2d does) -> Y = relu(BN(Z)); max-pool with saved arg-max.
  backward  hand-written: max-pool routing, ReLU + BatchNorm backward (training- or eval-mode), weight gradients
            dW = dZ^T X and data gradients dX = dZ W on the tensor cores, the grouping gradient as an atomic-free
            segmented sum per cloud.  Gradients for all 29 PointNet++ parameters; none for the input points.

Row-major [row][channel] activations are what the GEMM kernels read as K-major operands, so nothing is transposed
and the reference's (B, C, npoint, nsample) tensors never exist (csrc/sa_train.cu).
"""
from __future__ import annotations

import ctypes
from typing import List

import torch

from . import _C, ops
from .fused import gemm_tc

c_int, c_ll, c_float, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p

_C.register_protos({
    "msr3d_sa_train_stat_chunks": [c_ll],
    "msr3d_sa_group_rows_f32": [c_vp, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_int, c_int, c_ll, c_vp, c_int, c_vp],
    "msr3d_bn_train_fwd_f32": [c_vp, c_ll, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_float, c_float, c_vp, c_vp, c_vp, c_vp],
    "msr3d_bn_train_bwd_f32": [c_vp, c_vp, c_vp, c_ll, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_maxpool_rows_f32": [c_vp, c_ll, c_int, c_int, c_vp, c_ll, c_vp, c_vp],
    "msr3d_maxpool_rows_bwd_f32": [c_vp, c_ll, c_vp, c_ll, c_int, c_int, c_vp, c_vp],
    "msr3d_sa_group_rows_bwd_f32": [c_vp, c_int, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
})

BN_EPS, BN_MOMENTUM = 1e-5, 0.1       # nn.BatchNorm2d defaults (pytorch_utils.py:39-58)
# split-bf16 products of the tensor-core GEMMs: 6 = three-way split (fp32-grade), 3 = two-way (1.5e-5 relative).
# The training path defaults to fp32-grade everywhere: max-pool arg-max and ReLU decisions sit on exact ties of the
# forward values, and every decision that flips against an fp32 evaluation moves a whole gradient contribution
# (tools/debug_enc_train.py: with the two-way split ~1e-4 of the decisions flip and single weight-gradient rows move
# by 10 % while all norms agree to 1e-4).  MSR3D_ENC_TRAIN_SPLIT=3 trades that for speed.
FWD_SPLIT = int(__import__("os").environ.get("MSR3D_ENC_TRAIN_SPLIT", "6"))
WGRAD_SPLIT = 6
DGRAD_SPLIT = 3      # data gradients decide nothing discrete: two-way split (1.5e-5) is ample


# GEMM kernel: "tf32x3" = the TMA-fed tcgen05 kind::tf32 big / small kernel of the transformer (csrc/gemm3x.cu, fp32-grade,
# operands fetched by tensor-map TMA straight from the row-major activations; row pitches are padded to multiples of
# 4 floats for it), "bf16" = the register-path split-bf16 kernel of round 1 (csrc/linear_tc.cu; FWD_SPLIT applies)
GEMM_IMPL = __import__("os").environ.get("MSR3D_ENC_TRAIN_GEMM", "tf32x3")


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def _tma_ok(t: torch.Tensor) -> bool:
    if t.data_ptr() % 16:
        return False
    if t.stride(1) == 1:
        return t.stride(0) % 4 == 0 and t.stride(0) >= t.shape[1]
    return t.stride(0) == 1 and t.stride(1) % 4 == 0 and t.stride(1) >= t.shape[0]


def _gemm(A, Bm, split, **kw):
    """out = A @ Bm^T through the configured tensor-core kernel (falls back to the register-path kernel for operands
    the TMA path cannot address)."""
    out = kw.get("out")
    if GEMM_IMPL == "tf32x3" and _tma_ok(A) and _tma_ok(Bm) and (out is None or out.stride(0) % 4 == 0):
        return gemm_tc(A, Bm, split=split, impl="tf32x3", **kw)
    return gemm_tc(A, Bm, split=split, **kw)


def _weight(W2):
    """(cout, cin) weight with the row pitch padded to a multiple of 4 floats (zero columns), matching the padded X."""
    cout, cin = W2.shape
    if cin % 4 == 0:
        return W2
    Wp = torch.zeros(cout, _pad4(cin), dtype=W2.dtype, device=W2.device)
    Wp[:, :cin] = W2
    return Wp


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ksplit_for(k_rows: int, out_tiles: int) -> int:
    """Split-K factor of a weight-gradient GEMM: the contraction runs over all rows (up to 245 760 at config 2) while
    the output is a few tiles — spread the rows over about two waves of CTAs."""
    want = max(1, (2 * ops.NUM_SMS) // max(1, out_tiles))
    return max(1, min(want, k_rows // 512 if k_rows >= 1024 else 1))


def supported(net, pc: torch.Tensor) -> bool:
    """Same structural envelope as the fused frozen path: three SA levels (two QueryAndGroup + GroupAll), xyz + 3
    feature channels, three conv+BN+ReLU blocks per level."""
    if len(net.encoder) != 3 or not net.use_xyz or pc.size(-1) != 6 or not pc.is_cuda:
        return False
    if net.sa_n_points[2] is not None or net.sa_n_points[0] is None or net.sa_n_points[1] is None:
        return False
    for sa in net.encoder:
        mlp = sa.mlps[0]
        if len(sa.mlps) != 1 or any(len(blk) != 3 for blk in mlp):
            return False
        if any(blk[0].out_channels % 4 for blk in mlp):
            return False
    return net.sa_n_points[1] * max(blk[0].out_channels for blk in net.encoder[1].mlps[0]) * 4 <= 200 * 1024 \
        and net.sa_n_points[0] * net.encoder[0].mlps[0][-1][0].out_channels * 4 <= 200 * 1024


class _Level:
    """Saved tensors of one SA level between forward and backward."""
    __slots__ = ("X", "Z", "Y", "stats", "arg", "idx", "rows", "S", "widths", "batch_stats", "n_src", "pairs")


def _mlp_forward(lv: _Level, X, layers, training_bn: List[bool]):
    """Three conv1x1 + BN + ReLU blocks over row-major X; keeps Z, Y, stats per layer."""
    dev = X.device
    rows = X.shape[0]
    lv.X, lv.Z, lv.Y, lv.stats, lv.batch_stats = X, [], [], [], []
    a = X
    for (W, gamma, beta, rmean, rvar), bn_train in zip(layers, training_bn):
        cout = W.shape[0]
        Z = _gemm(a, _weight(W.view(cout, -1)), FWD_SPLIT)
        chunks = _C.lib().msr3d_sa_train_stat_chunks(rows)
        partial = torch.empty(2 * cout * max(1, chunks), dtype=torch.float32, device=dev)
        stats = torch.empty(2, cout, dtype=torch.float32, device=dev)
        Y = torch.empty_like(Z)
        _C.call("msr3d_bn_train_fwd_f32", Z.data_ptr(), rows, cout, gamma.data_ptr(), beta.data_ptr(),
                rmean.data_ptr(), rvar.data_ptr(), int(bn_train), BN_EPS, BN_MOMENTUM, partial.data_ptr(),
                stats.data_ptr(), Y.data_ptr(), _st())
        lv.Z.append(Z); lv.Y.append(Y); lv.stats.append(stats); lv.batch_stats.append(bool(bn_train))
        a = Y
    return a


def _pool(lv: _Level, Y, S):
    groups, C = Y.shape[0] // S, Y.shape[1]
    out = torch.empty(groups, C, dtype=torch.float32, device=Y.device)
    lv.arg = torch.empty(groups, C, dtype=torch.int32, device=Y.device)
    lv.S = S
    _C.call("msr3d_maxpool_rows_f32", Y.data_ptr(), groups, S, C, out.data_ptr(), C, lv.arg.data_ptr(), _st())
    return out


def _sink(sinks, key_tensor, like):
    """Destination of a parameter gradient: the caller's buffer (accumulated into: the engine's flat gradient views) or a
    fresh zero tensor (autograd mode)."""
    if sinks is not None and id(key_tensor) in sinks:
        return sinks[id(key_tensor)]
    return torch.zeros_like(like)


def _mlp_backward(lv: _Level, d_pool, layers, grads, need_dx: bool, sinks=None):
    """d_pool (groups, C_last) -> parameter gradients of the level's three blocks (+ dX if needed)."""
    dev = d_pool.device
    rows = lv.X.shape[0]
    C_last = lv.Y[-1].shape[1]
    dY = torch.empty(rows, C_last, dtype=torch.float32, device=dev)
    _C.call("msr3d_maxpool_rows_bwd_f32", d_pool.data_ptr(), d_pool.stride(0), lv.arg.data_ptr(), rows // lv.S, lv.S,
            C_last, dY.data_ptr(), _st())
    dX = None
    for j in reversed(range(len(layers))):
        W, gamma, beta, _, _ = layers[j]
        cout = W.shape[0]
        W2 = _weight(W.view(cout, -1))          # (cout, cin padded to the pitch of X)
        x_in = lv.X if j == 0 else lv.Y[j - 1]
        chunks = _C.lib().msr3d_sa_train_stat_chunks(rows)
        partial = torch.empty(2 * cout * max(1, chunks), dtype=torch.float32, device=dev)
        bstat = torch.empty(2, cout, dtype=torch.float32, device=dev)
        dgamma, dbeta = _sink(sinks, gamma, gamma), _sink(sinks, beta, beta)
        # ReLU + BatchNorm backward, in place: dY becomes dZ
        _C.call("msr3d_bn_train_bwd_f32", dY.data_ptr(), lv.Y[j].data_ptr(), lv.Z[j].data_ptr(), rows, cout,
                gamma.data_ptr(), lv.stats[j].data_ptr(), int(lv.batch_stats[j]), partial.data_ptr(), bstat.data_ptr(),
                dgamma.data_ptr(), dbeta.data_ptr(), dY.data_ptr(), _st())
        dZ = dY
        # dW (cout, cin) = dZ^T X: contraction over every row -> split-K with red.add into a zeroed buffer
        cin = W2.shape[1]
        cin_real = W.numel() // cout
        dst = _sink(sinks, W, W).view(cout, cin_real)
        dW = dst if cin_real == cin else torch.zeros(cout, cin, dtype=torch.float32, device=dev)
        tiles = -(-cout // 128) * -(-cin // 64)
        _gemm(dZ.t(), x_in.t(), WGRAD_SPLIT, out=dW, accumulate=2, ksplit=_ksplit_for(rows, tiles))
        if dW is not dst:
            dst.add_(dW[:, :cin_real])            # padded pitch (first layer of a level): fold the real columns in
        grads[id(W)] = dst.view_as(W)
        grads[id(gamma)], grads[id(beta)] = dgamma, dbeta
        if j > 0 or need_dx:
            dY = _gemm(dZ, W2.t(), DGRAD_SPLIT)                # dX (rows, cin) = dZ W
            if j == 0:
                dX = dY
    return dX


class _State:
    """Everything the backward needs from one forward (explicit API: the engine keeps it; autograd mode: the ctx)."""
    __slots__ = ("levels", "layer_sets", "g", "dims")


def forward_explicit(net, pc: torch.Tensor):
    """(B', P, 6) -> ((B', C_out) embeddings, state) without autograd bookkeeping."""
    pc = pc.contiguous().float()
    Bc, P, _ = pc.shape
    dev = pc.device
    np1, np2 = net.sa_n_points[0], net.sa_n_points[1]
    S1, S2 = net.sa_n_samples[0], net.sa_n_samples[1]
    layer_sets = []
    for sa in net.encoder:
        ls = []
        for blk in sa.mlps[0]:
            bn = blk[1][0]
            ls.append((blk[0].weight.detach(), bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var))
        layer_sets.append(ls)
    bn_mode = [[blk[1][0].training for blk in sa.mlps[0]] for sa in net.encoder]
    with torch.cuda.device(dev):
        # ---- level 1: neighbours are the raw (x, y, z, r, g, b) records
        _, xyz1, ball1 = ops.sa_group(pc, np1, net.sa_radii[0], S1)
        l1 = _Level()
        rows1 = Bc * np1 * S1
        X1 = torch.empty(rows1, _pad4(6), dtype=torch.float32, device=dev)
        _C.call("msr3d_sa_group_rows_f32", pc.data_ptr(), 6, pc.data_ptr() + 12, 6, 3, xyz1.data_ptr(),
                ball1.data_ptr(), P, np1, S1, rows1, X1.data_ptr(), X1.shape[1], _st())
        f1 = _pool(l1, _mlp_forward(l1, X1, layer_sets[0], bn_mode[0]), S1)          # (Bc * np1, C1)
        C1 = f1.shape[1]
        # ---- level 2
        _, xyz2, ball2 = ops.sa_group(xyz1, np2, net.sa_radii[1], S2)
        l2 = _Level()
        rows2 = Bc * np2 * S2
        X2 = torch.empty(rows2, _pad4(3 + C1), dtype=torch.float32, device=dev)
        _C.call("msr3d_sa_group_rows_f32", xyz1.data_ptr(), 3, f1.data_ptr(), C1, C1, xyz2.data_ptr(),
                ball2.data_ptr(), np1, np2, S2, rows2, X2.data_ptr(), X2.shape[1], _st())
        l2.idx, l2.n_src, l2.pairs = ball2, np1, np2 * S2
        f2 = _pool(l2, _mlp_forward(l2, X2, layer_sets[1], bn_mode[1]), S2)          # (Bc * np2, C2)
        C2 = f2.shape[1]
        # ---- level 3 (GroupAll): rows = the np2 remaining points of every cloud, raw xyz
        l3 = _Level()
        rows3 = Bc * np2
        X3 = torch.empty(rows3, _pad4(3 + C2), dtype=torch.float32, device=dev)
        _C.call("msr3d_sa_group_rows_f32", xyz2.data_ptr(), 3, f2.data_ptr(), C2, C2, None, None, np2, 1, 1, rows3,
                X3.data_ptr(), X3.shape[1], _st())
        g = _pool(l3, _mlp_forward(l3, X3, layer_sets[2], bn_mode[2]), np2)          # (Bc, C3)
        out = _gemm(g, net.fc.weight.detach(), FWD_SPLIT, bias=net.fc.bias.detach())
    for sa, modes in zip(net.encoder, bn_mode):
        for blk, tr in zip(sa.mlps[0], modes):
            if tr:
                blk[1][0].num_batches_tracked += 1
    st = _State()
    st.levels, st.layer_sets, st.g, st.dims = (l1, l2, l3), layer_sets, g, (Bc, np1, np2, C1, C2)
    return out, st


def backward_explicit(net, st: _State, d_out: torch.Tensor, sinks=None) -> dict:
    """Gradients of the 29 PointNet++ parameters from d(embeddings): {id(parameter): gradient}.  `sinks` maps
    id(parameter) to a tensor the gradient is ACCUMULATED into (the engine's views of its flat gradient buffer, already
    zeroed); without it every gradient is a fresh tensor."""
    (l1, l2, l3), layer_sets, g = st.levels, st.layer_sets, st.g
    Bc, np1, _, C1, C2 = st.dims
    d_out = d_out.contiguous().float()
    dev = d_out.device
    grads = {}
    by_detached = {}        # sinks are keyed by the module's parameters; the layer sets hold detached aliases
    if sinks is not None:
        for sa, ls in zip(net.encoder, layer_sets):
            for blk, (W, gamma, beta, _, _) in zip(sa.mlps[0], ls):
                bn = blk[1][0]
                for alias, p in ((W, blk[0].weight), (gamma, bn.weight), (beta, bn.bias)):
                    if id(p) in sinks:
                        by_detached[id(alias)] = sinks[id(p)]
        for p in (net.fc.weight, net.fc.bias):
            if id(p) in sinks:
                by_detached[id(p)] = sinks[id(p)]
    sk = by_detached if sinks is not None else None
    with torch.cuda.device(dev):
        Wfc = net.fc.weight.detach()
        dWfc = _sink(sk, net.fc.weight, Wfc)
        _gemm(d_out.t(), g.t(), WGRAD_SPLIT, out=dWfc, accumulate=2, ksplit=1)
        grads[id(net.fc.weight)] = dWfc
        dbfc = _sink(sk, net.fc.bias, net.fc.bias)
        _C.call("msr3d_colsum_add_f32", d_out.data_ptr(), d_out.stride(0), d_out.shape[0], d_out.shape[1],
                dbfc.data_ptr(), None, None, _st())
        grads[id(net.fc.bias)] = dbfc
        dg = _gemm(d_out, Wfc.t(), DGRAD_SPLIT)                                        # (Bc, C3)
        dX3 = _mlp_backward(l3, dg, layer_sets[2], grads, True, sk)                    # (Bc * np2, 3 + C2)
        d_f2 = dX3[:, 3:3 + C2]                                                        # GroupAll: rows = points
        dX2 = _mlp_backward(l2, d_f2, layer_sets[1], grads, True, sk)                  # (Bc * np2 * S2, 3 + C1)
        d_f1 = torch.empty(Bc * np1, C1, dtype=torch.float32, device=dev)
        _C.call("msr3d_sa_group_rows_bwd_f32", dX2.data_ptr(), dX2.shape[1], 3, l2.idx.data_ptr(), Bc, l2.pairs,
                l2.n_src, C1, d_f1.data_ptr(), _st())
        _mlp_backward(l1, d_f1, layer_sets[0], grads, False, sk)                       # inputs need no gradient
    out = {}
    for sa, ls in zip(net.encoder, layer_sets):
        for blk, (W, gamma, beta, _, _) in zip(sa.mlps[0], ls):
            bn = blk[1][0]
            out[id(blk[0].weight)] = grads[id(W)]
            out[id(bn.weight)] = grads[id(gamma)]
            out[id(bn.bias)] = grads[id(beta)]
    out[id(net.fc.weight)] = grads[id(net.fc.weight)]
    out[id(net.fc.bias)] = grads[id(net.fc.bias)]
    return out


class _PointNetPPTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, pc, *params):
        out, st = forward_explicit(net, pc)
        ctx.net, ctx.st, ctx.params = net, st, params
        return out

    @staticmethod
    def backward(ctx, d_out):
        grads = backward_explicit(ctx.net, ctx.st, d_out)
        return (None, None) + tuple(grads.get(id(p)) for p in ctx.params)


def forward_train(net, pc: torch.Tensor) -> torch.Tensor:
    """(B', P, 6) -> (B', C_out) with autograd through this repo's kernels (see module docstring)."""
    params = [p for p in net.parameters()]
    return _PointNetPPTrainFn.apply(net, pc, *params)

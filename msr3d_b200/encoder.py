"""PointNet++ object encoder of MSR3D, B200-native.

Drop-in for the reference classes (same names, constructor arguments, forward
signatures and state-dict keys):

  SharedMLP / Conv2d / BatchNorm2d   modules/third_party/pointnet2/pytorch_utils.py:11-58,157-188
  QueryAndGroup / GroupAll           modules/third_party/pointnet2/pointnet2_utils.py:291-419
  PointnetSAModule(MSG)              modules/third_party/pointnet2/pointnet2_modules.py:26-161
  PointNetPP                         modules/layers/pointnet.py:22-63
  PcdObjEncoder                      modules/vision/pcd_pointnet_encoder.py:10-74

Two execution paths, both on this repo's CUDA kernels:

* fused (what every shipped config runs: encoder frozen, BatchNorm in eval,
  no_grad — pcd_pointnet_encoder.py:42-57): per SA level one front-end call
  (FPS + centre gather + ball query) and ONE tensor-core kernel (gather + 3 x
  [conv1x1+BN+ReLU] + max-pool); GroupAll level and `fc` on the tcgen05 GEMM.
  Tensors between levels are point-major, BatchNorm is folded into the packed
  weights, the grouped (B,C,npoint,nsample) tensors never exist.
* training (unfrozen encoder, pcd_pointnet_encoder.py:58-62): msr3d_b200/encoder_train.py — one autograd
  Function over row-major kernels (training-mode BatchNorm with batch statistics, tensor-core GEMMs
  for the convolutions and every gradient, atomic-free grouping gradient).
* layer-wise (odd hyper-parameters, A/B): the module tree below, built from the `_ext` operators of
  msr3d_b200.ext with autograd, and torch for conv/BN.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import ext, ops


# ---------------------------------------------------------------------------
# autograd wrappers of the _ext operators (pointnet2_utils.py:48-288)
# ---------------------------------------------------------------------------
class _FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        idx = ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


class _Gather(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.saved = (idx, features.size(2))
        return ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.saved
        return ext.gather_points_grad(grad_out.contiguous(), idx, n), None


class _Grouping(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.saved = (idx, features.size(2))
        return ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.saved
        return ext.group_points_grad(grad_out.contiguous(), idx, n), None


class _BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        idx = ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad=None):
        return None, None, None, None


class _ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = ext.three_nn(unknown, known)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


class _ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.saved = (idx, weight, features.size(2))
        return ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.saved
        return ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


furthest_point_sample = _FurthestPointSampling.apply
gather_operation = _Gather.apply
grouping_operation = _Grouping.apply
ball_query = _BallQuery.apply
three_nn = _ThreeNN.apply
three_interpolate = _ThreeInterpolate.apply


class QueryAndGroup(nn.Module):
    """Ball query + grouping (+ centre subtraction, xyz channels first)."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        if sample_uniformly or ret_unique_cnt:
            raise NotImplementedError("sample_uniformly is a VoteNet option MSR3D never enables")
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz, self.normalize_xyz = ret_grouped_xyz, normalize_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz = grouped_xyz / self.radius
        if features is not None:
            grouped = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features


class GroupAll(nn.Module):
    """All points form one group; xyz is NOT centred (pointnet2_utils.py:407)."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped


# ---------------------------------------------------------------------------
# SharedMLP with the reference's parameter names
# ---------------------------------------------------------------------------
class BatchNorm2d(nn.Sequential):
    def __init__(self, in_size: int, name: str = ""):
        super().__init__()
        self.add_module(name + "bn", nn.BatchNorm2d(in_size))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0)


class Conv2d(nn.Sequential):
    """1x1 conv (no bias when followed by BN) -> BN -> activation, keys conv / bn.bn / activation."""

    def __init__(self, in_size, out_size, *, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_,
                 bias=True, preact=False, name=""):
        super().__init__()
        if preact:
            raise NotImplementedError("preact SharedMLPs are not used on the MSR3D path")
        conv = nn.Conv2d(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding,
                         bias=bias and not bn)
        init(conv.weight)
        if conv.bias is not None:
            nn.init.constant_(conv.bias, 0)
        self.add_module(name + "conv", conv)
        if bn:
            self.add_module(name + "bn", BatchNorm2d(out_size))
        if activation is not None:
            self.add_module(name + "activation", activation)


class SharedMLP(nn.Sequential):
    def __init__(self, args: List[int], *, bn=False, activation=nn.ReLU(inplace=True), preact=False,
                 first=False, name=""):
        super().__init__()
        for i in range(len(args) - 1):
            self.add_module(name + f"layer{i}", Conv2d(args[i], args[i + 1], bn=bn, activation=activation))


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint, self.groupers, self.mlps = None, None, None

    def forward(self, xyz, features=None):
        """Layer-wise path.  xyz (B,N,3), features (B,C,N) -> new_xyz (B,npoint,3), (B,C',npoint)."""
        new_xyz = None
        if self.npoint is not None:
            idx = furthest_point_sample(xyz, self.npoint)
            new_xyz = gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        outs = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            h = mlp(grouper(xyz, new_xyz, features))
            outs.append(h.max(dim=3).values)
        return new_xyz, torch.cat(outs, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    def __init__(self, *, npoint, radii, nsamples, mlps, bn=True, use_xyz=True, sample_uniformly=False):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.radii, self.nsamples = list(radii), list(nsamples)
        self.groupers, self.mlps = nn.ModuleList(), nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            self.groupers.append(QueryAndGroup(radius, nsample, use_xyz=use_xyz)
                                 if npoint is not None else GroupAll(use_xyz))
            if use_xyz:
                spec[0] += 3  # the reference mutates the caller's list too (pointnet2_modules.py:121-122)
            self.mlps.append(SharedMLP(spec, bn=bn))


class PointnetSAModule(PointnetSAModuleMSG):
    def __init__(self, *, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz)


# ---------------------------------------------------------------------------
# BatchNorm folding + weight packing for the tensor-core kernels
# ---------------------------------------------------------------------------
def _bn_fold(sd, p):
    """eval-mode BatchNorm as y = s*x + t (pytorch_utils.py:39-58; eps 1e-5)."""
    if f"{p}.bn.bn.weight" not in sd:
        w = sd[f"{p}.conv.weight"]
        bias = sd.get(f"{p}.conv.bias")
        return None, (bias if bias is not None else torch.zeros(w.shape[0], device=w.device))
    s = sd[f"{p}.bn.bn.weight"] / torch.sqrt(sd[f"{p}.bn.bn.running_var"] + 1e-5)
    t = sd[f"{p}.bn.bn.bias"] - sd[f"{p}.bn.bn.running_mean"] * s
    return s.contiguous(), t.contiguous()


def _first_layer_perm(nfeat: int, kp: int, device) -> torch.Tensor:
    """Packed K order [feat(nfeat) | xyz(3) | 0...] -> reference input channels [xyz(3) | feat]."""
    perm = torch.full((kp,), -1, dtype=torch.int32)
    perm[:nfeat] = torch.arange(3, 3 + nfeat, dtype=torch.int32)
    perm[nfeat:nfeat + 3] = torch.arange(0, 3, dtype=torch.int32)
    return perm.to(device)


def fold_sa_level(sd, prefix: str, nfeat: int, fused: bool, bn_last: int = 128, bn_first: int = 128):
    """Packs the three conv+BN layers under `prefix` (…mlps.0) for the tensor cores.

    fused=True : tiles for sa_mlp_fused (BN = min(C,128), BK = whole padded K of the layer)
    fused=False: tiles for linear_tc    (BN = 128, BK = 64; `bn_last` = N tile of the last layer)
    Returns ([PackedWeight]*3, [bias]*3)."""
    weights, biases = [], []
    i = 0
    while f"{prefix}.layer{i}.conv.weight" in sd:
        p = f"{prefix}.layer{i}"
        w = sd[f"{p}.conv.weight"]
        w2 = w.reshape(w.shape[0], w.shape[1])
        s, t = _bn_fold(sd, p)
        n, k = w2.shape
        if fused:
            bn_, bk = min(n, 128), -(-k // 16) * 16
        else:
            bn_, bk = 128, 64
            if bn_last != 128 and f"{prefix}.layer{i + 1}.conv.weight" not in sd and n % bn_last == 0:
                bn_ = bn_last
            if i == 0 and bn_first != 128:
                bn_ = bn_first
        perm = None
        if i == 0:
            kp = -(-k // bk) * bk
            perm = _first_layer_perm(nfeat, kp, w.device)
        weights.append(ops.pack_weight(w2, row_scale=s, col_perm=perm, BN=bn_, BK=bk, K=k))
        biases.append(t.float().contiguous())
        i += 1
    return weights, biases


_FUSED_WIDTHS = {(6, 64, 64, 128), (131, 128, 128, 256)}


class PointNetPP(nn.Module):
    """PointNet++ encoder: SA levels + fc (modules/layers/pointnet.py:22-63)."""

    def __init__(self, sa_n_points: list, sa_n_samples: list, sa_radii: list, sa_mlps: list,
                 bn=True, use_xyz=True):
        super().__init__()
        n_sa = len(sa_n_points)
        if not (n_sa == len(sa_n_samples) == len(sa_radii) == len(sa_mlps)):
            raise ValueError('Lens of given hyper-params are not compatible')
        self.sa_n_points, self.sa_n_samples, self.sa_radii = list(sa_n_points), list(sa_n_samples), list(sa_radii)
        self.use_xyz = use_xyz
        self.encoder = nn.ModuleList()
        for i in range(n_sa):
            self.encoder.append(PointnetSAModule(npoint=sa_n_points[i], nsample=sa_n_samples[i],
                                                 radius=sa_radii[i], mlp=sa_mlps[i], bn=bn, use_xyz=use_xyz))
        out_n_points = sa_n_points[-1] if sa_n_points[-1] is not None else 1
        self.fc = nn.Linear(out_n_points * sa_mlps[-1][-1], sa_mlps[-1][-1])
        self._pack_key = None
        self._packed = None
        self._ws = {}
        # False: the unfrozen encoder runs the layer-wise module tree (torch conv / BN over the _ext operators)
        self.train_kernels = __import__("os").environ.get("MSR3D_ENCODER_TRAIN_KERNELS", "1") == "1"

    # ---- layer-wise path (autograd, training-mode BN allowed)
    def forward_layerwise(self, pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        for sa in self.encoder:
            xyz, features = sa(xyz, features)
        return self.fc(features.reshape(features.size(0), -1))

    # ---- fused path
    def _widths(self, i):
        mlp = self.encoder[i].mlps[0]
        ws = [mlp[0][0].in_channels] + [blk[0].out_channels for blk in mlp]
        return tuple(ws)

    def fused_supported(self, pc) -> bool:
        if len(self.encoder) != 3 or not self.use_xyz or pc.size(-1) != 6:
            return False
        if self.sa_n_points[2] is not None or self.sa_n_points[0] is None or self.sa_n_points[1] is None:
            return False
        for i in (0, 1):
            if self.sa_n_samples[i] != 32 or self.sa_n_points[i] % 4 != 0 or self._widths(i) not in _FUSED_WIDTHS:
                return False
            if any(len(blk) != 3 for blk in self.encoder[i].mlps[0]):  # conv, bn, relu
                return False
        if self._widths(0)[0] != 6 or self._widths(1)[0] != 131 or len(self.encoder[2].mlps[0]) != 3:
            return False
        if self.sa_n_points[1] not in (1, 2, 4, 8, 16, 32):   # GroupAll pool group must divide a warp
            return False
        bns = [m for m in self.modules() if isinstance(m, nn.BatchNorm2d)]
        return all(not m.training for m in bns)

    def _packed_weights_key(self):
        params = self.__dict__.get("_pk_tensors")
        if params is None or len(params) != self.__dict__.get("_pk_count", -1):
            params = list(self.parameters()) + list(self.buffers())
            self.__dict__["_pk_tensors"], self.__dict__["_pk_count"] = params, len(params)
        return tuple((p.data_ptr(), p._version) for p in params)

    def _packed_weights(self):
        key = self._packed_weights_key()
        if key != self._pack_key:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            packed = {}
            packed["sa1"] = fold_sa_level(sd, "encoder.0.mlps.0", 3, fused=True)
            packed["sa2"] = fold_sa_level(sd, "encoder.1.mlps.0", self._widths(0)[-1], fused=True)
            packed["sa3"] = fold_sa_level(sd, "encoder.2.mlps.0", self._widths(1)[-1], fused=False)
            # last SA3 layer a second time with 192-wide N tiles: 3840 x 768 is 180 tiles of 128 x 128
            # (1.2 waves on 148 SMs) but 120 tiles of 128 x 192 (one wave) — chosen per call from M
            packed["sa3_w192"] = fold_sa_level(sd, "encoder.2.mlps.0", self._widths(1)[-1], fused=False,
                                               bn_last=192)[0][2]
            packed["fc"] = (ops.pack_weight(sd["fc.weight"], BN=64 if ops.ENC_SMALL_TILES else 128),
                            sd["fc.bias"].float().contiguous())
            if ops.ENC_SMALL_TILES:
                packed["sa3"][0][0] = fold_sa_level(sd, "encoder.2.mlps.0", self._widths(1)[-1], fused=False,
                                                    bn_first=64)[0][0]
            self._packed, self._pack_key = packed, key
        return self._packed

    def _workspace(self, name, shape, device, zero=False):
        t = self._ws.get(name)
        if t is None or t.shape != torch.Size(shape) or t.device != device:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=device)
            self._ws[name] = t
        return t

    def _workspace16(self, name, shape, device):
        t = self._ws.get(name)
        if t is None or t.shape != torch.Size(shape) or t.device != device:
            t = torch.empty(shape, dtype=torch.bfloat16, device=device)
            self._ws[name] = t
        return t

    @torch.no_grad()
    def forward_fused(self, pc, split: Optional[int] = None, trace: Optional[dict] = None,
                      out: Optional[torch.Tensor] = None, aux_stream: Optional[torch.cuda.Stream] = None):
        """pc: (B', P, 6) contiguous CUDA f32 -> (B', C_out).  11 kernel launches.
        aux_stream: the grouping of level 2 (FPS + ball query on the 32 centres of level 1) depends on the centres only,
        not on level 1's MLP — it is issued there, beside the SA1 kernel, and joined in front of SA2."""
        pk = self._packed_weights()
        Bc, P, _ = pc.shape
        dev = pc.device
        np1, np2 = self.sa_n_points[0], self.sa_n_points[1]
        w1, w2, w3 = self._widths(0), self._widths(1), self._widths(2)
        # level 1: neighbours are the raw (x,y,z,r,g,b) records
        fps1, xyz1, ball1 = ops.sa_group(pc, np1, self.sa_radii[0], 32, tag="sa1_group")
        f1 = self._workspace("f1", (Bc, np1, w1[-1]), dev)
        if aux_stream is not None:
            main = torch.cuda.current_stream()
            aux_stream.wait_stream(main)
            with torch.cuda.stream(aux_stream):
                fps2, xyz2, ball2 = ops.sa_group(xyz1, np2, self.sa_radii[1], 32, tag="sa2_group")
            for t in (fps2, xyz2, ball2):
                t.record_stream(main)
        ops.sa_mlp_fused(pc[..., 3:], 3, pc, xyz1, ball1, pk["sa1"][0], pk["sa1"][1], w1[1:], f1, split, tag="sa1_mlp")
        # level 2
        if aux_stream is not None:
            torch.cuda.current_stream().wait_stream(aux_stream)
        else:
            fps2, xyz2, ball2 = ops.sa_group(xyz1, np2, self.sa_radii[1], 32, tag="sa2_group")
        k3 = w2[-1] + 3
        ld3 = -(-k3 // 16) * 16
        x3 = self._workspace("x3", (Bc, np2, ld3), dev, zero=True)   # [feat | xyz | 0]: GroupAll operand
        ops.sa_mlp_fused(f1, w1[-1], xyz1, xyz2, ball2, pk["sa2"][0], pk["sa2"][1], w2[1:], x3, split, tag="sa2_mlp")
        from . import _C
        _C.call("msr3d_scatter_xyz_cols_f32", xyz2.data_ptr(), x3.data_ptr(), ld3, w2[-1], Bc * np2,
                torch.cuda.current_stream().cuda_stream)
        # level 3 (GroupAll): three GEMMs, max over the np2 points of a cloud in the last epilogue
        a = x3.view(Bc * np2, ld3)[:, :k3]
        chain = (ops.SA3_CHAIN and (split or ops.DEFAULT_SPLIT) == 3 and all(w.BN == 128 for w in pk["sa3"][0])
                 and w3[1] % 16 == 0 and w3[2] % 16 == 0)
        if chain:
            # the two inner activations leave their producer as bf16 hi / lo images: the consumer's TMA boxes land as
            # operand tiles, no CTA converts anything (csrc/linear_tma.cu)
            M3 = Bc * np2
            h1s = (self._workspace16("h1_hi", (M3, w3[1]), dev), self._workspace16("h1_lo", (M3, w3[1]), dev))
            h2s = (self._workspace16("h2_hi", (M3, w3[2]), dev), self._workspace16("h2_lo", (M3, w3[2]), dev))
            ops.linear_chain(a, pk["sa3"][0][0], pk["sa3"][1][0], act="relu", out_split=h1s, tag="sa3_l1")
            ops.linear_chain(h1s, pk["sa3"][0][1], pk["sa3"][1][1], act="relu", out_split=h2s, tag="sa3_l2")
            g = ops.linear_chain(h2s, pk["sa3"][0][2], pk["sa3"][1][2], act="relu", pool=np2,
                                 out=self._workspace("g", (Bc, w3[3]), dev), tag="sa3_l3")
        else:
            h1 = ops.linear_tc(a, pk["sa3"][0][0], pk["sa3"][1][0], act="relu", split=split,
                               out=self._workspace("h1", (Bc * np2, w3[1]), dev), tag="sa3_l1")
            h2 = ops.linear_tc(h1, pk["sa3"][0][1], pk["sa3"][1][1], act="relu", split=split,
                               out=self._workspace("h2", (Bc * np2, w3[2]), dev), tag="sa3_l2")
            w_l3 = pk["sa3"][0][2]
            mt = -(-(Bc * np2) // 128)
            if (ops.SA3_WIDE_TILES and mt * -(-w3[3] // 128) > ops.NUM_SMS and pk["sa3_w192"].BN == 192
                    and mt * (w3[3] // 192) <= ops.NUM_SMS):
                w_l3 = pk["sa3_w192"]
            g = ops.linear_tc(h2, w_l3, pk["sa3"][1][2], act="relu", pool=np2, split=split,
                              out=self._workspace("g", (Bc, w3[3]), dev), tag="sa3_l3")
        out = ops.linear_tc(g, pk["fc"][0], pk["fc"][1], split=split, tag="fc", out=out)
        if trace is not None:
            trace.update(fps1=fps1, ball1=ball1, xyz1=xyz1, f1=f1, fps2=fps2, ball2=ball2, xyz2=xyz2,
                         x3=x3, g=g)
        return out

    def forward(self, features):
        """features: (B', P, 3 + C) -> (B', C_out)."""
        frozen = not torch.is_grad_enabled() or not any(p.requires_grad for p in self.parameters())
        if features.is_cuda and frozen and self.fused_supported(features):
            return self.forward_fused(features.contiguous().float())
        if features.is_cuda and not frozen and self.train_kernels:
            # unfrozen encoder (pcd_pointnet_encoder.py:58-62): training- or eval-mode BatchNorm with autograd, forward
            # and hand-written backward on this repo's kernels (msr3d_b200/encoder_train.py)
            from . import encoder_train
            if encoder_train.supported(self, features):
                return encoder_train.forward_train(self, features)
        return self.forward_layerwise(features)


def get_mlp_head(input_size, hidden_size, output_size, dropout=0.0):
    """modules/utils.py:25-32."""
    return nn.Sequential(nn.Linear(input_size, hidden_size), nn.ReLU(), nn.LayerNorm(hidden_size, eps=1e-12),
                         nn.Dropout(dropout), nn.Linear(hidden_size, output_size))


class PcdObjEncoder(nn.Module):
    """Registry plugin `PcdObjEncoder` (VISION_REGISTRY): per-object PointNet++ embedding plus the
    607-way semantic head whose output the MSR3D prompter discards."""

    def __init__(self, cfg=None, sa_n_points=[32, 16, None], sa_n_samples=[32, 32, None],
                 sa_radii=[0.2, 0.4, None],
                 sa_mlps=[[3, 64, 64, 128], [128, 128, 128, 256], [256, 256, 512, 768]],
                 dropout=0.1, path=None, freeze=False):
        super().__init__()
        sa_mlps = [list(m) for m in sa_mlps]  # keep the caller's config intact (the `+= 3` mutation stays local)
        self.pcd_net = PointNetPP(sa_n_points=list(sa_n_points), sa_n_samples=list(sa_n_samples),
                                  sa_radii=list(sa_radii), sa_mlps=sa_mlps)
        self.obj3d_clf_pre_head = get_mlp_head(sa_mlps[-1][-1], 384, 607, dropout=0.3)
        self.dropout = nn.Dropout(dropout)
        if path is not None:
            self.load_state_dict(torch.load(path), strict=False)
        self.freeze = freeze
        if freeze:
            for p in self.parameters():
                p.requires_grad = False

    def freeze_bn(self, m):
        for layer in m.modules():
            if isinstance(layer, nn.BatchNorm2d):
                layer.eval()

    def encode(self, obj_pcds):
        """(B,O,P,6) -> obj_embeds (B,O,C); the part of forward() the prompter consumes."""
        B, O = obj_pcds.shape[:2]
        flat = obj_pcds.reshape(B * O, obj_pcds.shape[2], obj_pcds.shape[3])
        if self.freeze:
            self.freeze_bn(self.pcd_net)
            with torch.no_grad():
                emb = self.pcd_net(flat)
            return emb.view(B, O, -1).detach()
        return self.pcd_net(flat).view(B, O, -1)

    def forward(self, obj_pcds, obj_locs=None, obj_masks=None, obj_sem_masks=None, **kwargs):
        obj_embeds = self.encode(obj_pcds)
        return obj_embeds, self.obj3d_clf_pre_head(obj_embeds)

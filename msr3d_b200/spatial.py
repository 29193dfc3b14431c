"""Situated spatial-attention transformer of MSR3D, B200-native.

Drop-in for (same class names, ctor arguments, forward signatures, state-dict keys)

  MultiHeadAttentionSpatial        modules/layers/transformers.py:167-252
  TransformerSpatialEncoderLayer   modules/layers/transformers.py:298-329
  TransformerEncoderLayer          modules/layers/transformers.py:125-164
  CrossAttentionLayer              modules/layers/transformers.py:22-73
  calc_pairwise_locs, quaternion_to_matrix, transform_to_agent_coor
                                   modules/utils.py:60-137
  generate_fourier_features        model/ose3d_situation.py:31-59

Behaviour kept on purpose: the attention block returns
LayerNorm(out + residual) and the encoder layer adds that to its input again before
norm1 (double residual); masked keys get attn = -inf and loc = 0 -> log(1e-6).
Dropped on purpose: the per-layer host sync `assert torch.sum(torch.isnan(...))`
(transformers.py:246), which only fires when *every* element is NaN.
"""
from __future__ import annotations

import copy
import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def get_activation_fn(name: str):
    if name not in ("relu", "gelu", "glu"):
        raise RuntimeError(f"activation function currently support relu/gelu, not {name}")
    return getattr(F, name)


def layer_repeat(module: nn.Module, n: int) -> nn.ModuleList:
    """N-1 deep copies followed by the original (modules/utils.py:35-36)."""
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n - 1)] + [module])


def init_weights_bert(module: nn.Module, std: float = 0.02) -> None:
    """modules/weights.py:3-19."""
    if isinstance(module, nn.Linear):
        module.weight.data.normal_(mean=0.0, std=std)
        if module.bias is not None:
            module.bias.data.zero_()
    elif isinstance(module, nn.Embedding):
        module.weight.data.normal_(mean=0.0, std=std)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)


# ---------------------------------------------------------------------------
# geometry helpers
# ---------------------------------------------------------------------------
def calc_pairwise_locs(obj_centers, obj_whls=None, eps=1e-10, pairwise_rel_type="center",
                       spatial_dist_norm=True, spatial_dim=5):
    """(B,L,3) centres -> (B,L,L,spatial_dim) relative geometry [d/max d, dz/d, d2/d, dy/d2, dx/d2]."""
    if pairwise_rel_type == "mlp":
        locs = torch.cat([obj_centers, obj_whls], 2)
        L = locs.size(1)
        return torch.cat([locs.unsqueeze(2).expand(-1, -1, L, -1), locs.unsqueeze(1).expand(-1, L, -1, -1)], 3)
    delta = obj_centers.unsqueeze(2) - obj_centers.unsqueeze(1)
    dist = torch.sqrt(delta.pow(2).sum(3) + eps)
    norm = dist / dist.flatten(1).max(1).values.view(-1, 1, 1) if spatial_dist_norm else dist
    if spatial_dim == 1:
        return norm.unsqueeze(3)
    dist2d = torch.sqrt(delta[..., :2].pow(2).sum(3) + eps)
    if pairwise_rel_type == "center":
        out = torch.stack([norm, delta[..., 2] / dist, dist2d / dist, delta[..., 1] / dist2d,
                           delta[..., 0] / dist2d], 3)
    elif pairwise_rel_type == "vertical_bottom":
        bottom = obj_centers.clone()
        bottom[:, :, 2] -= obj_whls[:, :, 2]
        bdelta = bottom.unsqueeze(2) - bottom.unsqueeze(1)
        bdist = torch.sqrt(bdelta.pow(2).sum(3) + eps)
        bdist2d = torch.sqrt(bdelta[..., :2].pow(2).sum(3) + eps)
        out = torch.stack([norm, bdelta[..., 2] / bdist, bdist2d / bdist, delta[..., 1] / dist2d,
                           delta[..., 0] / dist2d], 3)
    else:
        raise NotImplementedError(pairwise_rel_type)
    return out[..., 1:] if spatial_dim == 4 else out


def quaternion_to_matrix(q):
    """Rotation of the conjugate quaternion (x,y,z negated): world -> agent frame."""
    x, y, z, w = -q[:, 0], -q[:, 1], -q[:, 2], q[:, 3]
    xx, yy, zz = x * x, y * y, z * z
    xy, xz, xw, yz, yw, zw = x * y, x * z, x * w, y * z, y * w, z * w
    return torch.stack([
        torch.stack([1 - 2 * (yy + zz), 2 * (xy + zw), 2 * (xz - yw)], -1),
        torch.stack([2 * (xy - zw), 1 - 2 * (xx + zz), 2 * (yz + xw)], -1),
        torch.stack([2 * (xz + yw), 2 * (yz - xw), 1 - 2 * (xx + yy)], -1)], -2)


def transform_to_agent_coor(obj_centers, anchor_loc, anchor_ori):
    return torch.matmul(obj_centers - anchor_loc.unsqueeze(1), quaternion_to_matrix(anchor_ori))


def generate_fourier_features(pos, num_bands=10, max_freq=15, concat_pos=True, sine_only=False):
    """(B,N,C) -> (B,N,C*(2*bands)+C): [pos | sin(pi pos f) | cos(pi pos f)], bands per dim contiguous."""
    freqs = torch.linspace(1.0, max_freq, num_bands, device=pos.device)
    scaled = (pos.unsqueeze(-1) * freqs).flatten(2)
    feats = torch.sin(math.pi * scaled) if sine_only else \
        torch.cat([torch.sin(math.pi * scaled), torch.cos(math.pi * scaled)], -1)
    return torch.cat([pos, feats], -1) if concat_pos else feats


# ---------------------------------------------------------------------------
# attention layers
# ---------------------------------------------------------------------------
class MultiHeadAttentionSpatial(nn.Module):
    def __init__(self, d_model, n_head, dropout=0.1, spatial_multihead=True, spatial_dim=5,
                 spatial_attn_fusion="mul"):
        super().__init__()
        assert d_model % n_head == 0, "d_model: %d, n_head: %d" % (d_model, n_head)
        self.n_head, self.d_model, self.d_per_head = n_head, d_model, d_model // n_head
        self.spatial_multihead, self.spatial_dim = spatial_multihead, spatial_dim
        self.spatial_attn_fusion = spatial_attn_fusion
        self.w_qs = nn.Linear(d_model, d_model)
        self.w_ks = nn.Linear(d_model, d_model)
        self.w_vs = nn.Linear(d_model, d_model)
        self.fc = nn.Linear(d_model, d_model)
        self.dropout = nn.Dropout(p=dropout)
        self.layer_norm = nn.LayerNorm(d_model)
        self.spatial_n_head = n_head if spatial_multihead else 1
        if spatial_attn_fusion in ("mul", "bias", "add"):
            self.pairwise_loc_fc = nn.Linear(spatial_dim, self.spatial_n_head)
        elif spatial_attn_fusion == "ctx":
            self.pairwise_loc_fc = nn.Linear(spatial_dim, d_model)
        elif spatial_attn_fusion == "cond":
            self.lang_cond_fc = nn.Linear(d_model, self.spatial_n_head * (spatial_dim + 1))
        else:
            raise NotImplementedError("unsupported spatial_attn_fusion %s" % spatial_attn_fusion)

    def forward(self, q, k, v, pairwise_locs, key_padding_mask=None, txt_embeds=None):
        residual = q
        B, L, D = q.shape
        T = k.shape[1]
        H, dh = self.n_head, self.d_per_head
        qh = self.w_qs(q).view(B, L, H, dh).permute(2, 0, 1, 3)
        kh = self.w_ks(k).view(B, T, H, dh).permute(2, 0, 1, 3)
        vh = self.w_vs(v).view(B, T, H, dh).permute(2, 0, 1, 3)
        attn = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(dh)
        fusion = self.spatial_attn_fusion
        if fusion in ("mul", "bias", "add"):
            loc = self.pairwise_loc_fc(pairwise_locs).permute(3, 0, 1, 2)
            if fusion == "mul":
                loc = F.relu(loc)
            if not self.spatial_multihead:
                loc = loc.expand(H, -1, -1, -1)
        elif fusion == "ctx":
            loc = self.pairwise_loc_fc(pairwise_locs).view(B, L, T, H, dh)
            loc = torch.einsum("hblk,blthk->hblt", qh, loc) / math.sqrt(dh)
        else:  # cond
            sw = self.lang_cond_fc(residual).view(B, L, self.spatial_n_head, self.spatial_dim + 1).permute(2, 0, 1, 3)
            if self.spatial_n_head == 1:
                sw = sw.expand(H, -1, -1, -1)
            loc = torch.sigmoid(torch.einsum("hbld,bltd->hblt", sw[..., 1:], pairwise_locs) + sw[..., :1])
        if key_padding_mask is not None:
            m = key_padding_mask.view(1, B, 1, T)
            attn = attn.masked_fill(m, float("-inf"))
            loc = loc.masked_fill(m, 0.0 if fusion in ("mul", "cond") else float("-inf"))
        if fusion == "add":
            fused = (torch.softmax(attn, 3) + torch.softmax(loc, 3)) / 2
        else:
            fused = torch.log(torch.clamp(loc, min=1e-6)) + attn if fusion in ("mul", "cond") else loc + attn
            fused = torch.softmax(fused, 3)
        out = torch.matmul(fused, vh).permute(1, 2, 0, 3).reshape(B, L, D)
        out = self.dropout(self.fc(out))
        return self.layer_norm(out + residual), fused


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, batch_first=True, dropout=0.1,
                 activation="relu", prenorm=False):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=batch_first)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.activation = get_activation_fn(activation)
        self.prenorm = prenorm

    def _ffn(self, x):
        return self.linear2(self.dropout(self.activation(self.linear1(x))))

    def forward(self, tgt, tgt_mask=None, tgt_key_padding_mask=None):
        t2 = self.norm1(tgt) if self.prenorm else tgt
        t2, att = self.self_attn(query=t2, key=t2, value=t2, attn_mask=tgt_mask,
                                 key_padding_mask=tgt_key_padding_mask)
        tgt = tgt + self.dropout1(t2)
        tgt = self.norm2(tgt) if self.prenorm else self.norm1(tgt)
        tgt = tgt + self.dropout2(self._ffn(tgt))
        if not self.prenorm:
            tgt = self.norm2(tgt)
        return tgt, att


class TransformerSpatialEncoderLayer(TransformerEncoderLayer):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu",
                 spatial_multihead=True, spatial_dim=5, spatial_attn_fusion="mul"):
        super().__init__(d_model, nhead, dim_feedforward=dim_feedforward, dropout=dropout, activation=activation)
        del self.self_attn
        self.self_attn = MultiHeadAttentionSpatial(d_model, nhead, dropout=dropout,
                                                   spatial_multihead=spatial_multihead,
                                                   spatial_dim=spatial_dim,
                                                   spatial_attn_fusion=spatial_attn_fusion)
        self.activation_name = activation

    def forward(self, tgt, tgt_pairwise_locs, tgt_mask=None, tgt_key_padding_mask=None):
        t2, att = self.self_attn(tgt, tgt, tgt, tgt_pairwise_locs, key_padding_mask=tgt_key_padding_mask)
        tgt = self.norm1(tgt + self.dropout1(t2))
        tgt = self.norm2(tgt + self.dropout2(self._ffn(tgt)))
        return tgt, att


class CrossAttentionLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu",
                 k_dim=None, v_dim=None, prenorm=True):
        super().__init__()
        self.prenorm = prenorm
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True,
                                                    kdim=k_dim or d_model, vdim=v_dim or d_model)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.activation = get_activation_fn(activation)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None):
        t2 = self.norm1(tgt) if self.prenorm else tgt
        t2, att = self.multihead_attn(query=t2, key=memory, value=memory, attn_mask=memory_mask,
                                      key_padding_mask=memory_key_padding_mask)
        tgt = tgt + self.dropout2(t2)
        if not self.prenorm:
            tgt = self.norm1(tgt)
        if self.prenorm:
            t2 = self.norm3(tgt)   # post-norm keeps feeding the attention output (reference quirk, :65-68)
        t2 = self.linear2(self.dropout(self.activation(self.linear1(t2))))
        tgt = tgt + self.dropout3(t2)
        if not self.prenorm:
            tgt = self.norm3(tgt)
        return tgt, att

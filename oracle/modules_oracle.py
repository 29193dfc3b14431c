"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by msr3d_b200/).

Plain-PyTorch fp32 CPU restatement of the module-level hot path of the MSR3D
reference, written functionally over a reference-keyed ``state_dict`` so it can be
checked 1:1 against the reference's own modules (imported through
oracle/refshim.py in the authoring container -> tests/golden/*.npz) and then travel
to the GPU box, where /root/reference does not exist.

Everything runs in eval semantics (dropout off, BatchNorm on running statistics —
what the shipped configs do for the frozen encoder, pcd_pointnet_encoder.py:42-57).
Autograd works through every function here, which is how the backward of the CUDA
path is checked.

Parity pin: tests/test_oracle_golden.py compares these functions with fixtures
generated from the unmodified reference modules (tests/golden/gen_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

from . import ext_oracle as ext

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# PointNet++ object encoder
# --------------------------------------------------------------------------
def shared_mlp(sd: SD, prefix: str, x: torch.Tensor, bn_stats: Optional[dict] = None) -> torch.Tensor:
    """SharedMLP of (conv1x1 no-bias -> BatchNorm2d -> ReLU) blocks.

    pytorch_utils.py:11-36 (SharedMLP), :67-120 (_ConvBase order conv, bn, act),
    :39-58 (BatchNorm2d wrapper => key `...bn.bn.*`).  x: (B, C, npoint, nsample).
    bn_stats None: eval BatchNorm (running statistics; the frozen encoder).  A dict: TRAINING-mode BatchNorm
    (unfrozen encoder, pcd_pointnet_encoder.py:58-62 — batch statistics, nn.BatchNorm2d defaults momentum 0.1,
    eps 1e-5); the updated running_mean / running_var are returned in it, `sd` is left untouched.
    """
    i = 0
    while f"{prefix}.layer{i}.conv.weight" in sd:
        p = f"{prefix}.layer{i}"
        x = F.conv2d(x, sd[f"{p}.conv.weight"], sd.get(f"{p}.conv.bias"))
        if f"{p}.bn.bn.weight" in sd and bn_stats is not None:
            rm = sd[f"{p}.bn.bn.running_mean"].detach().clone()
            rv = sd[f"{p}.bn.bn.running_var"].detach().clone()
            x = F.batch_norm(x, rm, rv, sd[f"{p}.bn.bn.weight"], sd[f"{p}.bn.bn.bias"], True, 0.1, 1e-5)
            bn_stats[f"{p}.bn.bn.running_mean"], bn_stats[f"{p}.bn.bn.running_var"] = rm, rv
        elif f"{p}.bn.bn.weight" in sd:
            x = F.batch_norm(x, sd[f"{p}.bn.bn.running_mean"], sd[f"{p}.bn.bn.running_var"],
                             sd[f"{p}.bn.bn.weight"], sd[f"{p}.bn.bn.bias"], False, 0.0, 1e-5)
        x = F.relu(x)
        i += 1
    return x


def query_and_group(xyz, new_xyz, feats, radius: float, nsample: int):
    """QueryAndGroup.forward with default flags (pointnet2_utils.py:314-373)."""
    idx = ext.ball_query(new_xyz.contiguous(), xyz.contiguous(), radius, nsample)
    long_idx = idx.long()
    B, M, S = idx.shape
    xyz_t = xyz.transpose(1, 2)  # (B,3,N)
    g_xyz = torch.gather(xyz_t.unsqueeze(2).expand(-1, -1, M, -1), 3,
                         long_idx.unsqueeze(1).expand(-1, 3, -1, -1))
    g_xyz = g_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)  # :347 centre subtraction
    if feats is None:
        return g_xyz, idx
    C = feats.shape[1]
    g_f = torch.gather(feats.unsqueeze(2).expand(-1, -1, M, -1), 3,
                       long_idx.unsqueeze(1).expand(-1, C, -1, -1))
    return torch.cat([g_xyz, g_f], dim=1), idx  # :354-356 xyz channels first


def group_all(xyz, feats):
    """GroupAll.forward (pointnet2_utils.py:389-419): raw, un-centred xyz."""
    g_xyz = xyz.transpose(1, 2).unsqueeze(2)
    if feats is None:
        return g_xyz
    return torch.cat([g_xyz, feats.unsqueeze(2)], dim=1)


def sa_module(sd: SD, prefix: str, xyz, feats, npoint: Optional[int], radius, nsample,
              trace: Optional[dict] = None, bn_stats: Optional[dict] = None):
    """_PointnetSAModuleBase.forward, single-scale (pointnet2_modules.py:34-75)."""
    if npoint is not None:
        fps_idx = ext.furthest_point_sampling(xyz.contiguous(), npoint)  # :54-58
        new_xyz = torch.gather(xyz, 1, fps_idx.long().unsqueeze(-1).expand(-1, -1, 3))
        grouped, ball_idx = query_and_group(xyz, new_xyz, feats, radius, nsample)
        if trace is not None:
            trace[prefix + ".fps_idx"] = fps_idx
            trace[prefix + ".ball_idx"] = ball_idx
            trace[prefix + ".new_xyz"] = new_xyz
    else:
        new_xyz = None
        grouped = group_all(xyz, feats)
    h = shared_mlp(sd, f"{prefix}.mlps.0", grouped, bn_stats)
    h = h.max(dim=3).values  # F.max_pool2d(kernel=[1,nsample]) + squeeze, :68-71
    if trace is not None:
        trace[prefix + ".features"] = h
    return new_xyz, h


def pointnetpp(sd: SD, prefix: str, pc: torch.Tensor, sa_n_points: Sequence, sa_n_samples: Sequence,
               sa_radii: Sequence, trace: Optional[dict] = None, bn_stats: Optional[dict] = None) -> torch.Tensor:
    """PointNetPP.forward (modules/layers/pointnet.py:55-63) incl. break_up_pc (:6-19)."""
    xyz = pc[..., 0:3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
    for i in range(len(sa_n_points)):
        xyz, feats = sa_module(sd, f"{prefix}.encoder.{i}", xyz, feats, sa_n_points[i],
                               sa_radii[i], sa_n_samples[i], trace, bn_stats)
    flat = feats.reshape(feats.size(0), -1)
    return F.linear(flat, sd[f"{prefix}.fc.weight"], sd[f"{prefix}.fc.bias"])


def mlp_head(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """get_mlp_head (modules/utils.py:25-32), eval: Linear, ReLU, LayerNorm(eps 1e-12), Linear."""
    x = F.relu(F.linear(x, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"]))
    x = F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.2.weight"], sd[f"{prefix}.2.bias"], 1e-12)
    return F.linear(x, sd[f"{prefix}.4.weight"], sd[f"{prefix}.4.bias"])


def pcd_obj_encoder(sd: SD, prefix: str, obj_pcds: torch.Tensor, sa_n_points=(32, 16, None),
                    sa_n_samples=(32, 32, None), sa_radii=(0.2, 0.4, None),
                    trace: Optional[dict] = None, bn_stats: Optional[dict] = None):
    """PcdObjEncoder.forward (modules/vision/pcd_pointnet_encoder.py:48-74)."""
    B, O = obj_pcds.shape[:2]
    flat = obj_pcds.reshape(B * O, obj_pcds.shape[2], obj_pcds.shape[3])
    emb = pointnetpp(sd, f"{prefix}.pcd_net", flat, sa_n_points, sa_n_samples, sa_radii, trace, bn_stats)
    emb = emb.reshape(B, O, -1)
    return emb, mlp_head(sd, f"{prefix}.obj3d_clf_pre_head", emb)


# --------------------------------------------------------------------------
# Situated spatial-attention transformer
# --------------------------------------------------------------------------
def pairwise_locs_center(centers: torch.Tensor, eps: float = 1e-10, dist_norm: bool = True):
    """calc_pairwise_locs, pairwise_rel_type='center', spatial_dim=5 (modules/utils.py:88-137)."""
    diff = centers.unsqueeze(2) - centers.unsqueeze(1)  # (B,L,L,3)
    dist = torch.sqrt((diff ** 2).sum(3) + eps)
    if dist_norm:
        norm = dist / dist.flatten(1).max(dim=1).values.view(-1, 1, 1)
    else:
        norm = dist
    dist2d = torch.sqrt((diff[..., :2] ** 2).sum(3) + eps)
    return torch.stack([norm, diff[..., 2] / dist, dist2d / dist,
                        diff[..., 1] / dist2d, diff[..., 0] / dist2d], dim=3)


def quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    """quaternion_to_matrix (modules/utils.py:60-75): conjugate (x,y,z negated)."""
    x, y, z, w = -q[:, 0], -q[:, 1], -q[:, 2], q[:, 3]
    xx, yy, zz = x * x, y * y, z * z
    xy, xz, xw, yz, yw, zw = x * y, x * z, x * w, y * z, y * w, z * w
    row0 = torch.stack([1 - 2 * (yy + zz), 2 * (xy + zw), 2 * (xz - yw)], -1)
    row1 = torch.stack([2 * (xy - zw), 1 - 2 * (xx + zz), 2 * (yz + xw)], -1)
    row2 = torch.stack([2 * (xz + yw), 2 * (yz - xw), 1 - 2 * (xx + yy)], -1)
    return torch.stack([row0, row1, row2], -2)


def to_agent_frame(centers, anchor_loc, anchor_ori):
    """transform_to_agent_coor (modules/utils.py:77-82): (p - anchor) @ R."""
    return torch.matmul(centers - anchor_loc.unsqueeze(1), quat_to_rot(anchor_ori))


def fourier_features(pos: torch.Tensor, num_bands: int = 10, max_freq: float = 15.0):
    """generate_fourier_features, concat_pos=True (model/ose3d_situation.py:31-59).

    Layout: [pos(d) | sin(pi*pos_0*f_0..f_9), sin(pos_1..), ... | cos(...)]."""
    freqs = torch.linspace(1.0, max_freq, num_bands, device=pos.device)
    scaled = (pos.unsqueeze(-1) * freqs).flatten(2)  # (B,N,d*bands), dim-major
    return torch.cat([pos, torch.sin(math.pi * scaled), torch.cos(math.pi * scaled)], dim=-1)


def linear_ln(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """nn.Sequential(Linear, LayerNorm) (ose3d_situation.py:251-258)."""
    x = F.linear(x, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"])
    return F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.1.weight"], sd[f"{prefix}.1.bias"], 1e-5)


def spatial_mha(sd: SD, prefix: str, x: torch.Tensor, pairwise: torch.Tensor,
                key_padding_mask: Optional[torch.Tensor], n_head: int, fusion: str = "cond",
                spatial_multihead: bool = True, drop_fc: Optional[torch.Tensor] = None):
    """MultiHeadAttentionSpatial.forward with q=k=v=x (modules/layers/transformers.py:200-252)."""
    B, L, D = x.shape
    dh = D // n_head

    def proj(name):
        return F.linear(x, sd[f"{prefix}.{name}.weight"], sd[f"{prefix}.{name}.bias"]) \
            .view(B, L, n_head, dh).permute(2, 0, 1, 3)  # (h,B,L,dh)

    q, k, v = proj("w_qs"), proj("w_ks"), proj("w_vs")
    attn = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)  # (h,B,L,T)
    sp_heads = n_head if spatial_multihead else 1
    sdim = pairwise.shape[-1]
    if fusion == "cond":  # :218-227
        sw = F.linear(x, sd[f"{prefix}.lang_cond_fc.weight"], sd[f"{prefix}.lang_cond_fc.bias"])
        sw = sw.view(B, L, sp_heads, sdim + 1).permute(2, 0, 1, 3)
        if sp_heads == 1:
            sw = sw.expand(n_head, -1, -1, -1)
        loc = torch.einsum("hbld,bltd->hblt", sw[..., 1:], pairwise) + sw[..., :1]
        loc = torch.sigmoid(loc)
    elif fusion in ("mul", "bias", "add"):  # :207-213
        loc = F.linear(pairwise, sd[f"{prefix}.pairwise_loc_fc.weight"],
                       sd[f"{prefix}.pairwise_loc_fc.bias"]).permute(3, 0, 1, 2)
        if fusion == "mul":
            loc = F.relu(loc)
        if not spatial_multihead:
            loc = loc.expand(n_head, -1, -1, -1)
    elif fusion == "ctx":  # :214-217
        loc = F.linear(pairwise, sd[f"{prefix}.pairwise_loc_fc.weight"],
                       sd[f"{prefix}.pairwise_loc_fc.bias"]).view(B, L, L, n_head, dh)
        loc = torch.einsum("hblk,blthk->hblt", q, loc) / math.sqrt(dh)
    else:
        raise NotImplementedError(fusion)
    if key_padding_mask is not None:  # :229-235
        m = key_padding_mask.view(1, B, 1, L)
        attn = attn.masked_fill(m, float("-inf"))
        loc = loc.masked_fill(m, 0.0 if fusion in ("mul", "cond") else float("-inf"))
    if fusion == "add":
        fused = (torch.softmax(attn, 3) + torch.softmax(loc, 3)) / 2
    else:
        if fusion in ("mul", "cond"):
            fused = torch.log(torch.clamp(loc, min=1e-6)) + attn
        else:
            fused = loc + attn
        fused = torch.softmax(fused, 3)
    out = torch.matmul(fused, v).permute(1, 2, 0, 3).reshape(B, L, D)
    out = F.linear(out, sd[f"{prefix}.fc.weight"], sd[f"{prefix}.fc.bias"])
    if drop_fc is not None:  # explicit multiplicative dropout mask (training-mode checks)
        out = out * drop_fc
    out = F.layer_norm(out + x, (D,), sd[f"{prefix}.layer_norm.weight"],
                       sd[f"{prefix}.layer_norm.bias"], 1e-5)
    return out, fused


def spatial_encoder_layer(sd: SD, prefix: str, x, pairwise, key_padding_mask, n_head: int,
                          fusion: str = "cond", activation: str = "gelu",
                          spatial_multihead: bool = True, masks: Optional[dict] = None):
    """TransformerSpatialEncoderLayer.forward (transformers.py:314-329).

    Note the double residual: self_attn already returns LN(out + x) (:250-251)."""
    D = x.shape[-1]
    mk = masks or {}
    t2, att = spatial_mha(sd, f"{prefix}.self_attn", x, pairwise, key_padding_mask, n_head,
                          fusion, spatial_multihead, mk.get("fc"))
    if "d1" in mk:
        t2 = t2 * mk["d1"]
    x = F.layer_norm(x + t2, (D,), sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"], 1e-5)
    act = {"gelu": F.gelu, "relu": F.relu}[activation]
    hid = act(F.linear(x, sd[f"{prefix}.linear1.weight"], sd[f"{prefix}.linear1.bias"]))
    if "ffn" in mk:
        hid = hid * mk["ffn"]
    h = F.linear(hid, sd[f"{prefix}.linear2.weight"], sd[f"{prefix}.linear2.bias"])
    if "d2" in mk:
        h = h * mk["d2"]
    x = F.layer_norm(x + h, (D,), sd[f"{prefix}.norm2.weight"], sd[f"{prefix}.norm2.bias"], 1e-5)
    return x, att


def cross_attention_layer(sd: SD, prefix: str, tgt: torch.Tensor, memory: torch.Tensor, n_head: int,
                          activation: str = "gelu", masks: Optional[dict] = None) -> torch.Tensor:
    """CrossAttentionLayer.forward, prenorm (modules/layers/transformers.py:22-73) around nn.MultiheadAttention
    (batch_first, kdim = vdim = d_model: packed in_proj).  The reference passes neither an attention mask nor a key
    padding mask (ose3d_situation.py:418 hands `tgt_key_padding_mask`, which the layer never forwards).
    masks (training-mode checks): 'attn' (B, h, L, T) on the attention probabilities, 'd2', 'ffn', 'd3'."""
    B, L, D = tgt.shape
    T = memory.shape[1]
    dh = D // n_head
    mk = masks or {}
    W, b = sd[f"{prefix}.multihead_attn.in_proj_weight"], sd[f"{prefix}.multihead_attn.in_proj_bias"]
    t2 = F.layer_norm(tgt, (D,), sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"], 1e-5)
    q = F.linear(t2, W[:D], b[:D]).view(B, L, n_head, dh).transpose(1, 2)
    k = F.linear(memory, W[D:2 * D], b[D:2 * D]).view(B, T, n_head, dh).transpose(1, 2)
    v = F.linear(memory, W[2 * D:], b[2 * D:]).view(B, T, n_head, dh).transpose(1, 2)
    att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh), -1)
    if "attn" in mk:
        att = att * mk["attn"]
    o = torch.matmul(att, v).transpose(1, 2).reshape(B, L, D)
    o = F.linear(o, sd[f"{prefix}.multihead_attn.out_proj.weight"], sd[f"{prefix}.multihead_attn.out_proj.bias"])
    if "d2" in mk:
        o = o * mk["d2"]
    x = tgt + o
    t3 = F.layer_norm(x, (D,), sd[f"{prefix}.norm3.weight"], sd[f"{prefix}.norm3.bias"], 1e-5)
    act = {"gelu": F.gelu, "relu": F.relu}[activation]
    hid = act(F.linear(t3, sd[f"{prefix}.linear1.weight"], sd[f"{prefix}.linear1.bias"]))
    if "ffn" in mk:
        hid = hid * mk["ffn"]
    h = F.linear(hid, sd[f"{prefix}.linear2.weight"], sd[f"{prefix}.linear2.bias"])
    if "d3" in mk:
        h = h * mk["d3"]
    return x + h


def ose3d_situation(sd: SD, data: dict, *, n_head: int = 8, num_layers: int = 3,
                    situation_type: str = "as_transform_for_objects",
                    obj_loc_encoding: str = "same_all", fusion: str = "cond",
                    activation: str = "gelu", spatial_dist_norm: bool = True,
                    sa_n_points=(32, 16, None), sa_n_samples=(32, 32, None),
                    sa_radii=(0.2, 0.4, None), obj_embeds: Optional[torch.Tensor] = None,
                    trace: Optional[dict] = None, layer_masks: Optional[list] = None,
                    spatial_multihead: bool = True, unfrozen_bn_stats: Optional[dict] = None,
                    cross_masks: Optional[list] = None):
    """OSE3DSituation.forward (model/ose3d_situation.py:288-454) for vision_backbone
    'gtpcd', use_spatial_attn/use_anchor/use_orientation True, attn_flat off, and
    situation_type in {'as_transform_for_objects' (shipped, configs/msr3d.yaml:177),
    'as_object' (LEO baseline, configs/leo_3_dataset.yaml:180-181), 'as_object_add_loc', 'as_embedding',
    'as_cross_attention'}.

    Returns (obj_tokens (B,N',H), obj_masks (B,N') bool with True = valid)."""
    if obj_embeds is None:
        obj_embeds = pcd_obj_encoder(sd, "obj_encoder", data["obj_fts"], sa_n_points,
                                     sa_n_samples, sa_radii, trace, unfrozen_bn_stats)[0]
        if unfrozen_bn_stats is None:
            obj_embeds = obj_embeds.detach()  # frozen encoder, pcd_pointnet_encoder.py:50-57
        # else: vision.args.freeze False (:58-62) — training-mode BatchNorm, gradients flow into the encoder
    feat = F.linear(obj_embeds, sd["obj_linear_projection.weight"], sd["obj_linear_projection.bias"])
    B, N, H = feat.shape
    mask = ~data["obj_masks"]  # True = padded (:315)
    loc = data["obj_locs"]
    type_w = sd["object_type_embedding.weight"]
    ori_feat = sd["object_orientation_feat"].expand(B, N, -1)
    type_emb = type_w[0].view(1, 1, H).expand(B, N, -1)
    if situation_type in ("as_object", "as_object_add_loc"):  # :334-357 prepend the agent token
        a_feat = sd["anchor_feat"].expand(B, -1, -1)
        a_ori = F.linear(fourier_features(data["anchor_orientation"].unsqueeze(1)),
                         sd["orientation_encoder.weight"], sd["orientation_encoder.bias"])
        a_loc = torch.cat([data["anchor_locs"].unsqueeze(1), sd["anchor_size"].expand(B, -1, -1)], -1)
        feat = torch.cat([a_feat, feat], 1)
        mask = torch.cat([torch.zeros(B, 1, dtype=torch.bool, device=mask.device), mask], 1)
        ori_feat = torch.cat([a_ori, ori_feat], 1)
        loc = torch.cat([a_loc, loc], 1)
        type_emb = torch.cat([type_w[1].view(1, 1, H).expand(B, 1, -1), type_emb], 1)
    x = feat + ori_feat + type_emb  # :359-363
    pairwise = pairwise_locs_center(loc[..., :3], dist_norm=spatial_dist_norm)  # :368-375
    for i in range(num_layers):  # :378-435
        if obj_loc_encoding == "diff_all":
            qpos = linear_ln(sd, f"loc_layers.{i}", loc)
        elif situation_type == "as_object_add_loc":
            qpos = linear_ln(sd, "loc_embedding_encoder", fourier_features(loc[..., :3])) + \
                linear_ln(sd, "size_embedding_encoder", loc[..., 3:])
        elif situation_type == "as_embedding":  # :385-393 situation embedding added to every object
            n_tok = loc.size(1)
            s_loc = linear_ln(sd, "loc_embedding_encoder",
                              fourier_features(data["anchor_locs"].unsqueeze(1).expand(-1, n_tok, -1)))
            s_ori = F.linear(fourier_features(data["anchor_orientation"].unsqueeze(1).expand(-1, n_tok, -1)),
                             sd["orientation_encoder.weight"], sd["orientation_encoder.bias"])
            qpos = linear_ln(sd, "loc_embedding_encoder", fourier_features(loc[..., :3])) + \
                linear_ln(sd, "size_embedding_encoder", loc[..., 3:]) + s_ori + s_loc
        elif situation_type == "as_transform_for_objects":  # :399-404
            tl = to_agent_frame(loc[..., :3], data["anchor_locs"], data["anchor_orientation"])
            qpos = linear_ln(sd, "loc_embedding_encoder", fourier_features(tl)) + \
                linear_ln(sd, "size_embedding_encoder", loc[..., 3:])
        else:
            qpos = linear_ln(sd, "loc_layers.0", loc)
        if obj_loc_encoding in ("same_all", "diff_all") or i == 0:
            x = x + qpos
        if situation_type == "as_cross_attention":  # :411-418 situation as the memory of a cross-attention layer
            n_tok = loc.size(1)
            s_feat = linear_ln(sd, "loc_embedding_encoder",
                               fourier_features(data["anchor_locs"].unsqueeze(1).expand(-1, n_tok, -1))) + \
                F.linear(fourier_features(data["anchor_orientation"].unsqueeze(1).expand(-1, n_tok, -1)),
                         sd["orientation_encoder.weight"], sd["orientation_encoder.bias"])
            x = cross_attention_layer(sd, f"situation_condition_layer.{i}", x, s_feat, n_head, activation,
                                      masks=cross_masks[i] if cross_masks else None)
        x, att = spatial_encoder_layer(sd, f"spatial_encoder.{i}", x, pairwise, mask, n_head,
                                       fusion, activation, spatial_multihead,
                                       masks=layer_masks[i] if layer_masks else None)
        if trace is not None:
            trace[f"layer{i}.out"] = x
            trace[f"layer{i}.attn"] = att
    return x, ~mask


# --------------------------------------------------------------------------
# Scene-token splice into the LLM input (the step right after the prompter)
# --------------------------------------------------------------------------
def scene_splice(W: torch.Tensor, b: torch.Tensor, inputs_embeds: torch.Tensor,
                 attention_mask: torch.Tensor, input_ids: torch.Tensor, obj_tokens: torch.Tensor,
                 obj_masks: torch.Tensor, scene_sp_token: int = 31495):
    """MSR3D.build_embeds, scene branch (model/msr3d/msr3d.py:277-287).

    scene_embeds = llm_proj(obj_tokens) cast to the embedding dtype (:277-279); written over the
    rows of `inputs_embeds` whose id is the scene placeholder, in the row-major order of
    `torch.where` (:280,282); the attention mask becomes bool and takes `obj_masks` at the same
    positions (:283-287).  Returns new tensors (the reference writes in place into the embedding
    lookup's output); differentiable w.r.t. W, b, obj_tokens.
    Pinned by tests/golden/splice.npz (tests/golden/gen_golden_splice.py runs the reference method).
    """
    scene = F.linear(obj_tokens, W, b).to(inputs_embeds.dtype)
    idx = torch.where(input_ids == scene_sp_token)
    out = inputs_embeds.clone()
    out[idx] = scene.reshape(-1, scene.shape[-1])
    am = attention_mask.to(obj_masks.dtype).clone()
    am[idx] = obj_masks.reshape(-1)
    return out, am

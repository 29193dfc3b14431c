"""`OSE3DSituation` — the MSR3D visual prompter — as a drop-in nn.Module.

Same registry name, constructor (`cfg.model.*` read attribute-style), `forward(data_dict)`
contract, `.device`, `.forward_gtpcd` and state-dict keys as model/ose3d_situation.py:157-454,
so `MSR3D.__init__` (model/msr3d/msr3d.py:52) can build it through MODEL_REGISTRY unchanged and
reference checkpoints load with strict=False.

Differences in *how*, not *what*:
  * the frozen PointNet++ encoder runs the fused tensor-core path (msr3d_b200/encoder.py) and
    skips the 607-way head whose output the reference discards (ose3d_situation.py:285);
  * `query_pos` under obj_loc_encoding in {same_0, same_all} does not depend on the layer index
    and is computed once instead of once per layer (ose3d_situation.py:378-410);
  * no per-layer device->host NaN assert (transformers.py:246).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .registry import MODEL_REGISTRY, BaseModel, build_module
from .spatial import (CrossAttentionLayer, TransformerEncoderLayer, TransformerSpatialEncoderLayer,
                      calc_pairwise_locs, generate_fourier_features, init_weights_bert, layer_repeat,
                      transform_to_agent_coor)


def _cfg_get(cfg, key, default=None):
    if hasattr(cfg, "get"):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class ObjColorEncoder(nn.Module):
    """GMM colour encoder of the unused 'gt' backbone (ose3d_situation.py:137-154)."""

    def __init__(self, hidden_size, dropout=0):
        super().__init__()
        self.ft_linear = nn.Sequential(nn.Linear(3, hidden_size), nn.ReLU(),
                                       nn.LayerNorm(hidden_size, eps=1e-12), nn.Dropout(dropout))

    def forward(self, obj_colors):
        return torch.sum(self.ft_linear(obj_colors[..., 1:]) * obj_colors[..., :1], 2)


class _AttFlat(nn.Module):
    """MCAN attentional flattening (ose3d_situation.py:62-134); off in every shipped config."""

    def __init__(self, hidden_size, flat_mlp_size=512, flat_glimpses=1, flat_out_size=1024, pdrop=0.1):
        super().__init__()
        fc = nn.Module()
        fc.linear = nn.Linear(hidden_size, flat_mlp_size)
        fc.gelu = nn.GELU()
        fc.dropout = nn.Dropout(pdrop)
        self.mlp = nn.Module()
        self.mlp.fc = fc
        self.mlp.linear = nn.Linear(flat_mlp_size, flat_glimpses)
        self.flat_glimpses = flat_glimpses
        self.linear_merge = nn.Linear(hidden_size * flat_glimpses, flat_out_size)

    def forward(self, x, x_mask):
        att = self.mlp.linear(self.mlp.fc.dropout(self.mlp.fc.gelu(self.mlp.fc.linear(x))))
        if x_mask is not None:
            att = att.masked_fill(x_mask.squeeze(1).squeeze(1).unsqueeze(2), -1e9)
        att = torch.softmax(att, dim=1)
        pooled = torch.cat([torch.sum(att[:, :, i:i + 1] * x, dim=1) for i in range(self.flat_glimpses)], 1)
        return self.linear_merge(pooled), att


@MODEL_REGISTRY.register()
class OSE3DSituation(BaseModel):
    def __init__(self, cfg):
        super().__init__(cfg)
        self.cfg = c = cfg.model
        H = c.hidden_size
        se = c.spatial_encoder
        self.vision_backbone_name = c.vision_backbone_name
        self.use_spatial_attn = c.use_spatial_attn
        self.use_anchor = c.use_anchor
        self.use_orientation = c.use_orientation
        if self.use_anchor:
            self.anchor_feat = nn.Parameter(torch.zeros(1, 1, H))
            self.anchor_size = nn.Parameter(torch.ones(1, 1, 3), requires_grad=False)
        if self.use_orientation:
            self.object_orientation_feat = nn.Parameter(torch.zeros(1, 1, H))
            self.orientation_encoder = nn.Linear(c.fourier_size, H)
        self.object_type_embedding = nn.Embedding(2, embedding_dim=H)
        if self.vision_backbone_name == "gt":
            self.obj_sem_encoder = nn.Linear(c.label_size, H)
            self.color_encoder = ObjColorEncoder(H)
        else:
            self.obj_encoder = build_module("vision", c.vision)
            if c.vision.name == "PcdObjEncoder":
                self.obj_linear_projection = nn.Linear(c.vision.args.sa_mlps[-1][-1], H)
            elif c.vision.name == "PointBERTPcdObjEncoder":
                self.obj_linear_projection = nn.Linear(c.vision.args.trans_dim * 2, H)
        if c.use_spatial_attn:
            layer = TransformerSpatialEncoderLayer(
                H, nhead=se.num_attention_heads, dim_feedforward=se.dim_feedforward, dropout=se.dropout,
                activation=se.activation, spatial_dim=se.spatial_dim,
                spatial_multihead=se.spatial_multihead, spatial_attn_fusion=se.spatial_attn_fusion)
        else:
            layer = TransformerEncoderLayer(H, nhead=se.num_attention_heads,
                                            dim_feedforward=se.dim_feedforward, dropout=se.dropout,
                                            activation=se.activation)
        self.spatial_encoder = layer_repeat(layer, se.num_layers)
        n_loc = se.num_layers if se.obj_loc_encoding == "diff_all" else 1
        self.loc_layers = layer_repeat(nn.Sequential(nn.Linear(se.dim_loc, H), nn.LayerNorm(H)), n_loc)
        self.spatial_encoder.apply(init_weights_bert)
        self.loc_layers.apply(init_weights_bert)
        if c.attn_flat.use_attn_flat:
            self.attflat_visual = _AttFlat(H, c.attn_flat.mcan_flat_mlp_size, c.attn_flat.mcan_flat_glimpses,
                                           c.attn_flat.mcan_flat_out_size, 0.1)
        if self.use_anchor:
            nn.init.normal_(self.anchor_feat, std=.02)
        self.situation_type = _cfg_get(c, "situation_type", "as_object")
        if self.situation_type in ("as_object_add_loc", "as_embedding", "as_transform_for_objects",
                                   "as_cross_attention", "as_dit_attention"):
            self.loc_embedding_encoder = nn.Sequential(nn.Linear(c.loc_fourier_dim, H), nn.LayerNorm(H))
            self.size_embedding_encoder = nn.Sequential(nn.Linear(3, H), nn.LayerNorm(H))
        if self.situation_type == "as_cross_attention":
            self.situation_condition_layer = layer_repeat(
                CrossAttentionLayer(H, nhead=se.num_attention_heads, dim_feedforward=se.dim_feedforward,
                                    dropout=se.dropout, activation=se.activation), se.num_layers)
        elif self.situation_type == "as_dit_attention":
            raise NotImplementedError("as_dit_attention is unreachable in the reference forward "
                                      "(ose3d_situation.py:273 vs :423) and needs timm")

    @property
    def device(self):
        return next(self.parameters()).device

    def obj_encoder_is_native(self) -> bool:
        return hasattr(self, "obj_encoder") and hasattr(self.obj_encoder, "encode")

    # ------------------------------------------------------------------ encoder
    def forward_gtpcd(self, data_dict):
        enc = self.obj_encoder
        embeds = enc.encode(data_dict["obj_fts"]) if hasattr(enc, "encode") else enc(data_dict["obj_fts"])[0]
        return self.obj_linear_projection(embeds)

    # ------------------------------------------------------------------ query position
    def _situation_feat(self, data_dict, n):
        loc = data_dict["anchor_locs"].unsqueeze(1).expand(-1, n, -1)
        ori = data_dict["anchor_orientation"].unsqueeze(1).expand(-1, n, -1)
        return (self.loc_embedding_encoder(generate_fourier_features(loc)),
                self.orientation_encoder(generate_fourier_features(ori)))

    def _query_pos(self, i, loc, data_dict):
        se = self.cfg.spatial_encoder
        if se.obj_loc_encoding == "diff_all":
            return self.loc_layers[i](loc)
        st = self.situation_type
        if st == "as_object_add_loc":
            return self.loc_embedding_encoder(generate_fourier_features(loc[:, :, :3])) + \
                self.size_embedding_encoder(loc[:, :, 3:])
        if st == "as_embedding":
            s_loc, s_ori = self._situation_feat(data_dict, loc.size(1))
            return self.loc_embedding_encoder(generate_fourier_features(loc[:, :, :3])) + \
                self.size_embedding_encoder(loc[:, :, 3:]) + s_ori + s_loc
        if st == "as_transform_for_objects":
            agent = transform_to_agent_coor(loc[:, :, :3], data_dict["anchor_locs"],
                                            data_dict["anchor_orientation"])
            return self.loc_embedding_encoder(generate_fourier_features(agent)) + \
                self.size_embedding_encoder(loc[:, :, 3:])
        return self.loc_layers[0](loc)

    # ------------------------------------------------------------------ forward
    def forward(self, data_dict):
        if self.vision_backbone_name == "gt":
            raise AssertionError("the 'gt' backbone is disabled in the reference (ose3d_situation.py:301)")
        if "single_obj" in data_dict:  # token of a single object (ose3d_situation.py:306-312)
            keep = data_dict["obj_fts"]
            data_dict["obj_fts"] = data_dict["single_obj"]
            data_dict["single_obj_token"] = self.forward_gtpcd(data_dict)
            data_dict["obj_fts"] = keep
            return data_dict
        if self.obj_encoder_is_native() and data_dict["obj_fts"].is_cuda:
            # the reference's call path (MSR3D.forward -> loss.backward(), trainer/leo_trainer.py:189-195) on the
            # hand-written kernels: forward and backward of the fused engine behind a torch.autograd.Function
            from . import autograd as _ag
            if _ag.eligible(self, data_dict) is None:
                data_dict["obj_tokens"] = _ag.fused_forward(self, data_dict)
                data_dict["oatt"] = None
                data_dict["obj_masks"] = _ag.token_masks(self, data_dict)
                return data_dict
        feat = self.forward_gtpcd(data_dict)
        mask = ~data_dict["obj_masks"]          # True = padded
        B, N, H = feat.shape
        dev = feat.device
        loc = data_dict["obj_locs"]
        type_w = self.object_type_embedding.weight
        type_emb = type_w[0].view(1, 1, H).expand(B, N, -1)
        ori_feat = self.object_orientation_feat.expand(B, N, -1) if self.use_orientation else None
        if self.use_anchor and self.situation_type in ("as_object", "as_object_add_loc"):
            a_loc = data_dict["anchor_locs"]
            assert a_loc.shape[-1] == 3
            feat = torch.cat((self.anchor_feat.expand(B, -1, -1), feat), 1)
            mask = torch.cat((torch.zeros((B, 1), device=dev, dtype=torch.bool), mask), 1)
            if self.use_orientation:
                a_ori = self.orientation_encoder(
                    generate_fourier_features(data_dict["anchor_orientation"].unsqueeze(1).to(dev)))
                ori_feat = torch.cat((a_ori, ori_feat), 1)
            loc = torch.cat((torch.cat((a_loc.unsqueeze(1), self.anchor_size.expand(B, -1, -1)), -1), loc), 1)
            type_emb = torch.cat((type_w[1].view(1, 1, H).expand(B, 1, -1), type_emb), 1)
        x = feat + type_emb if ori_feat is None else feat + ori_feat + type_emb

        se = self.cfg.spatial_encoder
        pairwise = None
        if self.cfg.use_spatial_attn:
            pairwise = calc_pairwise_locs(loc[:, :, :3], loc[:, :, 3:], pairwise_rel_type=se.pairwise_rel_type,
                                          spatial_dist_norm=se.spatial_dist_norm, spatial_dim=se.spatial_dim)
        with torch.autocast(device_type=x.device.type, enabled=False):  # ose3d_situation.py:377
            x = x.float()
            qpos = None
            for i, layer in enumerate(self.spatial_encoder):
                if qpos is None or se.obj_loc_encoding == "diff_all":
                    qpos = self._query_pos(i, loc, data_dict)   # layer-invariant otherwise
                if se.obj_loc_encoding in ("same_all", "diff_all") or i == 0:
                    x = x + qpos
                if self.situation_type == "as_cross_attention":
                    s_loc, s_ori = self._situation_feat(data_dict, loc.size(1))
                    x, _ = self.situation_condition_layer[i](x, s_loc + s_ori, tgt_key_padding_mask=mask)
                if self.cfg.use_spatial_attn:
                    x, _ = layer(x, pairwise, tgt_key_padding_mask=mask)
                else:
                    x, _ = layer(x, tgt_key_padding_mask=mask)
        if self.cfg.attn_flat.use_attn_flat:
            x, data_dict["oatt"] = self.attflat_visual(x, mask)
        else:
            data_dict["oatt"] = None
        data_dict["obj_tokens"] = x
        data_dict["obj_masks"] = ~mask
        return data_dict

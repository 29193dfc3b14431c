"""Fused execution engine for the MSR3D prompter.

`FusedPrompter` takes a (drop-in) `OSE3DSituation` and runs its forward, backward and optimiser step as a fixed
program of ~70 launches of this repo's kernels (TMA-fed 3xTF32 tcgen05 GEMMs from csrc/gemm3x.cu, the one-launch
attention block of csrc/attn_block.cu, everything else from csrc/spatial.cu, csrc/modes.cu, csrc/optim.cu, and the
fused PointNet++ encoder), replayed as ONE CUDA graph per step in the pipelined schedules.  It replaces what the
reference does with ~500 ATen launches plus a host sync per layer.

Parameters are re-homed into one flat fp32 buffer (each `p.data` becomes a view) laid out as
[ late biases | late weights | early weights | early biases | parameters the configured mode never touches ]
(dist.flat_layout) with the input projections of every attention block (w_qs, w_ks, w_vs, lang_cond_fc) adjacent, so
that QKV+cond is a single (816 x 256) GEMM, its weight gradient another one, and the gradient all-reduce and AdamW are
single flat operations.  State-dict keys and shapes are unchanged.

Covered (`supported()`): hidden 256 / 8 heads / spatial_dim 5 / gelu, obj_loc_encoding same_all, same_0 or diff_all,
spatial_attn_fusion cond (every shipped config) | mul | bias | ctx | add, situation_type as_transform_for_objects
(shipped) | as_object | as_object_add_loc | as_embedding | as_cross_attention, encoder frozen (shipped) or unfrozen
(encoder_train.py).  Anything else (attn_flat, no spatial attention) keeps running through the module tree
in spatial.py / prompter.py.  The pipelined / prefetch / data-parallel schedules serve the shipped situation mode with
the frozen encoder; the other modes and the unfrozen encoder run the sequential schedule or the module call.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _C, ops

c_int, c_uint, c_ll, c_float, c_vp = ctypes.c_int, ctypes.c_uint, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p
c_ull = ctypes.c_ulonglong


class GemmDesc(ctypes.Structure):
    """Mirror of `msr3d_gemm_desc` (include/msr3d_b200.h)."""
    _fields_ = [("A", c_vp), ("stride_am", c_ll), ("stride_ak", c_ll),
                ("B", c_vp), ("stride_bn", c_ll), ("stride_bk", c_ll),
                ("bias", c_vp), ("residual", c_vp), ("ld_residual", c_ll),
                ("aux", c_vp), ("ld_aux", c_ll), ("aux_out", c_vp), ("out", c_vp), ("ld_out", c_ll),
                ("M", c_int), ("N", c_int), ("K", c_int),
                ("act", c_int), ("split", c_int), ("accumulate", c_int), ("epilogue", c_int), ("ksplit", c_int),
                ("a_xform", c_int), ("b_xform", c_int),
                ("a_site", c_uint), ("b_site", c_uint), ("epi_site", c_uint),
                ("dropout_p", c_float), ("seed", c_ull), ("seed_ptr", c_vp),
                ("row_map", c_vp), ("out_dtype", c_int), ("tile_n", c_int)]


STAGE_MAX = 8


class StageDesc(ctypes.Structure):
    """Mirror of `msr3d_stage_desc` (include/msr3d_b200.h)."""
    _fields_ = [("rot_a", c_vp * STAGE_MAX), ("rot_b", c_vp * STAGE_MAX), ("rot_c", c_vp * STAGE_MAX), ("rot_n", c_ll * STAGE_MAX),
                ("cp_dst", c_vp * STAGE_MAX), ("cp_src", c_vp * STAGE_MAX), ("cp_n", c_ll * STAGE_MAX),
                ("mask_a", c_vp), ("mask_b", c_vp), ("mask_c", c_vp), ("key_pad", c_vp), ("n_mask", c_ll),
                ("n_rotate", c_int), ("n_copy", c_int)]


AR_MAX_WORLD, AR_MAX_BLOCKS = 8, 64


class AllreduceDesc(ctypes.Structure):
    """Mirror of `msr3d_allreduce_desc` (include/msr3d_b200.h)."""
    _fields_ = [("buf", c_vp * AR_MAX_WORLD), ("flags", c_vp * AR_MAX_WORLD), ("partials", c_vp * AR_MAX_WORLD),
                ("epoch", c_vp), ("gate", c_vp), ("rank", c_int), ("world", c_int), ("channel", c_int)]


_C.register_protos({
    "msr3d_gemm_tc_f32": [ctypes.POINTER(GemmDesc), c_vp],
    "msr3d_gemm_tf32x3_supported": [ctypes.POINTER(GemmDesc)],
    "msr3d_gemm_tf32x3_f32": [ctypes.POINTER(GemmDesc), c_int, c_vp],
    "msr3d_ffn_pair_fwd_f32": [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_ull, c_vp, c_float, c_uint,
                               c_vp, c_vp, c_vp, c_int, c_vp],
    "msr3d_ffn_pair_bwd_f32": [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_vp],
    "msr3d_spatial_geometry_f32": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "msr3d_qpos_fwd_f32": [c_vp, c_vp, c_int] + [c_vp] * 14 + [c_vp],
    "msr3d_qpos_bwd_f32": [c_vp, c_vp, c_vp, c_vp, c_int] + [c_vp] * 14 + [c_vp],
    "msr3d_attn_fwd_f32": [c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp],
    "msr3d_attn_bwd_f32": [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp],
    "msr3d_attn_bwd_tc_f32": [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp],
    "msr3d_attn_tail_fwd_f32": [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_ull, c_vp, c_float, c_uint, c_uint,
                                c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_attn_tail_bwd_f32": [c_vp, c_int, c_vp, c_vp, c_ull, c_vp, c_float, c_uint, c_uint,
                                c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_ffn_tail_fwd_f32": [c_vp, c_vp, c_int, c_vp, c_vp, c_ull, c_vp, c_float, c_uint,
                               c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_ffn_tail_bwd_f32": [c_vp, c_int, c_vp, c_ull, c_vp, c_float, c_uint, c_vp, c_vp,
                               c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_colsum_add_f32": [c_vp, c_ll, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "msr3d_dropout_mask_f32": [c_ull, c_vp, c_float, c_uint, c_ll, c_vp, c_vp],
    "msr3d_adamw_clip_step_f32": [c_vp, c_vp, c_vp, c_vp, c_ll, c_ll, c_ll, c_float, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_dot_f32": [c_vp, c_vp, c_ll, c_vp, c_vp],
    "msr3d_adamw_clip_step_zg_f32": [c_vp, c_vp, c_vp, c_vp, c_ll, c_ll, c_ll, c_float, c_vp, c_vp, c_vp, c_vp, c_int, c_vp],
    "msr3d_adamw_clip_step_split_f32": [c_vp, c_vp, c_vp, c_vp, c_ll, c_ll, c_ll, c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_qpos_fwd_acc_f32": [c_vp, c_vp, c_int] + [c_vp] * 15 + [c_vp],
    "msr3d_stage_step": [ctypes.POINTER(StageDesc), c_vp],
    "msr3d_adamw_apply_f32": [c_vp, c_vp, c_vp, c_vp, c_ll, c_ll, c_ll, c_float, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_accum_tick": [c_vp, c_vp, c_int, c_vp],
    "msr3d_zero_gated_f32": [c_vp, c_ll, c_vp, c_vp],
    "msr3d_adamw_clip_step_gated_f32": [c_vp, c_vp, c_vp, c_vp, c_ll, c_ll, c_ll, c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_allreduce_sum_f32": [ctypes.POINTER(AllreduceDesc), c_ll, c_ll, c_int, c_vp],
    "msr3d_grad_norm_finalize_f32": [c_vp, c_int, c_int, c_int, c_float, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_split_tf32_f32": [c_vp, c_vp, c_vp, c_ll, c_vp],
    "msr3d_sumsq_partials_f32": [c_vp, c_ll, c_ll, c_int, c_vp, c_int, c_vp],
    "msr3d_ffn_pair_fwd_ps_f32": [c_vp] * 8 + [c_int, c_int, c_ull, c_vp, c_float, c_uint, c_vp, c_vp, c_vp, c_int, c_vp],
    "msr3d_ffn_pair_bwd_ps_f32": [c_vp] * 7 + [c_int, c_int, c_vp, c_vp, c_int, c_vp],
    "msr3d_attn_block_fwd_ps_f32": [c_vp] * 9 + [c_int, c_int] + [c_vp] * 4 + [c_ull, c_vp, c_float, c_uint, c_uint]
                                   + [c_vp] * 11 + [c_vp],
    "msr3d_ffn_tail_bwd_ps_f32": [c_vp, c_int, c_vp, c_ull, c_vp, c_float, c_uint, c_vp, c_vp,
                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_attn_block_supported": [c_int, c_int, c_int],
    "msr3d_anchor_inputs": [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp],
    "msr3d_loc_layer_fwd_f32": [c_vp, c_int] + [c_vp] * 10 + [c_vp],
    "msr3d_cross_vec_fwd_f32": [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp],
    "msr3d_cross_residual_ln_fwd_f32": [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_ull, c_vp, c_float, c_uint, c_uint,
                                        c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "msr3d_cross_ffn_residual_fwd_f32": [c_vp, c_vp, c_int, c_ull, c_vp, c_float, c_uint, c_vp, c_vp],
    "msr3d_cross_ffn_residual_bwd_f32": [c_vp, c_int, c_ull, c_vp, c_float, c_uint, c_vp, c_vp, c_vp],
    "msr3d_cross_ln_bwd_f32": [c_vp, c_vp, c_int, c_vp, c_ull, c_vp, c_float, c_uint, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                               c_vp, c_vp],
    "msr3d_cross_vec_bwd_f32": [c_vp, c_vp, c_int, c_int] + [c_vp] * 8 + [c_vp],
    "msr3d_loc_layer_bwd_f32": [c_vp, c_vp, c_vp, c_vp, c_int] + [c_vp] * 7 + [c_vp],
    "msr3d_anchor_assemble_f32": [c_vp] * 8 + [c_int, c_int, c_vp],
    "msr3d_anchor_disassemble_f32": [c_vp, c_vp, c_int, c_int] + [c_vp] * 5 + [c_vp],
    "msr3d_situation_embed_fwd_f32": [c_vp, c_vp, c_int, c_int] + [c_vp] * 11 + [c_vp],
    "msr3d_situation_embed_bwd_f32": [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int] + [c_vp] * 9 + [c_vp],
    "msr3d_attn_fusion_fwd_f32": [c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp],
    "msr3d_attn_fusion_bwd_f32": [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    "msr3d_spatial_weights_fwd_f32": [c_vp, c_vp, c_int, c_int, c_vp, c_int, c_int, c_int, c_vp],
    "msr3d_spatial_weights_bwd_f32": [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp],
    "msr3d_attn_block_fwd_f32": [c_vp] * 9 + [c_int, c_int] + [c_vp] * 4 + [c_ull, c_vp, c_float, c_uint, c_uint]
                                + [c_vp] * 9 + [c_vp],
})

def _os_env(name, default):
    import os
    return os.environ.get(name, default)


H = 256          # hidden size the fused kernels are built for
NHEAD = 8
QKVC = 3 * H + NHEAD * 6
# spatial_attn_fusion -> `fusion` argument of msr3d_attn_fusion_{fwd,bwd}_f32 (`ctx` is `bias` with spatial weights
# derived from q: msr3d_spatial_weights_fwd_f32 ctx_mode 1)
# situation modes that put an agent ("anchor") token in front of the objects (ose3d_situation.py:334-357; csrc/modes.cu)
ANCHOR_MODES = ("as_object", "as_object_add_loc")
FUSION_CODE = {"cond": 0, "mul": 1, "bias": 2, "ctx": 2, "add": 3}


def encoder_frozen(model) -> bool:
    return all(not p.requires_grad for p in model.obj_encoder.parameters())


def supported(model) -> Optional[str]:
    """None if `model` (an OSE3DSituation) can run on the fused engine, else the reason.  An UNFROZEN object encoder
    (vision.args.freeze False) is fine for the transformer half: the module call (autograd.py) then runs the encoder's
    own training path (encoder_train.py) in front of it and receives d(embeddings) from `backward(need_d_emb=True)`;
    the engine's own optimiser schedules (train_step*, prefetch) encode without autograd and refuse it."""
    c = model.cfg
    se = c.spatial_encoder
    checks = [
        (c.hidden_size == H, "hidden_size != 256"),
        (se.num_attention_heads == NHEAD, "heads != 8"),
        (bool(c.use_spatial_attn), "use_spatial_attn is off"),
        (se.spatial_attn_fusion in FUSION_CODE and se.spatial_dim == 5, "spatial_attn_fusion / spatial_dim != 5"),
        (se.spatial_attn_fusion != "cond" or bool(se.spatial_multihead), "cond fusion without spatial_multihead"),
        (se.pairwise_rel_type == "center", "pairwise_rel_type != center"),
        (se.activation == "gelu", "activation != gelu"),
        (se.obj_loc_encoding in ("same_all", "same_0", "diff_all"), f"obj_loc_encoding {se.obj_loc_encoding}"),
        (se.obj_loc_encoding != "diff_all" or se.dim_loc == 6, "dim_loc != 6"),
        (model.situation_type in ("as_transform_for_objects", "as_embedding", "as_cross_attention") + ANCHOR_MODES,
         f"situation_type {model.situation_type}"),
        (model.situation_type not in ANCHOR_MODES or bool(c.use_anchor), "anchor mode without use_anchor"),
        (model.situation_type == "as_object" or c.loc_fourier_dim == 63, "loc_fourier_dim != 63"),
        (model.situation_type not in ("as_object", "as_cross_attention") or se.dim_loc == 6, "dim_loc != 6"),
        (not c.attn_flat.use_attn_flat, "attn_flat on"),
        (bool(c.use_orientation), "use_orientation off"),
        (c.vision.name == "PcdObjEncoder", "vision backbone is not PcdObjEncoder"),
        (se.dim_feedforward % 64 == 0, "dim_feedforward % 64"),
        (hasattr(model, "obj_encoder"), "no object encoder"),
    ]
    for ok, why in checks:
        if not ok:
            return why
    return None


class FusedPrompter:
    def __init__(self, model, lr=3e-5, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8,
                 grad_norm: Optional[float] = 5.0, world_size: int = 1, split: int = 3,
                 grad_split: int = 6, gemm_impl: Optional[str] = None, no_decay_bias: bool = False,
                 alias_grads: bool = True, chain_sms: Optional[int] = None,
                 extra_params: Optional[Dict[str, torch.nn.Parameter]] = None, accum_steps: int = 1):
        # NOTE: weight-gradient GEMMs and bias reductions are issued on a side stream (see _run)
        why = supported(model)
        if why is not None:
            raise NotImplementedError(f"fused prompter: unsupported configuration ({why})")
        self.model = model
        # weight decay: the reference decays every optimised parameter (one uniform AdamW group,
        # model/msr3d/msr3d.py:165-172); no_decay_bias=True exempts names containing "bias" (opt-in)
        self.no_decay_bias = bool(no_decay_bias)
        # alias_grads: p.grad of every trainable parameter is a view of the flat gradient buffer (engine API:
        # the fused optimiser and the all-reduce work on it).  False = module / autograd mode (autograd.py):
        # p.grad belongs to autograd, the flat buffer is scratch of one backward.
        self.alias_grads = bool(alias_grads)
        # extra_params: trainable parameters of the CALLER (llm_proj, llm_proj_img, clip_proj: model/msr3d/msr3d.py:84-130)
        # re-homed into the same flat buffers, so that the reference's ONE gradient exchange, ONE global-norm clip and
        # ONE AdamW over all trainable parameters (13.67 M elements) are this engine's single flat
        # operations.  Their forward / backward stay the caller's (torch autograd accumulates into p.grad, which is a
        # view of the flat gradient buffer); they sit in the early bucket (final before the prompter's backward).
        # accum_steps = k > 1: gradient accumulation over k micro-steps (configs/msr3d.yaml:33 gradient_accumulation_steps:
        # 5; trainer/leo_trainer.py:66,211) in the prefetch schedule — gradients are cleared at the first micro-step of a
        # window, exchanged / clipped / applied (scaled by 1 / (world * k), the mean over micro-batches as
        # accelerator.backward produces it) at the last; a device-side counter gates those kernels, so the SAME CUDA
        # graph replays for every micro-step.
        self.accum_steps = int(accum_steps)
        assert self.accum_steps >= 1
        self.extra_params = dict(extra_params or {})
        assert alias_grads or not self.extra_params, "extra_params need alias_grads=True"
        # transformer GEMMs: "tf32x3" = TMA-fed tcgen05 kind::tf32 big/small pairs (csrc/gemm3x.cu, default),
        # "bf16" = round-1 register-path split-bf16 kernel (csrc/linear_tc.cu; `split` / `grad_split` apply)
        import os as _os0
        self.gemm_impl = gemm_impl or _os0.environ.get("MSR3D_GEMM_IMPL", "tf32x3")
        assert self.gemm_impl in ("tf32x3", "bf16")
        # fused FFN pair kernels (FFN1 + GELU + dropout + FFN2 / du + dy2): on with the TMA-fed GEMMs
        self.ffn_pair = self.gemm_impl == "tf32x3" and _os0.environ.get("MSR3D_FFN_PAIR", "1") == "1"
        self.ffn_slice = int(_os0.environ.get("MSR3D_FFN_SLICE", "64"))
        # attention block of a layer (QKV+cond GEMM, attention core, out-projection, LayerNorm tail) as ONE cluster
        # kernel (csrc/attn_block.cu); needs the pre-split weight images (flat_big / flat_small)
        self.attn_block = self.gemm_impl == "tf32x3" and _os0.environ.get("MSR3D_ATTN_BLOCK", "1") == "1"
        # FFN pair fed with PRE-SPLIT operands (activation images written by the LayerNorm tails, weight images by the
        # once-per-step split): needs the attention block (producer of the y2 images)
        self.ffn_presplit = self.ffn_pair and self.attn_block and _os0.environ.get("MSR3D_FFN_PRESPLIT", "1") == "1"
        self.split = split                    # transformer GEMMs (forward + data gradients)
        self.enc_split = 3 if split == 6 else split   # frozen encoder: two-way split is ample (1.5e-5)
        self.grad_split = grad_split if split >= 3 else split
        self.world_size = world_size
        self.dev = next(model.parameters()).device
        assert self.dev.type == "cuda", "the fused engine runs on CUDA tensors only"
        se = model.cfg.spatial_encoder
        self.nl, self.F = se.num_layers, se.dim_feedforward
        self.same_all = se.obj_loc_encoding == "same_all"
        # diff_all (ose3d_situation.py:379-381): layer i adds its OWN loc_layers[i](loc) = LN(Linear_6(loc)) as the query
        # position, whatever the situation mode (the modes' Fourier / size encoders are then never evaluated)
        self.diff_all = se.obj_loc_encoding == "diff_all"
        # attention fusion (modules/layers/transformers.py:207-244): `cond` (every shipped config) has the one-launch
        # attention block and the tensor-core backward; mul / bias / ctx / add run the four-launch path with the
        # fusion-templated attention kernels and pairwise_loc_fc as spatial weights (csrc/spatial.cu)
        self.fusion = se.spatial_attn_fusion
        # situation mode: the shipped as_transform_for_objects, or one of the agent-token modes (tokens per scene = L + 1)
        self.stype = model.situation_type
        self.n_anchor = 1 if self.stype in ANCHOR_MODES else 0
        # as_embedding: one situation embedding per scene added to every object's query position (csrc/modes.cu)
        self.sit_embed = self.stype == "as_embedding"
        # as_cross_attention: a pre-norm cross-attention + FFN block on the situation vector in front of every spatial
        # layer (csrc/modes.cu: cross_* kernels; the FFN on the fused pair kernels)
        self.cross = self.stype == "as_cross_attention"
        assert not self.cross or self.ffn_pair, "as_cross_attention needs the fused FFN pair kernels (gemm_impl tf32x3)"
        self.multihead = int(bool(se.spatial_multihead))
        self.dist_norm = int(bool(se.spatial_dist_norm))
        self.p_drop = float(se.dropout)
        self._rehome()
        dev = self.dev
        self.m = torch.zeros_like(self.flat_p[:self.n_opt])
        self.v = torch.zeros_like(self.flat_p[:self.n_opt])
        # {sumsq, step, lr, beta1, beta2, eps, weight_decay, max_norm}
        # {sumsq, step, lr, beta1, beta2, eps, weight_decay, max_norm} + scratch of the norm reduction
        # (MSR3D_OPT_STATE_FLOATS in include/msr3d_b200.h)
        self.opt_state = torch.zeros(8 + 4 * 148, dtype=torch.float32, device=dev)
        self.opt_state[:8] = torch.tensor([0.0, 0.0, lr, betas[0], betas[1], eps, weight_decay,
                                           grad_norm if grad_norm else 0.0])
        self.ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        # {micro, k, clear_next, sync_now} of the accumulation window (msr3d_accum_tick)
        if self.accum_steps > 1:
            assert self.world_size == 1 or self.p2p is not None, "gradient accumulation at N > 1 needs the P2P exchange"
        self.accum_state = torch.tensor([0, self.accum_steps, 1, 1 if self.accum_steps == 1 else 0], dtype=torch.int32, device=dev)
        if self.accum_steps > 1 and self.p2p is not None:
            for d in self.p2p["desc"]:
                d.gate = self.accum_state.data_ptr() + 12
        self.loss_buf = torch.zeros(2, dtype=torch.float32, device=dev)   # double-buffered (async read-back)
        self._loss_i = 0
        self.seed_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.seed = 0x5EED
        self._shape = None
        self._keep = []          # ctypes descriptors must outlive the program
        self.graph = None
        # SM partition (msr3d_b200/partition.py): the transformer chain of batch i and the frozen encoder of batch
        # i+1 run on disjoint sets of SMs; chain_sms = 0 keeps everything on one set (round-1 behaviour)
        chain_sms = int(_os0.environ.get("MSR3D_CHAIN_SMS", "0")) if chain_sms is None else int(chain_sms)
        self.partition = None
        if chain_sms > 0:
            from . import partition as _part
            self.partition = _part.make(dev, chain_sms)
        if self.partition is not None:
            part = self.partition
            self.side_stream, self.side_stream2 = part.stream(0), part.stream(0)
            self.hi_stream = part.stream(0, priority=-1)
            self.enc_stream_part = part.stream(1)
            _C.lib().msr3d_set_tunable(7, part.counts[1])      # persistent SA kernels: one CTA per encoder SM
        else:
            self.side_stream = torch.cuda.Stream(device=dev)
            self.side_stream2 = torch.cuda.Stream(device=dev)
            self.hi_stream = torch.cuda.Stream(device=dev, priority=-1)
            self.enc_stream_part = None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self._copied = [torch.cuda.Event(), torch.cuda.Event()]      # H2D of slot s complete
        self._slot_free = [torch.cuda.Event(), torch.cuda.Event()]   # encoder that read slot s complete
        self.enc_stream = self.enc_stream_part if self.enc_stream_part is not None else torch.cuda.Stream(device=dev)
        self.use_side = True
        self.wgrad_small = _os0.environ.get("MSR3D_WGRAD_SMALL", "1") == "1"
        # data parallel: all-reduce the gradients of layers >= 1 under the backward of layer 0 (two buckets)
        # instead of one collective after the backward.  A/B: MSR3D_DDP_BUCKETS=0/1
        import os as _os
        self.ddp_buckets = _os.environ.get("MSR3D_DDP_BUCKETS", "1") == "1"
        # data parallel, prefetch schedule: the NCCL all-reduces and the optimiser are CAPTURED in the step's CUDA graph
        # (one replay per step; round 1 cut the step into two graphs with the collectives and AdamW launched from the
        # host in between: +64 us per step already at 2 GPUs).  MSR3D_DDP_IN_GRAPH=0 restores that.
        self.ddp_in_graph = _os.environ.get("MSR3D_DDP_IN_GRAPH", "1") == "1"
        self.comm_stream = torch.cuda.Stream(device=dev) if world_size > 1 else None
        # single GPU, prefetch schedule: squared-norm partials of the early gradient bucket (layers >= 1) on a side stream
        # under the backward of layer 0; the tail of the step then sums 1.4 M instead of 4.2 M gradients before AdamW
        # (the peer-memory exchange does the same at N > 1).  A/B: MSR3D_EARLY_NORM=0/1
        self.early_norm = _os.environ.get("MSR3D_EARLY_NORM", "0") == "1" and world_size == 1 and self.accum_steps == 1
        if self.early_norm:
            self.norm_stream = (self.partition.stream(0) if self.partition is not None else torch.cuda.Stream(device=dev))
            self.norm_partials = torch.zeros(2 * AR_MAX_WORLD * AR_MAX_BLOCKS, dtype=torch.float32, device=dev)
        self._side_pending = [False, False]
        self.extra_sumsq: Optional[torch.Tensor] = None   # see optimizer_step()
        # one engine per model: the drop-in module call (autograd.py) reuses this one instead of re-homing
        # the parameters a second time
        object.__setattr__(model, "_fused_engine", self)

    # ------------------------------------------------------------------ parameters
    def _rehome(self):
        from . import dist as _dist
        # the object encoder's parameters (trainable only with vision.args.freeze False) stay where they are: their
        # forward / backward is encoder_train.py's, driven by autograd
        # ... unless this engine optimises them itself (engine API, vision.args.freeze False): then the 29 PointNet++
        # parameters join the flat buffers (late bucket: their gradients are the last to be final), so the ONE gradient
        # exchange, clip and AdamW cover "the trainable encoder / adapter parameters" of the north star; the 607-way head
        # the prompter never evaluates goes to the never-touched segment
        self.train_encoder = self.alias_grads and not encoder_frozen(self.model)
        named = [(n, p) for n, p in self.model.named_parameters()
                 if p.requires_grad and (self.train_encoder or not n.startswith("obj_encoder."))]
        n_model = len(named)
        named += [("ext." + k, p) for k, p in self.extra_params.items() if p.requires_grad]
        by_name = dict(named)
        # parameters the configured situation mode never touches (neither reduced nor updated)
        # What the configured (situation mode, obj_loc_encoding) evaluates; every other trainable parameter goes to the
        # never-touched segment (what DDP's find_unused_parameters achieves in the reference).  In as_cross_attention
        # nn.MultiheadAttention's q / k projections and norm1 get exactly zero gradients (all keys are one vector) and are
        # optimised like the reference does; norm2 is never evaluated.
        st = self.stype
        used = ["obj_linear_projection.", "object_orientation_feat", "object_type_embedding.", "spatial_encoder.", "ext.",
                "obj_encoder.pcd_net."]
        if st in ANCHOR_MODES:
            used += ["anchor_feat", "orientation_encoder."]
        if self.diff_all:
            used += ["loc_layers."]
        elif st in ("as_object", "as_cross_attention"):
            used += ["loc_layers.0."]
        else:
            used += ["loc_embedding_encoder.", "size_embedding_encoder."]
        if st == "as_embedding" and not self.diff_all:
            used += ["orientation_encoder."]
        if st == "as_cross_attention":
            used += ["loc_embedding_encoder.", "orientation_encoder."]
            used += [f"situation_condition_layer.{i}.{part}" for i in range(self.nl)
                     for part in ("multihead_attn.", "linear1.", "linear2.", "norm1.", "norm3.")]
        unused = tuple(n for n, _ in named if not n.startswith(tuple(used)))
        lay = _dist.flat_layout([n for n, _ in named], {n: p.numel() for n, p in named}, self.nl,
                                no_decay_bias=self.no_decay_bias, unused=unused)
        self.order, offs, total = lay["order"], lay["offs"], lay["total"]
        self.decay_range = lay["decay"]          # [d0, d1): decayed elements (everything optimised by default)
        self.bucket_early = lay["early"]         # gradients of layers >= 1: complete before layer 0's backward
        self.bucket_late = lay["late"]
        self.n_opt = lay["n_opt"]
        total4 = (total + 3) // 4 * 4
        self.flat_p = torch.zeros(total4, dtype=torch.float32, device=self.dev)[:total]
        self.p2p = None
        if self.world_size > 1:
            self.p2p = self._symmetric_gradients(total4)
        self.flat_g = (self.p2p["grad"][:total] if self.p2p is not None
                       else torch.zeros(total, dtype=torch.float32, device=self.dev))
        # tf32 big / small operand images of the whole parameter buffer, refreshed once per step (_programs)
        self.flat_big = torch.zeros(total4, dtype=torch.float32, device=self.dev)
        self.flat_small = torch.zeros(total4, dtype=torch.float32, device=self.dev)
        self.n_split = total4
        self.offs = offs
        for n in self.order:
            p = by_name[n]
            view = self.flat_p[offs[n]:offs[n] + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            if self.alias_grads:
                p.grad = self.flat_g[offs[n]:offs[n] + p.numel()].view_as(p)
        self.by_name = by_name
        self.autograd_names = [n for n, _ in named[:n_model]]
        self.n_trainable = sum(p.numel() for _, p in named)
        if self.world_size > 1:
            _dist.broadcast_(self.flat_p)     # identical replicas, as DDP's constructor guarantees

    def _symmetric_gradients(self, n):
        """Data parallel: the flat gradient buffer (and the flags / norm partials of csrc/allreduce.cu) in SYMMETRIC
        memory, so that the all-reduce is this library's own NVLink peer-memory kernel instead of ncclAllReduce.
        None (NCCL stays the exchange) when MSR3D_P2P_ALLREDUCE=0 or symmetric memory is not available."""
        import os
        import warnings
        if os.environ.get("MSR3D_P2P_ALLREDUCE", "1") != "1" or self.world_size > AR_MAX_WORLD:
            return None
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            group = dist.group.WORLD
            enable = getattr(symm_mem, "enable_symm_mem_for_group", None)
            if enable is not None:
                try:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        enable(group.group_name)
                except Exception:  # noqa: BLE001
                    pass
            out = {}
            for key, numel, dtype in (("grad", n, torch.float32), ("flags", 2 * AR_MAX_BLOCKS * AR_MAX_WORLD * 2, torch.int32),
                                      ("partials", 2 * AR_MAX_WORLD * AR_MAX_BLOCKS, torch.float32)):
                t = symm_mem.empty(numel, dtype=dtype, device=self.dev)
                t.zero_()
                hdl = symm_mem.rendezvous(t, group)
                out[key], out[key + "_ptrs"], out[key + "_hdl"] = t, [int(x) for x in hdl.buffer_ptrs], hdl
            out["epoch"] = torch.zeros(2 * AR_MAX_BLOCKS, dtype=torch.int32, device=self.dev)
            out["rank"] = dist.get_rank()
            torch.cuda.synchronize()
            dist.barrier()                      # every rank's flags are zero before anybody launches
            descs = []
            for ch in (0, 1):
                d = AllreduceDesc()
                for q in range(self.world_size):
                    d.buf[q], d.flags[q], d.partials[q] = out["grad_ptrs"][q], out["flags_ptrs"][q], out["partials_ptrs"][q]
                d.epoch, d.rank, d.world, d.channel = out["epoch"].data_ptr(), out["rank"], self.world_size, ch
                d.gate = None          # set by __init__ when gradients are accumulated over several micro-steps
                descs.append(d)
            out["desc"] = descs
            return out
        except Exception as e:  # noqa: BLE001
            warnings.warn(f"msr3d_b200: symmetric memory unavailable ({type(e).__name__}: {e}); gradients are exchanged with NCCL")
            return None

    def _comm_waits_for_gradients(self, main):
        """The early bucket is final once the main chain AND the weight-gradient side streams have got this far: the
        COMMUNICATION stream waits for them — the main chain does not (joining it here kept layer 0's backward waiting
        for the weight-gradient GEMMs of layers >= 1: +45 us per step at every N > 1)."""
        self.comm_stream.wait_stream(main)
        for k, side in enumerate((self.side_stream, self.side_stream2)):
            if self._side_pending[k]:
                self.comm_stream.wait_stream(side)

    def _backward_optimizer_early_norm(self, main, split):
        """Single-GPU tail of the prefetch step with the early bucket's norm taken beside the backward of layer 0."""
        (e0, e1), (l0, l1) = self.bucket_early, self.bucket_late
        nb0, nb1 = 48, 32
        ns, st = self.norm_stream, self.opt_state
        self._run(self.bwd_prog[:self.bwd_split])
        ns.wait_stream(main)
        for k, side in enumerate((self.side_stream, self.side_stream2)):
            if self._side_pending[k]:
                ns.wait_stream(side)
        _C.call("msr3d_sumsq_partials_f32", self.flat_g.data_ptr(), e0, e1, nb0, self.norm_partials.data_ptr(), 0, ns.cuda_stream)
        self._run(self.bwd_prog[self.bwd_split:])          # ends with a join of the weight-gradient streams
        _C.call("msr3d_sumsq_partials_f32", self.flat_g.data_ptr(), l0, l1, nb1, self.norm_partials.data_ptr(), 1, main.cuda_stream)
        main.wait_stream(ns)
        ex = self.extra_sumsq.data_ptr() if self.extra_sumsq is not None else None
        _C.call("msr3d_grad_norm_finalize_f32", self.norm_partials.data_ptr(), 1, nb0, nb1, 1.0, ex, st.data_ptr(),
                self.seed_counter.data_ptr(), None, main.cuda_stream)
        n, (d0, d1) = self.n_opt, self.decay_range
        _C.call("msr3d_adamw_apply_f32", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                n, d0, d1, 1.0, st.data_ptr(), self.flat_big.data_ptr() if split else None,
                self.flat_small.data_ptr() if split else None, None, main.cuda_stream)

    def _allreduce_p2p(self, channel, begin, end, nblocks, stream):
        rc = _C.lib().msr3d_allreduce_sum_f32(ctypes.byref(self.p2p["desc"][channel]), begin, end, nblocks, stream.cuda_stream)
        if rc != 0:
            raise _C.Msr3dError(f"allreduce_sum failed ({rc}): {_C.lib().msr3d_last_error().decode()}")

    def _optimizer_after_p2p(self, nb0, nb1, split):
        """Clip + AdamW on gradients reduced by _allreduce_p2p: the norm partials came with the data."""
        st = torch.cuda.current_stream().cuda_stream
        ex = self.extra_sumsq.data_ptr() if self.extra_sumsq is not None else None
        acc = self.accum_steps > 1
        gate = self.accum_state.data_ptr() + 12 if acc else None        # sync_now
        scale = 1.0 / (self.world_size * self.accum_steps)
        _C.call("msr3d_grad_norm_finalize_f32", self.p2p["partials"].data_ptr(), self.world_size, nb0, nb1,
                scale, ex, self.opt_state.data_ptr(), None if acc else self.seed_counter.data_ptr(), gate, st)
        n, (d0, d1) = self.n_opt, self.decay_range
        _C.call("msr3d_adamw_apply_f32", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                n, d0, d1, scale, self.opt_state.data_ptr(),
                self.flat_big.data_ptr() if split else None, self.flat_small.data_ptr() if split else None, gate, st)

    def _accum_tick(self, phase):
        if self.accum_steps > 1:
            _C.call("msr3d_accum_tick", self.accum_state.data_ptr(), self.seed_counter.data_ptr(), phase,
                    torch.cuda.current_stream().cuda_stream)

    def _optimizer_accum(self, split):
        """Single-GPU clip + AdamW of an accumulation window: norm and update gated by sync_now."""
        n, (d0, d1) = self.n_opt, self.decay_range
        ex = self.extra_sumsq.data_ptr() if self.extra_sumsq is not None else None
        _C.call("msr3d_adamw_clip_step_gated_f32", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(),
                self.v.data_ptr(), n, d0, d1, 1.0 / (self.world_size * self.accum_steps), self.opt_state.data_ptr(),
                self.ticket.data_ptr(), ex, self.flat_big.data_ptr() if split else None,
                self.flat_small.data_ptr() if split else None, self.accum_state.data_ptr() + 12,
                torch.cuda.current_stream().cuda_stream)

    def P(self, name):  # device pointer of a parameter / its gradient
        return self.flat_p.data_ptr() + 4 * self.offs[name]

    def G(self, name):
        return self.flat_g.data_ptr() + 4 * self.offs[name]

    def PB(self, name):  # big / small tf32 images of a parameter (msr3d_split_tf32_f32)
        return self.flat_big.data_ptr() + 4 * self.offs[name]

    def PS(self, name):
        return self.flat_small.data_ptr() + 4 * self.offs[name]

    # ------------------------------------------------------------------ workspaces + program
    def _buf(self, *shape, dtype=torch.float32, zero=False):
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.dev)
        return t

    def _build(self, B, L, O_pts):
        """Allocate static buffers and pre-bind every launch for (B scenes, L objects, P points)."""
        Lo = L                       # objects per scene; L = tokens per scene from here on
        L = Lo + self.n_anchor
        M, F, nl = B * L, self.F, self.nl
        f = self._buf
        w = self.ws = {}
        w["obj_fts"] = f(B, Lo, O_pts, 6)
        w["obj_locs"] = f(B, Lo, 6)
        w["obj_masks"] = torch.ones(B, Lo, dtype=torch.bool, device=self.dev)
        w["key_pad"] = torch.zeros(B, L, dtype=torch.uint8, device=self.dev)
        if self.n_anchor:
            w["all_loc"] = f(B, L, 6)
            w["proj"], w["dproj"] = f(B * Lo, H), f(B * Lo, H)
        if self.sit_embed:
            w["xhat_sit"], w["rstd_sit"] = f(B, H), f(B)
        if self.diff_all:
            w["qpos_l"] = [f(M, H) for _ in range(nl)]
            w["xhat_q"] = [f(M, H) for _ in range(nl)]
            w["rstd_q"] = [f(M) for _ in range(nl)]
            w["x0_scratch"] = f(M, H)
        w["anchor_locs"] = f(B, 3)
        w["anchor_ori"] = f(B, 4)
        w["upstream"] = f(M, H)
        w["pw"], w["ff"], w["sz"] = f(B, L, L, 5), f(M, 64), f(M, 3)
        # prefetch schedule: the geometry of the batch being encoded is computed beside the encoder (it depends on
        # the inputs only) and handed to the next step
        w["pw_n"], w["ff_n"], w["sz_n"] = f(B, L, L, 5), f(M, 64), f(M, 3)
        w["h1"], w["qpos"], w["qpos0"] = f(M, H), f(M, H), f(M, H)
        for k in ("xhat_l", "xhat_s"):
            w[k] = f(M, H)
        for k in ("rstd_l", "rstd_s"):
            w[k] = f(M)
        # split-K GEMM outputs (accumulated with red.add): carved out of two arenas that are zeroed with
        # one memset each per step (forward: x_in[0], qkvc, o, f of every layer; backward: dctx)
        cross = self.cross
        n_f = M * H + nl * (M * QKVC + 2 * M * H) + (nl * M * H if cross else 0)
        w["arena_f"] = f(n_f)
        w["arena_b"] = f(nl * M * H + ((nl * M * H + B * H) if cross else 0))
        cur = [0]

        def carve(arena, *shape):
            n = 1
            for d in shape:
                n *= d
            t = arena[cur[0]:cur[0] + n].view(*shape)
            cur[0] += n
            return t
        if cross:
            # x_pre[i]: input of layer i's cross-attention block (i = 0: the assembled layer-0 input, reduced into a
            # zeroed buffer); x_in[i]: its output = input of spatial layer i
            w["x_pre"] = [carve(w["arena_f"], M, H)] + [f(M, H) for _ in range(nl - 1)]
            w["x_in"] = [f(M, H) for _ in range(nl)]
        else:
            w["x_in"] = [carve(w["arena_f"], M, H)] + [f(M, H) for _ in range(nl - 1)]
        w["qkvc"] = [carve(w["arena_f"], M, QKVC) for _ in range(nl)]
        w["o"] = [carve(w["arena_f"], M, H) for _ in range(nl)]
        w["f"] = [carve(w["arena_f"], M, H) for _ in range(nl)]
        if cross:
            w["f_c"] = [carve(w["arena_f"], M, H) for _ in range(nl)]
        assert cur[0] == n_f
        cur[0] = 0
        w["dctx"] = [carve(w["arena_b"], M, H) for _ in range(nl)]
        if cross:
            w["dt3"] = [carve(w["arena_b"], M, H) for _ in range(nl)]
            w["ds_sit"] = carve(w["arena_b"], B, H)
            w["s_sit"], w["xhat_sit"], w["rstd_sit"] = f(B, H), f(B, H), f(B)
            w["v_c"] = [f(B, H) for _ in range(nl)]
            w["P_c"] = [f(B, NHEAD, H) for _ in range(nl)]
            w["scale_c"] = [f(M, NHEAD) for _ in range(nl)]
            for k in ("xa", "xhat3", "t3", "df_c", "dxa"):
                w[k] = [f(M, H) for _ in range(nl)]
            w["rstd3"] = [f(M) for _ in range(nl)]
            w["h_c"] = [f(M, F) for _ in range(nl)]
            w["gp_c"] = [f(M, F) for _ in range(nl)]
            w["du_c"] = [f(M, F) for _ in range(nl)]
            w["g_attn"] = f(M, H)
        w["P"] = [f(B, NHEAD, L, L) for _ in range(nl)]
        w["LOC"] = [f(B, NHEAD, L, L) for _ in range(nl)]
        w["ctx"] = [f(M, H) for _ in range(nl)]
        for k in ("xhat_a", "xhat_1", "xhat_2", "y2"):
            w[k] = [f(M, H) for _ in range(nl)]
        for k in ("rstd_a", "rstd_1", "rstd_2"):
            w[k] = [f(M) for _ in range(nl)]
        w["h"] = [f(M, F) for _ in range(nl)]     # dropout(gelu(u))
        w["gp"] = [f(M, F) for _ in range(nl)]    # gelu'(u) * dropout mask
        w["tokens"] = f(M, H)
        # operand images (big / small tf32) of y2 and df for the pre-split FFN pair: live inside one layer only
        w["y2_big"], w["y2_small"], w["df_big"], w["df_small"] = f(M, H), f(M, H), f(M, H), f(M, H)
        # backward
        # per-layer copies of everything a weight-gradient GEMM reads: those GEMMs run on a side
        # stream, off the critical path, while the main stream moves on to the next layer
        w["dy2"], w["dh1"] = f(M, H), f(M, H)
        w["df"] = [f(M, H) for _ in range(nl)]
        w["do"] = [f(M, H) for _ in range(nl)]
        w["du"] = [f(M, F) for _ in range(nl)]
        w["dqkvc"] = [f(M, QKVC) for _ in range(nl)]
        w["dx"] = [f(M, H) for _ in range(nl)]
        # pipelined mode: the frozen encoder of the NEXT batch overlaps the transformer step of the current one
        w["emb_next"] = None   # allocated in prepare() once the embedding width is known
        w["obj_locs_n"], w["anchor_locs_n"], w["anchor_ori_n"] = f(B, Lo, 6), f(B, 3), f(B, 4)
        w["obj_masks_n"] = torch.ones(B, Lo, dtype=torch.bool, device=self.dev)
        w["upstream_n"] = f(M, H)
        # prefetch mode (two batches in flight ahead of the transformer step): a second point-cloud slot
        # filled by an asynchronous H2D copy, and a second level of staged per-scene inputs
        w["obj_fts_b"] = f(B, Lo, O_pts, 6)
        for q in ("_q0", "_q1"):    # one per slot: the copy stream fills one while the step reads the other
            w["obj_locs" + q], w["anchor_locs" + q], w["anchor_ori" + q] = f(B, Lo, 6), f(B, 3), f(B, 4)
            w["obj_masks" + q] = torch.ones(B, Lo, dtype=torch.bool, device=self.dev)
            w["upstream" + q] = f(M, H)
        self._shape = (B, Lo, O_pts)
        self.graph = None
        self.graph_pipelined = None
        self.graph_prefetch = [None, None]
        self._primed = False
        self._primed2 = False

    def _gemm(self, prog, A, sa, Bm, sb, M, N, K, out, ldo, bias=0, residual=0, ldr=0, aux=0, ldaux=0,
              act=0, accumulate=0, epilogue=0, ksplit=1, a_xform=0, b_xform=0, a_site=0, b_site=0,
              epi_site=0, tag=None, wgrad=False, aux_out=0, side=False, tile_n=0):
        d = GemmDesc()
        d.A, d.stride_am, d.stride_ak = A, sa[0], sa[1]
        d.B, d.stride_bn, d.stride_bk = Bm, sb[0], sb[1]
        d.bias, d.residual, d.ld_residual, d.aux, d.ld_aux = bias or None, residual or None, ldr, aux or None, ldaux
        d.aux_out = aux_out or None
        d.out, d.ld_out, d.M, d.N, d.K = out, ldo, M, N, K
        d.act, d.accumulate, d.epilogue, d.ksplit = act, accumulate, epilogue, ksplit
        d.tile_n = tile_n
        if wgrad and side and self.wgrad_small:
            # weight gradients run beside the critical chain: few large CTAs, no split-K, so that the
            # chain's 100+-CTA GEMMs still find free SMs (each CTA of this kernel owns a whole SM)
            d.tile_n, d.ksplit = 128, 1
        # weight gradients contract over tokens and cancel heavily: three-way split (fp32-grade)
        d.split = self.grad_split if wgrad else self.split
        d.a_xform, d.b_xform, d.a_site, d.b_site, d.epi_site = a_xform, b_xform, a_site, b_site, epi_site
        d.dropout_p = self._p
        d.seed = self.seed
        d.seed_ptr = self.seed_counter.data_ptr()
        self._keep.append(d)
        ref = ctypes.byref(d)
        if self.partition is not None and not (wgrad and side) and d.ksplit > 1:
            # the chain owns `counts[0]` SMs: a GEMM wider than that would run in two waves
            bn = d.tile_n if d.tile_n in (64, 128) else 64
            tiles = -(-M // 128) * -(-N // bn)
            while d.ksplit > 1 and tiles * d.ksplit > self.partition.counts[0]:
                d.ksplit -= 1
            d.tile_n = bn
        if self.gemm_impl == "tf32x3" and _C.lib().msr3d_gemm_tf32x3_supported(ref):
            # TMA-fed 3xTF32 kernel (csrc/gemm3x.cu); the weights of the forward / data-gradient GEMMs are
            # "static" operands (written by the optimiser of the previous step only)
            if wgrad and side and self.wgrad_small:
                d.tile_n = 128
            fn = _C.lib().msr3d_gemm_tf32x3_f32
            prog.append(("msr3d_gemm_tf32x3_f32", fn, (ref, 0 if wgrad else 1), tag, side))
            return
        fn = _C.lib().msr3d_gemm_tc_f32
        prog.append(("msr3d_gemm_tc_f32", fn, (ref,), tag, side))

    def _call(self, prog, name, *args, tag=None, side=False):
        prog.append((name, getattr(_C.lib(), name), args, tag, side))

    def _programs(self, training: bool):
        """(forward, backward) launch lists for the current static buffers."""
        B, Lo, _ = self._shape
        L = Lo + self.n_anchor           # tokens per scene
        M, Mo, F, nl, w = B * L, B * Lo, self.F, self.nl, self.ws
        anchor, as_object, cross = bool(self.n_anchor), self.stype == "as_object", self.cross
        x0_buf = w["x_pre"][0] if cross else w["x_in"][0]        # the zero-filled buffer the layer-0 input is reduced into
        self._p = self.p_drop if training else 0.0
        p, seed, sptr = self._p, self.seed, self.seed_counter.data_ptr()
        P, G = self.P, self.G
        ptr = lambda t: t.data_ptr()  # noqa: E731
        pre_prog: List = []
        fwd: List = []
        bwd: List = []
        # ---------------- forward.  geometry -> loc_fc -> qpos depend only on the inputs: side stream,
        # concurrent with the PointNet++ encoder on the main stream
        # split-K factors: tools/gemm_sweep.py (one 64-wide K chunk per CTA wherever the epilogue is linear)
        ks_h = max(1, -(-H // 64))
        pre_prog.append(("memset", None, (w["arena_f"],), None, False))
        geo: List = []
        if anchor:
            # agent-token modes: all_loc / key_pad with the agent in row 0, geometry of the raw centres (no agent frame)
            self._call(geo, "msr3d_anchor_inputs", ptr(w["obj_locs"]), ptr(w["obj_masks"]), ptr(w["anchor_locs"]),
                       self.model.anchor_size.data_ptr(), B, Lo, ptr(w["all_loc"]), ptr(w["key_pad"]), tag="anchor_inputs",
                       side=True)
            self._call(geo, "msr3d_spatial_geometry_f32", ptr(w["all_loc"]), ptr(w["anchor_locs"]), ptr(w["anchor_ori"]),
                       B, L, self.dist_norm, 0, ptr(w["pw"]), ptr(w["ff"]), ptr(w["sz"]), tag="geometry", side=True)
        else:
            # as_embedding: Fourier features of the raw centres (the situation enters through its own embedding)
            self._call(geo, "msr3d_spatial_geometry_f32", ptr(w["obj_locs"]), ptr(w["anchor_locs"]), ptr(w["anchor_ori"]),
                       B, L, self.dist_norm, 0 if self.sit_embed else 1, ptr(w["pw"]), ptr(w["ff"]), ptr(w["sz"]),
                       tag="geometry", side=True)
        # prefetch schedule: geometry of the batch in point-cloud slot s, from its freshly copied inputs, for the NEXT step
        self.geo_prog = []
        for q in (() if (anchor or self.sit_embed or cross or self.diff_all) else ("_q0", "_q1")):
            gq: List = []
            self._call(gq, "msr3d_spatial_geometry_f32", ptr(w["obj_locs" + q]), ptr(w["anchor_locs" + q]), ptr(w["anchor_ori" + q]),
                       B, L, self.dist_norm, 1, ptr(w["pw_n"]), ptr(w["ff_n"]), ptr(w["sz_n"]), tag="geometry")
            self.geo_prog.append(gq)
        diff_all = self.diff_all
        store_qpos = as_object or cross or diff_all      # qpos is STORED into the layer-0 input: the projection follows it
        if diff_all:
            # loc_layers[i](loc) for every layer (ose3d_situation.py:379-381); layer 0's also starts the layer-0 input
            for i in range(nl):
                first = i == 0
                self._call(pre_prog, "msr3d_loc_layer_fwd_f32", ptr(w["all_loc"] if anchor else w["obj_locs"]), M,
                           P(f"loc_layers.{i}.0.weight"), P(f"loc_layers.{i}.0.bias"), P(f"loc_layers.{i}.1.weight"),
                           P(f"loc_layers.{i}.1.bias"), ptr(w["xhat_q"][i]), ptr(w["rstd_q"][i]), ptr(w["qpos_l"][i]),
                           ptr(x0_buf if first else w["x0_scratch"]),
                           P("object_orientation_feat") if (first and not anchor) else None,
                           P("object_type_embedding.weight") if (first and not anchor) else None, tag="qpos", side=True)
            if cross:
                self._call(pre_prog, "msr3d_situation_embed_fwd_f32", ptr(w["anchor_locs"]), ptr(w["anchor_ori"]), B, L,
                           P("loc_embedding_encoder.0.weight"), P("loc_embedding_encoder.0.bias"),
                           P("loc_embedding_encoder.1.weight"), P("loc_embedding_encoder.1.bias"),
                           P("orientation_encoder.weight"), P("orientation_encoder.bias"), ptr(w["xhat_sit"]),
                           ptr(w["rstd_sit"]), None, None, ptr(w["s_sit"]), tag="situation_vec", side=True)
        elif as_object or cross:
            # qpos = loc_layers[0](loc) (ose3d_situation.py:396), stored as the first term of the layer-0 input; without an
            # agent row (as_cross_attention) the objects' constant terms ride along
            self._call(pre_prog, "msr3d_loc_layer_fwd_f32", ptr(w["all_loc"] if as_object else w["obj_locs"]), M,
                       P("loc_layers.0.0.weight"), P("loc_layers.0.0.bias"), P("loc_layers.0.1.weight"),
                       P("loc_layers.0.1.bias"), ptr(w["xhat_l"]), ptr(w["rstd_l"]), ptr(w["qpos"]), ptr(x0_buf),
                       P("object_orientation_feat") if cross else None, P("object_type_embedding.weight") if cross else None,
                       tag="qpos", side=True)
            if cross:       # the situation vector s_b (B, 256): keys and values of every cross-attention layer
                self._call(pre_prog, "msr3d_situation_embed_fwd_f32", ptr(w["anchor_locs"]), ptr(w["anchor_ori"]), B, L,
                           P("loc_embedding_encoder.0.weight"), P("loc_embedding_encoder.0.bias"),
                           P("loc_embedding_encoder.1.weight"), P("loc_embedding_encoder.1.bias"),
                           P("orientation_encoder.weight"), P("orientation_encoder.bias"), ptr(w["xhat_sit"]),
                           ptr(w["rstd_sit"]), None, None, ptr(w["s_sit"]), tag="situation_vec", side=True)
        else:
            self._gemm(pre_prog, ptr(w["ff"]), (64, 1), P("loc_embedding_encoder.0.weight"), (63, 1), M, H, 63, ptr(w["h1"]), H,
                       bias=P("loc_embedding_encoder.0.bias"), tag="loc_fc", side=True)
            # qpos0 is reduced into the (zero-filled) layer-0 input together with the split-K partial sums of the projection
            # (agent-token modes: qpos alone; the other terms differ per row and come from anchor_assemble)
            self._call(pre_prog, "msr3d_qpos_fwd_acc_f32", ptr(w["h1"]), ptr(w["sz"]), M, P("size_embedding_encoder.0.weight"),
                       P("size_embedding_encoder.0.bias"), P("loc_embedding_encoder.1.weight"), P("loc_embedding_encoder.1.bias"),
                       P("size_embedding_encoder.1.weight"), P("size_embedding_encoder.1.bias"),
                       None if anchor else P("object_orientation_feat"), None if anchor else P("object_type_embedding.weight"),
                       ptr(w["xhat_l"]), ptr(w["rstd_l"]), ptr(w["xhat_s"]), ptr(w["rstd_s"]), ptr(w["qpos"]), ptr(w["qpos0"]),
                       ptr(w["x_in"][0]), tag="qpos", side=True)
            if self.sit_embed:
                self._call(pre_prog, "msr3d_situation_embed_fwd_f32", ptr(w["anchor_locs"]), ptr(w["anchor_ori"]), B, L,
                           P("loc_embedding_encoder.0.weight"), P("loc_embedding_encoder.0.bias"),
                           P("loc_embedding_encoder.1.weight"), P("loc_embedding_encoder.1.bias"),
                           P("orientation_encoder.weight"), P("orientation_encoder.bias"), ptr(w["xhat_sit"]),
                           ptr(w["rstd_sit"]), ptr(w["qpos"]), ptr(w["x_in"][0]), None, tag="situation_embed", side=True)
        pre_prog[1:1] = geo                              # geometry first on the side stream
        assert pre_prog[0][0] == "memset"
        cond = self.fusion == "cond"
        fcode, NQ = FUSION_CODE[self.fusion], (QKVC if cond else 3 * H)
        use_ab = cond and self.attn_block and bool(_C.lib().msr3d_attn_block_supported(L, NHEAD, H))
        self.attn_block_active = use_ab
        if use_ab:
            # operand images of the weights: once per step, beside the input encoders
            self._call(pre_prog, "msr3d_split_tf32_f32", self.flat_p.data_ptr(), self.flat_big.data_ptr(),
                       self.flat_small.data_ptr(), self.n_split, tag="split_w", side=2)
        E = self.emb_dim
        # the projection runs beside the query-position kernels (side stream); both reduce into x_in[0]
        if anchor:
            self._gemm(fwd, self.emb_ptr, (E, 1), P("obj_linear_projection.weight"), (E, 1), Mo, H, E, ptr(w["proj"]), H,
                       bias=P("obj_linear_projection.bias"), tag="obj_proj")
            fwd.append(("join", None, (), None, False))
            self._call(fwd, "msr3d_anchor_assemble_f32", ptr(w["x_in"][0]), ptr(w["proj"]), P("object_orientation_feat"),
                       P("object_type_embedding.weight"), P("anchor_feat"), P("orientation_encoder.weight"),
                       P("orientation_encoder.bias"), ptr(w["anchor_ori"]), B, Lo, tag="anchor_assemble")
        else:
            if store_qpos:  # qpos is STORED into the layer-0 input (loc_layer_fwd): the projection reduces into it afterwards
                fwd.append(("join", None, (), None, False))
            self._gemm(fwd, self.emb_ptr, (E, 1), P("obj_linear_projection.weight"), (E, 1), M, H, E, ptr(x0_buf), H,
                       bias=P("obj_linear_projection.bias"), accumulate=2, ksplit=max(1, -(-E // 64)), tag="obj_proj")
            fwd.append(("join", None, (), None, False))
        for i in range(nl):
            pre = f"spatial_encoder.{i}."
            sa = pre + "self_attn."
            s0 = i * 8
            if cross:
                # situation_condition_layer[i] (ose3d_situation.py:411-418): x_pre[i] -> x_in[i]
                pc = f"situation_condition_layer.{i}."
                Wv = P(pc + "multihead_attn.in_proj_weight") + 4 * 2 * H * H
                bv = P(pc + "multihead_attn.in_proj_bias") + 4 * 2 * H
                self._call(fwd, "msr3d_cross_vec_fwd_f32", ptr(w["s_sit"]), Wv, bv, P(pc + "multihead_attn.out_proj.weight"), B,
                           ptr(w["v_c"][i]), ptr(w["P_c"][i]), tag="cross_vec")
                self._call(fwd, "msr3d_cross_residual_ln_fwd_f32", ptr(w["x_pre"][i]), ptr(w["P_c"][i]),
                           P(pc + "multihead_attn.out_proj.bias"), M, L, P(pc + "norm3.weight"), P(pc + "norm3.bias"), seed, sptr,
                           p, s0 + 4, s0 + 5, ptr(w["scale_c"][i]), ptr(w["xa"][i]), ptr(w["xhat3"][i]), ptr(w["rstd3"][i]),
                           ptr(w["t3"][i]), tag="cross_ln")
                self._call(fwd, "msr3d_ffn_pair_fwd_f32", ptr(w["t3"][i]), P(pc + "linear1.weight"), P(pc + "linear1.bias"),
                           P(pc + "linear2.weight"), P(pc + "linear2.bias"), M, F, seed, sptr, p, s0 + 6, ptr(w["h_c"][i]),
                           ptr(w["gp_c"][i]), ptr(w["f_c"][i]), self.ffn_slice, tag="cross_ffn_pair")
                self._call(fwd, "msr3d_cross_ffn_residual_fwd_f32", ptr(w["xa"][i]), ptr(w["f_c"][i]), M, seed, sptr, p, s0 + 7,
                           ptr(w["x_in"][i]), tag="cross_out")
            x_in = ptr(w["x_in"][i])
            use_ps = use_ab and self.ffn_presplit
            if use_ab and use_ps:
                self._call(fwd, "msr3d_attn_block_fwd_ps_f32", x_in, self.PB(sa + "w_qs.weight"), self.PS(sa + "w_qs.weight"),
                           P(sa + "w_qs.bias"), self.PB(sa + "fc.weight"), self.PS(sa + "fc.weight"), P(sa + "fc.bias"),
                           ptr(w["pw"]), ptr(w["key_pad"]), B, L, P(sa + "layer_norm.weight"), P(sa + "layer_norm.bias"),
                           P(pre + "norm1.weight"), P(pre + "norm1.bias"), seed, sptr, p, s0 + 0, s0 + 1,
                           ptr(w["qkvc"][i]), ptr(w["P"][i]), ptr(w["LOC"][i]), ptr(w["ctx"][i]),
                           ptr(w["xhat_a"][i]), ptr(w["rstd_a"][i]), ptr(w["xhat_1"][i]), ptr(w["rstd_1"][i]),
                           ptr(w["y2"][i]), ptr(w["y2_big"]), ptr(w["y2_small"]), tag="attn_block")
            elif use_ab:
                self._call(fwd, "msr3d_attn_block_fwd_f32", x_in, self.PB(sa + "w_qs.weight"), self.PS(sa + "w_qs.weight"),
                           P(sa + "w_qs.bias"), self.PB(sa + "fc.weight"), self.PS(sa + "fc.weight"), P(sa + "fc.bias"),
                           ptr(w["pw"]), ptr(w["key_pad"]), B, L, P(sa + "layer_norm.weight"), P(sa + "layer_norm.bias"),
                           P(pre + "norm1.weight"), P(pre + "norm1.bias"), seed, sptr, p, s0 + 0, s0 + 1,
                           ptr(w["qkvc"][i]), ptr(w["P"][i]), ptr(w["LOC"][i]), ptr(w["ctx"][i]),
                           ptr(w["xhat_a"][i]), ptr(w["rstd_a"][i]), ptr(w["xhat_1"][i]), ptr(w["rstd_1"][i]),
                           ptr(w["y2"][i]), tag="attn_block")
            else:
                self._gemm(fwd, x_in, (H, 1), P(sa + "w_qs.weight"), (H, 1), M, NQ, H, ptr(w["qkvc"][i]), QKVC,
                           bias=P(sa + "w_qs.bias"), accumulate=2, ksplit=ks_h, tag="qkvc")
                if not cond:      # pairwise_loc_fc -> the spatial-weight columns of the qkvc rows
                    self._call(fwd, "msr3d_spatial_weights_fwd_f32", P(sa + "pairwise_loc_fc.weight"),
                               P(sa + "pairwise_loc_fc.bias"), int(self.fusion == "ctx"), self.multihead,
                               ptr(w["qkvc"][i]), QKVC, M, NHEAD, tag="spatial_w")
                self._call(fwd, "msr3d_attn_fusion_fwd_f32", ptr(w["qkvc"][i]), QKVC, ptr(w["pw"]), ptr(w["key_pad"]), B, L,
                           NHEAD, fcode, ptr(w["P"][i]), ptr(w["LOC"][i]), ptr(w["ctx"][i]), H, tag="attn_core")
                self._gemm(fwd, ptr(w["ctx"][i]), (H, 1), P(sa + "fc.weight"), (H, 1), M, H, H, ptr(w["o"][i]), H,
                           bias=P(sa + "fc.bias"), accumulate=2, ksplit=ks_h, tag="out_proj")
                self._call(fwd, "msr3d_attn_tail_fwd_f32", ptr(w["o"][i]), x_in, M, P(sa + "layer_norm.weight"),
                           P(sa + "layer_norm.bias"), P(pre + "norm1.weight"), P(pre + "norm1.bias"), seed, sptr, p,
                           s0 + 0, s0 + 1, ptr(w["xhat_a"][i]), ptr(w["rstd_a"][i]), ptr(w["xhat_1"][i]),
                           ptr(w["rstd_1"][i]), ptr(w["y2"][i]), tag="attn_tail")
            # FFN: h = dropout(gelu(u)) and gelu'(u)*mask are materialised once (the backward reads both)
            if self.ffn_pair and use_ps:
                self._call(fwd, "msr3d_ffn_pair_fwd_ps_f32", ptr(w["y2_big"]), ptr(w["y2_small"]),
                           self.PB(pre + "linear1.weight"), self.PS(pre + "linear1.weight"), P(pre + "linear1.bias"),
                           self.PB(pre + "linear2.weight"), self.PS(pre + "linear2.weight"), P(pre + "linear2.bias"), M, F,
                           seed, sptr, p, s0 + 2, ptr(w["h"][i]), ptr(w["gp"][i]), ptr(w["f"][i]),
                           self.ffn_slice, tag="ffn_pair")
            elif self.ffn_pair:
                # both GEMMs in one kernel, the hidden tile stays in shared memory (csrc/gemm3x.cu)
                self._call(fwd, "msr3d_ffn_pair_fwd_f32", ptr(w["y2"][i]), P(pre + "linear1.weight"),
                           P(pre + "linear1.bias"), P(pre + "linear2.weight"), P(pre + "linear2.bias"), M, F,
                           seed, sptr, p, s0 + 2, ptr(w["h"][i]), ptr(w["gp"][i]), ptr(w["f"][i]),
                           self.ffn_slice, tag="ffn_pair")
            else:
                self._gemm(fwd, ptr(w["y2"][i]), (H, 1), P(pre + "linear1.weight"), (H, 1), M, F, H, ptr(w["h"][i]), F,
                           bias=P(pre + "linear1.bias"), epilogue=2, aux_out=ptr(w["gp"][i]), ldaux=F,
                           epi_site=s0 + 2, tag="ffn1")
                self._gemm(fwd, ptr(w["h"][i]), (F, 1), P(pre + "linear2.weight"), (F, 1), M, H, F, ptr(w["f"][i]), H,
                           bias=P(pre + "linear2.bias"), accumulate=2, ksplit=16, tag="ffn2")
            last = i == nl - 1
            add = 0 if (last or not (self.same_all or diff_all)) else ptr(w["qpos_l"][i + 1] if diff_all else w["qpos"])
            self._call(fwd, "msr3d_ffn_tail_fwd_f32", ptr(w["f"][i]), ptr(w["y2"][i]), M, P(pre + "norm2.weight"),
                       P(pre + "norm2.bias"), seed, sptr, p, s0 + 3, add or None, ptr(w["xhat_2"][i]),
                       ptr(w["rstd_2"][i]), ptr(w["tokens"]) if last else None,
                       None if last else ptr((w["x_pre"] if cross else w["x_in"])[i + 1]), tag="ffn_tail")
        # ---------------- backward (gradients accumulate into the flat buffer)
        dy3 = ptr(w["upstream"])
        bwd.append(("memset", None, (w["arena_b"],), None, False))
        for i in reversed(range(nl)):
            if i == 0:
                # everything below contributes to the LATE gradient bucket only (dist.flat_layout): the
                # gradients of layers >= 1 are final once the side streams have been joined here
                self.bwd_split = len(bwd)
            pre = f"spatial_encoder.{i}."
            sa = pre + "self_attn."
            s0 = i * 8
            dx = ptr(w["dx"][i])
            df, du, do_, dqkvc = ptr(w["df"][i]), ptr(w["du"][i]), ptr(w["do"][i]), ptr(w["dqkvc"][i])
            ps_b = self.ffn_pair and self.attn_block_active and self.ffn_presplit
            if ps_b:
                self._call(bwd, "msr3d_ffn_tail_bwd_ps_f32", dy3, M, P(pre + "norm2.weight"), seed, sptr, p, s0 + 3,
                           ptr(w["xhat_2"][i]), ptr(w["rstd_2"][i]), ptr(w["dy2"]), df, G(pre + "norm2.weight"),
                           G(pre + "norm2.bias"), G(pre + "linear2.bias"), ptr(w["df_big"]), ptr(w["df_small"]),
                           tag="ffn_tail_bwd")
            else:
                self._call(bwd, "msr3d_ffn_tail_bwd_f32", dy3, M, P(pre + "norm2.weight"), seed, sptr, p, s0 + 3,
                           ptr(w["xhat_2"][i]), ptr(w["rstd_2"][i]), ptr(w["dy2"]), df, G(pre + "norm2.weight"),
                           G(pre + "norm2.bias"), G(pre + "linear2.bias"), tag="ffn_tail_bwd")
            # dW2 (H x F) += df^T . h                                   [side stream]
            self._gemm(bwd, df, (1, H), ptr(w["h"][i]), (1, F), H, F, M, G(pre + "linear2.weight"), F,
                       accumulate=2, ksplit=2, tag="dW2", wgrad=True, side=True)
            if ps_b:
                self._call(bwd, "msr3d_ffn_pair_bwd_ps_f32", ptr(w["df_big"]), ptr(w["df_small"]),
                           self.PB(pre + "linear1.weight"), self.PS(pre + "linear1.weight"),
                           self.PB(pre + "linear2.weight"), self.PS(pre + "linear2.weight"),
                           ptr(w["gp"][i]), M, F, du, ptr(w["dy2"]), self.ffn_slice, tag="ffn_pair_bwd")
            elif self.ffn_pair:
                # du = (df . W2) * gelu'(u) * mask and dy2 += du . W1 in one kernel
                self._call(bwd, "msr3d_ffn_pair_bwd_f32", df, P(pre + "linear1.weight"), P(pre + "linear2.weight"),
                           ptr(w["gp"][i]), M, F, du, ptr(w["dy2"]), self.ffn_slice, tag="ffn_pair_bwd")
            else:
                # du = (df . W2) * gelu'(u) * mask
                self._gemm(bwd, df, (H, 1), P(pre + "linear2.weight"), (1, F), M, F, H, du, F,
                           epilogue=3, aux=ptr(w["gp"][i]), ldaux=F, tag="du")
            # dW1 (F x H) += du^T . y2 ; db1                            [side stream]
            self._gemm(bwd, du, (1, F), ptr(w["y2"][i]), (1, H), F, H, M, G(pre + "linear1.weight"), H,
                       accumulate=2, ksplit=2, tag="dW1", wgrad=True, side=2)
            self._call(bwd, "msr3d_colsum_add_f32", du, F, M, F, G(pre + "linear1.bias"), None, None, tag="db1",
                       side=2)
            if not self.ffn_pair:
                # dy2 += du . W1
                self._gemm(bwd, du, (F, 1), P(pre + "linear1.weight"), (1, H), M, H, F, ptr(w["dy2"]), H,
                           accumulate=2, ksplit=16, tag="dy2")
            self._call(bwd, "msr3d_attn_tail_bwd_f32", ptr(w["dy2"]), M, P(sa + "layer_norm.weight"),
                       P(pre + "norm1.weight"), seed, sptr, p, s0 + 0, s0 + 1, ptr(w["xhat_a"][i]),
                       ptr(w["rstd_a"][i]), ptr(w["xhat_1"][i]), ptr(w["rstd_1"][i]), dx, do_,
                       G(sa + "layer_norm.weight"), G(sa + "layer_norm.bias"), G(pre + "norm1.weight"),
                       G(pre + "norm1.bias"), G(sa + "fc.bias"), tag="attn_tail_bwd")
            # dWo (H x H) += do^T . ctx [side] ;  dctx = do . Wo
            self._gemm(bwd, do_, (1, H), ptr(w["ctx"][i]), (1, H), H, H, M, G(sa + "fc.weight"), H,
                       accumulate=2, ksplit=4, tag="dWo", wgrad=True, side=True)
            self._gemm(bwd, do_, (H, 1), P(sa + "fc.weight"), (1, H), M, H, H, ptr(w["dctx"][i]), H,
                       accumulate=2, ksplit=ks_h, tag="dctx")
            # attention backward: tensor-core kernel (one CTA per head and scene, csrc/attn_block.cu) when L <= 64
            bwd_tc = self.attn_block_active and _os_env("MSR3D_ATTN_BWD_TC", "1") == "1"
            if cond:
                self._call(bwd, "msr3d_attn_bwd_tc_f32" if bwd_tc else "msr3d_attn_bwd_f32", ptr(w["dctx"][i]), H,
                           ptr(w["qkvc"][i]), QKVC, ptr(w["pw"]), ptr(w["P"][i]), ptr(w["LOC"][i]), B, L, NHEAD, dqkvc,
                           tag="attn_core_bwd")
            else:
                self._call(bwd, "msr3d_attn_fusion_bwd_f32", ptr(w["dctx"][i]), H, ptr(w["qkvc"][i]), QKVC, ptr(w["pw"]),
                           ptr(w["P"][i]), ptr(w["LOC"][i]), B, L, NHEAD, fcode, dqkvc, tag="attn_core_bwd")
                # d(spatial weights) -> pairwise_loc_fc gradients (ctx: plus its share of dq)
                self._call(bwd, "msr3d_spatial_weights_bwd_f32", P(sa + "pairwise_loc_fc.weight"),
                           P(sa + "pairwise_loc_fc.bias"), int(self.fusion == "ctx"), self.multihead, ptr(w["qkvc"][i]),
                           dqkvc, QKVC, M, NHEAD, G(sa + "pairwise_loc_fc.weight"), G(sa + "pairwise_loc_fc.bias"),
                           tag="spatial_w_bwd")
            # dWqkvc (816 x H) += dqkvc^T . x_in ; bias [side] ; dx += dqkvc . Wqkvc     (768 rows without `cond`)
            self._gemm(bwd, dqkvc, (1, QKVC), ptr(w["x_in"][i]), (1, H), NQ, H, M, G(sa + "w_qs.weight"), H,
                       accumulate=2, ksplit=4, tag="dWqkvc", wgrad=True, side=2)
            self._call(bwd, "msr3d_colsum_add_f32", dqkvc, QKVC, M, NQ, G(sa + "w_qs.bias"), None, None,
                       tag="dbqkvc", side=2)
            self._gemm(bwd, dqkvc, (QKVC, 1), P(sa + "w_qs.weight"), (1, H), M, H, NQ, dx, H,
                       accumulate=2, ksplit=max(1, -(-NQ // 64)), tag="dx")
            dy3 = dx
            if cross:
                # backward of situation_condition_layer[i]: dx (grad wrt x_in[i]) -> dxa[i] (grad wrt x_pre[i])
                pc = f"situation_condition_layer.{i}."
                df_c, du_c = ptr(w["df_c"][i]), ptr(w["du_c"][i])
                self._call(bwd, "msr3d_cross_ffn_residual_bwd_f32", dx, M, seed, sptr, p, s0 + 7, df_c, G(pc + "linear2.bias"),
                           tag="cross_out_bwd")
                self._gemm(bwd, df_c, (1, H), ptr(w["h_c"][i]), (1, F), H, F, M, G(pc + "linear2.weight"), F,
                           accumulate=2, ksplit=2, tag="dW2c", wgrad=True, side=True)
                self._call(bwd, "msr3d_ffn_pair_bwd_f32", df_c, P(pc + "linear1.weight"), P(pc + "linear2.weight"),
                           ptr(w["gp_c"][i]), M, F, du_c, ptr(w["dt3"][i]), self.ffn_slice, tag="cross_ffn_pair_bwd")
                self._gemm(bwd, du_c, (1, F), ptr(w["t3"][i]), (1, H), F, H, M, G(pc + "linear1.weight"), H,
                           accumulate=2, ksplit=2, tag="dW1c", wgrad=True, side=2)
                self._call(bwd, "msr3d_colsum_add_f32", du_c, F, M, F, G(pc + "linear1.bias"), None, None, tag="db1c", side=2)
                self._call(bwd, "msr3d_cross_ln_bwd_f32", ptr(w["dt3"][i]), dx, M, P(pc + "norm3.weight"), seed, sptr, p, s0 + 5,
                           ptr(w["xhat3"][i]), ptr(w["rstd3"][i]), ptr(w["dxa"][i]), ptr(w["g_attn"]), G(pc + "norm3.weight"),
                           G(pc + "norm3.bias"), G(pc + "multihead_attn.out_proj.bias"), tag="cross_ln_bwd")
                self._call(bwd, "msr3d_cross_vec_bwd_f32", ptr(w["g_attn"]), ptr(w["scale_c"][i]), B, L, ptr(w["v_c"][i]),
                           ptr(w["s_sit"]), P(pc + "multihead_attn.in_proj_weight") + 4 * 2 * H * H,
                           P(pc + "multihead_attn.out_proj.weight"), G(pc + "multihead_attn.out_proj.weight"),
                           G(pc + "multihead_attn.in_proj_weight") + 4 * 2 * H * H,
                           G(pc + "multihead_attn.in_proj_bias") + 4 * 2 * H, ptr(w["ds_sit"]), tag="cross_vec_bwd")
                dy3 = ptr(w["dxa"][i])
        dx0 = ptr(w["dxa"][0]) if cross else ptr(w["dx"][0])
        dpr = dx0
        if anchor:
            # object rows of dx0 -> compact dproj; agent rows -> anchor_feat, type_embedding[1], orientation_encoder
            dpr = ptr(w["dproj"])
            self._call(bwd, "msr3d_anchor_disassemble_f32", dx0, ptr(w["anchor_ori"]), B, Lo, dpr, G("anchor_feat"),
                       G("object_type_embedding.weight") + 4 * H, G("orientation_encoder.weight"),
                       G("orientation_encoder.bias"), tag="anchor_disassemble")
        # x_in0 = emb Wp^T + bp + orientation_feat + type_emb[0] + qpos     (object rows)
        self._call(bwd, "msr3d_colsum_add_f32", dpr, H, Mo, H, G("obj_linear_projection.bias"),
                   G("object_orientation_feat"), G("object_type_embedding.weight"), tag="dbias_proj", side=True)
        E = self.emb_dim
        self._gemm(bwd, dpr, (1, H), self.emb_ptr, (1, E), H, E, Mo, G("obj_linear_projection.weight"), E,
                   accumulate=2, ksplit=4, tag="dWproj", wgrad=True, side=True)
        dq = [ptr((w["dxa"] if cross else w["dx"])[i]) for i in range(nl)] if self.same_all else [dx0]
        dq += [None] * (3 - len(dq))
        assert len(dq) == 3, "fused qpos backward sums at most three layers"
        if diff_all:
            for i in range(nl):          # every layer's own loc_layers[i] from the gradient of that layer's input
                self._call(bwd, "msr3d_loc_layer_bwd_f32", ptr((w["dxa"] if cross else w["dx"])[i]), None, None,
                           ptr(w["all_loc"] if anchor else w["obj_locs"]), M, P(f"loc_layers.{i}.1.weight"),
                           ptr(w["xhat_q"][i]), ptr(w["rstd_q"][i]), G(f"loc_layers.{i}.1.weight"), G(f"loc_layers.{i}.1.bias"),
                           G(f"loc_layers.{i}.0.weight"), G(f"loc_layers.{i}.0.bias"), tag="qpos_bwd")
            if cross:
                self._call(bwd, "msr3d_situation_embed_bwd_f32", ptr(w["ds_sit"]), None, None, ptr(w["anchor_locs"]),
                           ptr(w["anchor_ori"]), B, 1, P("loc_embedding_encoder.1.weight"), ptr(w["xhat_sit"]),
                           ptr(w["rstd_sit"]), G("loc_embedding_encoder.0.weight"), G("loc_embedding_encoder.0.bias"),
                           G("loc_embedding_encoder.1.weight"), G("loc_embedding_encoder.1.bias"),
                           G("orientation_encoder.weight"), G("orientation_encoder.bias"), tag="situation_vec_bwd")
        elif as_object or cross:
            self._call(bwd, "msr3d_loc_layer_bwd_f32", dq[0], dq[1], dq[2], ptr(w["all_loc"] if as_object else w["obj_locs"]),
                       M, P("loc_layers.0.1.weight"), ptr(w["xhat_l"]), ptr(w["rstd_l"]), G("loc_layers.0.1.weight"),
                       G("loc_layers.0.1.bias"), G("loc_layers.0.0.weight"), G("loc_layers.0.0.bias"), tag="qpos_bwd")
            if cross:       # d(situation vector), summed over the layers -> loc_embedding_encoder / orientation_encoder
                self._call(bwd, "msr3d_situation_embed_bwd_f32", ptr(w["ds_sit"]), None, None, ptr(w["anchor_locs"]),
                           ptr(w["anchor_ori"]), B, 1, P("loc_embedding_encoder.1.weight"), ptr(w["xhat_sit"]),
                           ptr(w["rstd_sit"]), G("loc_embedding_encoder.0.weight"), G("loc_embedding_encoder.0.bias"),
                           G("loc_embedding_encoder.1.weight"), G("loc_embedding_encoder.1.bias"),
                           G("orientation_encoder.weight"), G("orientation_encoder.bias"), tag="situation_vec_bwd")
        else:
            self._call(bwd, "msr3d_qpos_bwd_f32", dq[0], dq[1], dq[2], ptr(w["sz"]), M, P("loc_embedding_encoder.1.weight"),
                       P("size_embedding_encoder.1.weight"), ptr(w["xhat_l"]), ptr(w["rstd_l"]), ptr(w["xhat_s"]),
                       ptr(w["rstd_s"]), ptr(w["dh1"]), G("loc_embedding_encoder.1.weight"), G("loc_embedding_encoder.1.bias"),
                       G("size_embedding_encoder.1.weight"), G("size_embedding_encoder.1.bias"),
                       G("size_embedding_encoder.0.weight"), G("size_embedding_encoder.0.bias"),
                       G("loc_embedding_encoder.0.bias"), tag="qpos_bwd")
            self._gemm(bwd, ptr(w["dh1"]), (1, H), ptr(w["ff"]), (1, 64), H, 63, M, G("loc_embedding_encoder.0.weight"), 63,
                       accumulate=2, ksplit=4, tag="dWloc", wgrad=True)
            if self.sit_embed:
                self._call(bwd, "msr3d_situation_embed_bwd_f32", dq[0], dq[1], dq[2], ptr(w["anchor_locs"]),
                           ptr(w["anchor_ori"]), B, L, P("loc_embedding_encoder.1.weight"), ptr(w["xhat_sit"]),
                           ptr(w["rstd_sit"]), G("loc_embedding_encoder.0.weight"), G("loc_embedding_encoder.0.bias"),
                           G("loc_embedding_encoder.1.weight"), G("loc_embedding_encoder.1.bias"),
                           G("orientation_encoder.weight"), G("orientation_encoder.bias"), tag="situation_embed_bwd")
        bwd.append(("join", None, (), None, False))
        # prefetch schedule: the geometry was computed beside the encoder of the previous step
        # and the gradient buffer is cleared beside the input encoders (first written by the backward pass)
        # ... and the weight images were written by the optimiser of the previous step (refresh_split() after any other
        # change of the parameters)
        clear = ("memset", None, (self.flat_g,), None, 2)
        if self.accum_steps > 1:
            clear = ("msr3d_zero_gated_f32", _C.lib().msr3d_zero_gated_f32,
                     (self.flat_g.data_ptr(), (self.n_opt + 3) // 4 * 4, self.accum_state.data_ptr() + 8), "grad_clear", 2)
        self.pre_prog_nogeo = [e for e in pre_prog if e[3] not in ("geometry", "split_w")] + [clear]
        return pre_prog, fwd, bwd

    def _run(self, prog):
        """Issue a launch list.  Entries flagged `side` (1 or 2) go to that side stream after it has
        caught up with the main stream (fork); a "join" entry (and the end of the list) makes the main
        stream wait for the side streams.  Under CUDA-graph capture this becomes parallel graph branches."""
        main = torch.cuda.current_stream()
        sides = (self.side_stream, self.side_stream2)
        timer = ops.TIMER
        for name, fn, args, tag, on_side in prog:
            if name == "join":
                for k, side in enumerate(sides):
                    if self._side_pending[k]:
                        main.wait_stream(side)
                        self._side_pending[k] = False
                continue
            if fn is None:                  # memset of a workspace tensor
                if on_side and self.use_side:
                    k = int(on_side) - 1
                    sides[k].wait_stream(main)
                    self._side_pending[k] = True
                    with torch.cuda.stream(sides[k]):
                        args[0].zero_()
                else:
                    args[0].zero_()
                continue
            if on_side and self.use_side:
                k = int(on_side) - 1
                side = sides[k]
                side.wait_stream(main)
                self._side_pending[k] = True
                if timer is not None and tag:
                    with torch.cuda.stream(side), timer.range(tag):
                        rc = fn(*args, side.cuda_stream)
                else:
                    rc = fn(*args, side.cuda_stream)
            elif timer is not None and tag:
                with timer.range(tag):
                    rc = fn(*args, main.cuda_stream)
            else:
                rc = fn(*args, main.cuda_stream)
            if rc != 0:
                raise _C.Msr3dError(f"{name} [{tag}] failed ({rc}): {_C.lib().msr3d_last_error().decode()}")

    # ------------------------------------------------------------------ public API
    def prepare(self, B, L, n_points, training=True):
        """(Re)build buffers and launch programs for a batch geometry / train-eval mode."""
        key = (B, L, n_points, bool(training), self.split)
        if getattr(self, "_key", None) == key:
            return
        enc = self.model.obj_encoder
        self.emb_dim = enc.pcd_net.fc.out_features
        if self._shape != (B, L, n_points):
            self._build(B, L, n_points)
            self.ws["emb"] = self._buf(B * L, self.emb_dim)
            self.ws["emb_next"] = self._buf(B * L, self.emb_dim)
        self.emb_ptr = self.ws["emb"].data_ptr()
        self._keep = []
        self.pre_prog, self.fwd_prog, self.bwd_prog = self._programs(training)
        self._key = key
        self.graph = None

    def load_inputs(self, batch, upstream: Optional[torch.Tensor] = None, non_blocking=True):
        """Copy a SceneBatch (host pinned or device) into the static input buffers."""
        w = self.ws
        w["obj_fts"].copy_(batch.obj_fts, non_blocking=non_blocking)
        w["obj_locs"].copy_(batch.obj_locs, non_blocking=non_blocking)
        w["obj_masks"].copy_(batch.obj_masks, non_blocking=non_blocking)
        w["anchor_locs"].copy_(batch.anchor_locs, non_blocking=non_blocking)
        w["anchor_ori"].copy_(batch.anchor_orientation, non_blocking=non_blocking)
        if upstream is not None:
            w["upstream"].copy_(upstream.reshape(w["upstream"].shape), non_blocking=non_blocking)

    def _encode(self, out_key="emb", src_key="obj_fts"):
        w = self.ws
        if getattr(self, "train_encoder", False):
            # unfrozen encoder, engine API: explicit forward of the training path (encoder_train.py), state kept for backward()
            assert out_key == "emb", "the unfrozen encoder runs the sequential schedule (train_step_eager)"
            from . import encoder_train
            B, L, Pn, _ = w[src_key].shape
            out, self._enc_state = encoder_train.forward_explicit(self.model.obj_encoder.pcd_net, w[src_key].view(B * L, Pn, 6))
            w[out_key].copy_(out)
            return
        if not encoder_frozen(self.model):
            raise NotImplementedError("fused prompter: this engine was built for the module call (alias_grads=False); with "
                                      "vision.args.freeze False either call the module or build the engine with alias_grads=True")
        enc = self.model.obj_encoder
        B, L, Pn, _ = w[src_key].shape
        enc.freeze_bn(enc.pcd_net)
        enc.pcd_net.forward_fused(w[src_key].view(B * L, Pn, 6), self.enc_split, out=w[out_key], aux_stream=self._enc_aux())

    def _enc_aux(self):
        """Second stream of the encoder (same SM partition when there is one): level 2's grouping runs beside SA1's MLP.
        A/B: MSR3D_ENC_FORK=0/1."""
        if "_enc_aux_stream" not in self.__dict__:
            import os
            if os.environ.get("MSR3D_ENC_FORK", "0") != "1":
                self._enc_aux_stream = None
            elif self.partition is not None:
                self._enc_aux_stream = self.partition.stream(1)
            else:
                self._enc_aux_stream = torch.cuda.Stream(device=self.dev)
        return self._enc_aux_stream

    # ------------------------------------------------------------------ pipelined mode
    # The encoder is frozen, so the embedding of batch i+1 does not depend on the update made with
    # batch i: it is computed on its own stream while the transformer forward/backward/optimiser of
    # batch i runs (what a prefetching data loader does for the host side).  Call order per step:
    #   stage_points(batch_{i+1}) -> pipelined step (consumes batch_i) -> stage_scene(batch_{i+1})
    # Results are identical to the sequential order; only the wall-clock placement changes.
    def stage_points(self, batch, non_blocking=True):
        self.ws["obj_fts"].copy_(batch.obj_fts, non_blocking=non_blocking)

    def stage_scene(self, batch, upstream, non_blocking=True, level="_n"):
        w = self.ws
        w["obj_locs" + level].copy_(batch.obj_locs, non_blocking=non_blocking)
        w["obj_masks" + level].copy_(batch.obj_masks, non_blocking=non_blocking)
        w["anchor_locs" + level].copy_(batch.anchor_locs, non_blocking=non_blocking)
        w["anchor_ori" + level].copy_(batch.anchor_orientation, non_blocking=non_blocking)
        w["upstream" + level].copy_(upstream.reshape(w["upstream_n"].shape), non_blocking=non_blocking)

    def prime(self, batch, upstream):
        """Fill the pipeline: encode `batch` now; its transformer step happens in the next call."""
        assert not (self.n_anchor or self.sit_embed or self.cross or self.diff_all), \
            "the non-shipped situation modes run the sequential schedule (forward / backward / train_step)"
        self.stage_points(batch)
        self._encode("emb_next")
        self.stage_scene(batch, upstream)
        self._primed = True

    def _pipelined_body(self, with_encoder=True, slot=0, depth2=False, part=0):
        """part 0: the whole step.  Parts 1 / 2 (data-parallel runs): everything up to the point where the
        early gradient bucket is final / the rest of the backward — the all-reduce of the early bucket is
        issued between them and overlaps part 2."""
        w = self.ws
        main = torch.cuda.current_stream()
        if part == 2:
            self._run(self.bwd_prog[self.bwd_split:])
            return
        if depth2:
            self._accum_tick(0)
            # ONE launch moves every staged input one generation on (csrc/stage.cu)
            rc = _C.lib().msr3d_stage_step(ctypes.byref(self._stage_desc(slot)), main.cuda_stream)
            if rc != 0:
                raise _C.Msr3dError(f"stage_step failed ({rc}): {_C.lib().msr3d_last_error().decode()}")
            if with_encoder:                   # fork: geometry + encoder of the incoming batch
                self.enc_stream.wait_stream(main)
                with torch.cuda.stream(self.enc_stream):
                    self._run(self.geo_prog[slot])
                    self._encode("emb_next", "obj_fts_b" if slot else "obj_fts")
            self._run(self.pre_prog_nogeo)
        else:
            for k in ("obj_locs", "obj_masks", "anchor_locs", "anchor_ori", "upstream"):
                w[k].copy_(w[k + "_n"])
            w["emb"].copy_(w["emb_next"])
            if with_encoder:                       # fork: encoder of the incoming batch
                self.enc_stream.wait_stream(main)
                with torch.cuda.stream(self.enc_stream):
                    self._encode("emb_next", "obj_fts_b" if slot else "obj_fts")
            self.flat_g.zero_()
            torch.logical_not(w["obj_masks"], out=w["key_pad"].view(torch.bool))
            self._run(self.pre_prog)
        self._run(self.fwd_prog)
        if part == 1:
            self._run(self.bwd_prog[:self.bwd_split] + [("join", None, (), None, False)])
        elif part == 3:
            # whole data-parallel step in one stream-ordered program (capturable): gradients of layers >= 1 are
            # all-reduced on the communication stream under the backward of layer 0 and the input encoders, the rest
            # right after it; clip + AdamW follow on the device
            import torch.distributed as dist
            (e0, e1), (l0, l1) = self.bucket_early, self.bucket_late
            split_w = depth2 and self.attn_block_active
            if self.p2p is not None:
                # this library's own exchange over NVLink peer memory (csrc/allreduce.cu), norm partials fused in
                nb0, nb1 = 32, 48
                if self.ddp_buckets and e1 > e0:
                    self._run(self.bwd_prog[:self.bwd_split])
                    self._comm_waits_for_gradients(main)
                    self._allreduce_p2p(0, e0, e1, nb0, self.comm_stream)
                    self._run(self.bwd_prog[self.bwd_split:])
                    self._allreduce_p2p(1, l0, l1, nb1, main)
                    main.wait_stream(self.comm_stream)
                    self._optimizer_after_p2p(nb0, nb1, split_w)
                else:
                    self.backward()
                    self._allreduce_p2p(0, 0, self.n_opt, AR_MAX_BLOCKS, main)
                    self._optimizer_after_p2p(AR_MAX_BLOCKS, 0, split_w)
                self._accum_tick(1)
            else:
                if self.ddp_buckets and e1 > e0:
                    self._run(self.bwd_prog[:self.bwd_split])
                    self._comm_waits_for_gradients(main)
                    with torch.cuda.stream(self.comm_stream):
                        dist.all_reduce(self.flat_g[e0:e1])
                    self._run(self.bwd_prog[self.bwd_split:])
                    dist.all_reduce(self.flat_g[l0:l1])
                    main.wait_stream(self.comm_stream)
                else:
                    self.backward()
                    dist.all_reduce(self.flat_g[:self.n_opt])
                self.optimizer_step(split=split_w)
        elif self.world_size == 1 and depth2 and self.early_norm and self.bucket_early[1] > self.bucket_early[0]:
            self._backward_optimizer_early_norm(main, split=self.attn_block_active)
        else:
            self.backward()
            if self.world_size == 1:
                if depth2 and self.accum_steps > 1:
                    self._optimizer_accum(self.attn_block_active)
                    self._accum_tick(1)
                else:
                    self.optimizer_step(split=depth2 and self.attn_block_active)
        if with_encoder:
            main.wait_stream(self.enc_stream)  # join

    def _stage_desc(self, slot):
        """Descriptor of the staging launch of a prefetch-schedule step that encodes the batch in `slot` (kept alive
        in self._stage_keep: a captured graph holds no reference to host memory, but eager launches read it)."""
        cache = self.__dict__.setdefault("_stage_keep", {})
        if slot in cache and cache[slot][0] is self.ws:
            return cache[slot][1]
        w, q = self.ws, "_q1" if slot else "_q0"
        d = StageDesc()
        rot = [k for k in ("obj_locs", "anchor_locs", "anchor_ori", "upstream")]
        for i, k in enumerate(rot):
            d.rot_a[i], d.rot_b[i], d.rot_c[i], d.rot_n[i] = w[k].data_ptr(), w[k + "_n"].data_ptr(), w[k + q].data_ptr(), w[k].numel()
        d.n_rotate = len(rot)
        cps = [("emb", "emb_next"), ("pw", "pw_n"), ("ff", "ff_n"), ("sz", "sz_n")]
        for i, (a, b) in enumerate(cps):
            d.cp_dst[i], d.cp_src[i], d.cp_n[i] = w[a].data_ptr(), w[b].data_ptr(), w[a].numel()
        d.n_copy = len(cps)
        d.mask_a, d.mask_b, d.mask_c = w["obj_masks"].data_ptr(), w["obj_masks_n"].data_ptr(), w["obj_masks" + q].data_ptr()
        d.key_pad, d.n_mask = w["key_pad"].data_ptr(), w["obj_masks"].numel()
        cache[slot] = (self.ws, d)
        return d

    def train_step_pipelined(self, flush=False):
        """Transformer step on the staged batch (+ encoder of the batch in `obj_fts` unless flush)."""
        assert self._primed, "call prime() first"
        if self.graph_pipelined is not None and not flush:
            self.graph_pipelined.replay()
        else:
            self._pipelined_body(with_encoder=not flush)
        if self.world_size > 1:
            from . import dist as _dist
            _dist.allreduce_sum_(self.flat_g[:self.n_opt])
            self.optimizer_step()
        return self._tokens()

    def capture_pipelined(self):
        """CUDA graph of one pipelined step.  Warm-up replays leave parameters untouched."""
        assert self._primed
        saved = (self.flat_p.clone(), self.m.clone(), self.v.clone(), self.opt_state.clone(), self.seed_counter.clone())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._pipelined_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # captured on a HIGH-priority stream: the transformer chain is latency bound (small grids, one
        # after the other) while the overlapped encoder is throughput work on the default-priority
        # stream — the block scheduler should hand freed SMs to the chain first
        with torch.cuda.graph(g, stream=self.hi_stream, capture_error_mode="thread_local"):
            self._pipelined_body()
        self.graph_pipelined = g
        for dst, src in zip((self.flat_p, self.m, self.v, self.opt_state, self.seed_counter), saved):
            dst.copy_(src)
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ prefetch mode
    # One more stage in front of the pipeline above: the point clouds of batch i are copied host ->
    # device on a copy stream while the step that encodes batch i-1 and updates with batch i-2 runs,
    # i.e. what a prefetching data loader with pinned memory does; the 5.9 MB H2D copy (~0.12 ms at
    # config 2) leaves the critical path.  Call order per step i:
    #   prefetch(batch_i, upstream_i, i % 2) -> train_step_prefetch((i - 1) % 2)
    # Everything of batch i (points and the small per-scene inputs) travels on the copy stream into the
    # buffers of slot i % 2; nothing is copied on the compute stream (five small H2D copies in stream
    # order cost ~10 us of DMA latency each).
    def prefetch(self, batch, upstream, slot, non_blocking=True):
        cs = self.copy_stream
        cs.wait_event(self._slot_free[slot])       # the step that last read this slot has finished
        with torch.cuda.stream(cs):
            self.ws["obj_fts_b" if slot else "obj_fts"].copy_(batch.obj_fts, non_blocking=non_blocking)
            self.stage_scene(batch, upstream, non_blocking=non_blocking, level="_q1" if slot else "_q0")
            self._copied[slot].record(cs)

    def prime2(self, batch0, up0, batch1, up1):
        """Fill both stages: encode batch0 now, park batch1's points in slot 1."""
        assert not (self.n_anchor or self.sit_embed or self.cross or self.diff_all), \
            "the non-shipped situation modes run the sequential schedule (forward / backward / train_step)"
        main = torch.cuda.current_stream()
        self._slot_free[0].record()
        self._slot_free[1].record()
        self.prefetch(batch0, up0, 0)
        main.wait_event(self._copied[0])
        self._encode("emb_next", "obj_fts")
        self._run(self.geo_prog[0])
        if self.attn_block_active:
            self.refresh_split()
        for k in ("obj_locs", "obj_masks", "anchor_locs", "anchor_ori", "upstream"):
            self.ws[k + "_n"].copy_(self.ws[k + "_q0"])
        self._slot_free[0].record()
        self.prefetch(batch1, up1, 1)
        self._primed2 = True

    def train_step_prefetch(self, slot, flush=0):
        """Transformer step on the oldest staged batch + encoder of the batch in `slot`.
        flush=1: same, eagerly (drains the pipeline); flush=2: last step, nothing left to encode."""
        assert self._primed2, "call prime2() first"
        main = torch.cuda.current_stream()
        if flush != 2:
            main.wait_event(self._copied[slot])
        graphed = self.graph_prefetch[slot] is not None and not flush
        if self.world_size > 1 and self.ddp_in_graph:
            if graphed:
                self.graph_prefetch[slot].replay()
            else:
                self._pipelined_body(with_encoder=flush != 2, slot=slot, depth2=True, part=3)
            self._slot_free[slot].record(main)
        elif self.world_size == 1 or not self.ddp_buckets:
            if graphed:
                self.graph_prefetch[slot].replay()
            else:
                self._pipelined_body(with_encoder=flush != 2, slot=slot, depth2=True)
            self._slot_free[slot].record(main)
            if self.world_size > 1:
                from . import dist as _dist
                _dist.allreduce_sum_(self.flat_g[:self.n_opt])
                self.optimizer_step(split=self.attn_block_active)
        else:
            # two gradient buckets: layers >= 1 are reduced while layer 0 (and the input encoders) are
            # still in their backward pass; the optimiser waits for both collectives on the device
            import torch.distributed as dist
            (e0, e1), (l0, l1) = self.bucket_early, self.bucket_late
            if graphed:
                self.graph_prefetch[slot].replay()
            else:
                self._pipelined_body(with_encoder=flush != 2, slot=slot, depth2=True, part=1)
            self._slot_free[slot].record(main)
            w_early = dist.all_reduce(self.flat_g[e0:e1], async_op=True) if e1 > e0 else None
            if graphed:
                self.graph_prefetch_tail.replay()
            else:
                self._pipelined_body(part=2)
            w_late = dist.all_reduce(self.flat_g[l0:l1], async_op=True)
            if w_early is not None:
                w_early.wait()
            w_late.wait()
            self.optimizer_step(split=self.attn_block_active)
        return self._tokens()

    def capture_prefetch(self):
        """Two CUDA graphs (one per point-cloud slot).  Warm-up runs leave parameters and staged inputs untouched."""
        assert self._primed2
        w = self.ws
        keys = [k + lv for k in ("obj_locs", "obj_masks", "anchor_locs", "anchor_ori", "upstream")
                for lv in ("", "_n", "_q0", "_q1")]
        keys += ["pw", "ff", "sz", "pw_n", "ff_n", "sz_n"]
        saved = [t.clone() for t in (self.flat_p, self.m, self.v, self.opt_state, self.seed_counter, w["emb"], w["emb_next"],
                                     self.accum_state)]
        saved_in = {k: w[k].clone() for k in keys}
        torch.cuda.synchronize()
        for slot in (0, 1):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._pipelined_body(slot=slot, depth2=True, part=3 if (self.world_size > 1 and self.ddp_in_graph) else 0)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            in_graph = self.world_size > 1 and self.ddp_in_graph
            with torch.cuda.graph(g, stream=self.hi_stream, capture_error_mode="thread_local"):
                self._pipelined_body(slot=slot, depth2=True,
                                     part=3 if in_graph else (1 if (self.world_size > 1 and self.ddp_buckets) else 0))
            self.graph_prefetch[slot] = g
        if self.world_size > 1 and self.ddp_buckets and not self.ddp_in_graph:      # the rest of the backward (same for both slots)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.hi_stream, capture_error_mode="thread_local"):
                self._pipelined_body(part=2)
            self.graph_prefetch_tail = g
        for dst, src in zip((self.flat_p, self.m, self.v, self.opt_state, self.seed_counter, w["emb"], w["emb_next"],
                             self.accum_state), saved):
            dst.copy_(src)
        for k in keys:
            w[k].copy_(saved_in[k])
        if self.attn_block_active:
            self.refresh_split()
        torch.cuda.synchronize()

    def forward(self, emb: Optional[torch.Tensor] = None):
        """Encoder + transformer forward on the static buffers; returns tokens (B, L, 256) view."""
        w = self.ws
        if not self.n_anchor:         # agent-token modes: msr3d_anchor_inputs writes key_pad (first launch of pre_prog)
            torch.logical_not(w["obj_masks"], out=w["key_pad"].view(torch.bool))
        self._run(self.pre_prog)      # geometry + query positions on the side stream ...
        if emb is None:
            self._encode()            # ... while the encoder runs on the main stream
        else:                         # embeddings computed by the caller (unfrozen encoder, autograd.py)
            w["emb"].copy_(emb.reshape(w["emb"].shape))
        self._run(self.fwd_prog)      # starts with a join
        return self._tokens()

    def _tokens(self):
        """(B, tokens per scene, 256) view of the output buffer (tokens per scene = objects + 1 in the agent-token modes)."""
        B, L = self.ws["key_pad"].shape
        return self.ws["tokens"].view(B, L, H)

    def token_masks(self) -> torch.Tensor:
        """obj_masks as OSE3DSituation returns them (ose3d_situation.py:452): True = valid token, agent token included."""
        return self.ws["key_pad"] == 0

    # ------------------------------------------------------------------ drop-in module call as CUDA graphs
    # OSE3DSituation.forward / loss.backward() (autograd.py) launch ~40 / ~45 kernels through ctypes: at config 2 the host
    # needs longer to issue them than the GPU to run them.  The two programs are therefore captured once per
    # (shape, mode) and replayed; inputs, parameters, saved activations and the upstream gradient all live in static
    # buffers already.  The packed weights of the frozen encoder are part of the capture: a change of the encoder's
    # parameters (checkpoint load) invalidates the graphs.
    def module_graph(self, which: str, external_emb: bool):
        import os
        if os.environ.get("MSR3D_MODULE_GRAPH", "1") != "1":
            return None
        cache = self.__dict__.setdefault("_module_graphs", {})
        with_enc = which == "fwd" and not external_emb      # only the forward program contains the encoder
        enc_key = self.model.obj_encoder.pcd_net._packed_weights_key() if with_enc else None
        key = (which, self._key, external_emb)
        hit = cache.get(key)
        if hit is not None and hit[0] == enc_key and hit[2] is self.ws:
            return hit[1]
        if torch.cuda.is_current_stream_capturing():
            return None

        def body():
            if which == "fwd":
                w = self.ws
                if not self.n_anchor:
                    torch.logical_not(w["obj_masks"], out=w["key_pad"].view(torch.bool))
                self._run(self.pre_prog)
                if not external_emb:
                    self._encode()
                self._run(self.fwd_prog)
            else:
                self.flat_g.zero_()
                self._run(self.bwd_prog)
        saved_g = self.flat_g.clone()
        s_ = torch.cuda.Stream(device=self.dev)
        s_.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s_):
            body()                          # warm-up: lazy allocations (encoder workspaces, packed weights) happen here
        torch.cuda.current_stream().wait_stream(s_)
        torch.cuda.synchronize()
        if with_enc:
            enc_key = self.model.obj_encoder.pcd_net._packed_weights_key()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                body()
        except Exception as e:  # noqa: BLE001 — e.g. called while another capture is under way: stay on eager launches
            import warnings
            warnings.warn(f"msr3d_b200: module-call graph capture failed ({type(e).__name__}: {e}); launching eagerly")
            for k in range(2):
                self._side_pending[k] = False
            g = None
        if which == "bwd":
            self.flat_g.copy_(saved_g)
        cache[key] = (enc_key, g, self.ws)
        return g

    def backward(self, need_d_emb: bool = False):
        """Backward from ws['upstream']; accumulates into the flat gradient buffer.  need_d_emb: also returns the
        gradient wrt the object embeddings (B * objects, E) = d(projection output) . W_proj — what an unfrozen encoder
        back-propagates (tensor-core GEMM, msr3d_gemm_tc_f32)."""
        self._run(self.bwd_prog)
        if getattr(self, "train_encoder", False):
            # d(embeddings) -> the encoder's hand-written backward, gradients accumulated straight into the flat buffer
            from . import encoder_train
            net = self.model.obj_encoder.pcd_net
            sinks = {id(p): p.grad for p in net.parameters() if p.grad is not None}
            encoder_train.backward_explicit(net, self._enc_state, self.d_emb(), sinks)
            self._enc_state = None
        return self.d_emb() if need_d_emb else None

    def d_emb(self):
        """Gradient wrt the object embeddings from the finished backward (see backward)."""
        w = self.ws
        dpr = w["dproj"] if self.n_anchor else w["dx"][0]
        Wp = self.by_name["obj_linear_projection.weight"].detach()
        with torch.cuda.device(self.dev):
            return gemm_tc(dpr, Wp.t(), split=6)

    def refresh_split(self):
        """tf32 big / small images of the parameters (after anything but optimizer_step(split=True) changed them)."""
        _C.call("msr3d_split_tf32_f32", self.flat_p.data_ptr(), self.flat_big.data_ptr(), self.flat_small.data_ptr(),
                self.n_split, torch.cuda.current_stream().cuda_stream)

    def optimizer_step(self, zero_grads: bool = False, split: bool = False):
        """Global-norm clip + AdamW over the flat buffer.  `self.extra_sumsq` (a 1-element CUDA tensor, optional)
        is added to the squared gradient norm: set it to the squared norm of the trainable parameters that live
        outside this engine (llm_proj, LoRA, ...) to reproduce the reference's clip over ALL model parameters
        (trainer/leo_trainer.py:192-193); the caller then scales its own gradients by `clip_coef()`."""
        n, (d0, d1) = self.n_opt, self.decay_range
        ex = self.extra_sumsq.data_ptr() if self.extra_sumsq is not None else None
        if split:
            _C.call("msr3d_adamw_clip_step_split_f32", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(),
                    self.v.data_ptr(), n, d0, d1, 1.0 / self.world_size, self.opt_state.data_ptr(),
                    self.ticket.data_ptr(), self.seed_counter.data_ptr(), ex, self.flat_big.data_ptr(),
                    self.flat_small.data_ptr(), torch.cuda.current_stream().cuda_stream)
            return
        _C.call("msr3d_adamw_clip_step_zg_f32", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(),
                self.v.data_ptr(), n, d0, d1, 1.0 / self.world_size, self.opt_state.data_ptr(),
                self.ticket.data_ptr(), self.seed_counter.data_ptr(), ex, 1 if zero_grads else 0,
                torch.cuda.current_stream().cuda_stream)

    def clip_coef(self) -> torch.Tensor:
        """min(1, max_norm / (||g|| + 1e-6)) of the last optimizer_step (device scalar), for parameters the
        caller updates itself with the same global norm."""
        norm = self.opt_state[0].sqrt()
        mx = self.opt_state[7]
        return torch.where(mx > 0, torch.clamp(mx / (norm + 1e-6), max=1.0), torch.ones_like(norm))

    def set_lr(self, lr: float):
        self.opt_state[2] = lr

    @property
    def launches_per_step(self) -> int:
        """Kernel launches of this library in one training step (counted on the first eager step)."""
        return getattr(self, "_launches", 0)

    def train_step_eager(self):
        """zero grads -> fwd -> bwd -> (all-reduce) -> clip + AdamW, on whatever is in the input buffers."""
        n0 = _C.launch_count()
        self._train_step_eager()
        self._launches = _C.launch_count() - n0
        return self._tokens()

    def _train_step_eager(self):
        self.flat_g.zero_()
        tokens = self.forward()
        self.backward()
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat_g[:self.n_opt])
        self.optimizer_step()
        return tokens

    def capture(self):
        """Capture forward+backward (and, single-GPU, the optimiser) into one CUDA graph."""
        assert not getattr(self, "train_encoder", False), "the unfrozen encoder runs eager steps (train_step_eager)"
        for _ in range(2):       # warm-up on a side stream: lazy attribute setup must not happen in capture
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.flat_g.zero_()
                self.forward()
                self.backward()
            torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread polls CUDA events concurrently, which a "global"
        # capture would treat as an illegal call during capture
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self.flat_g.zero_()
            self.forward()
            self.backward()
            if self.world_size == 1:
                self.optimizer_step()
        self.graph = g

    def train_step(self):
        if self.graph is None:
            return self.train_step_eager()
        self.graph.replay()
        if self.world_size > 1:
            from . import dist as _dist
            _dist.allreduce_sum_(self.flat_g[:self.n_opt])
            self.optimizer_step()
        return self._tokens()

    def loss(self):
        """Surrogate scalar <tokens, upstream> (what bench.py reads back): one deterministic single-CTA
        kernel instead of torch's mul + two-pass reduction (20 us cold)."""
        t = self.ws["tokens"]
        self._loss_i ^= 1
        out = self.loss_buf[self._loss_i:self._loss_i + 1]
        _C.call("msr3d_dot_f32", t.data_ptr(), self.ws["upstream"].data_ptr(), t.numel(), out.data_ptr(),
                torch.cuda.current_stream().cuda_stream)
        return out[0]


def gemm_tc(A: torch.Tensor, Bm: torch.Tensor, bias: Optional[torch.Tensor] = None, split: int = 3,
            out: Optional[torch.Tensor] = None, accumulate: int = 0, ksplit: int = 1,
            residual: Optional[torch.Tensor] = None, impl: str = "bf16", tile_n: int = 0,
            aux: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = A[M,K] @ Bm[N,K]^T (+ bias + residual) on the tensor cores; A and Bm are any 2-D
    fp32 CUDA views (strides are passed through, nothing is packed or copied).
    impl "bf16": register-path split-bf16 kernel (linear_tc.cu); "tf32x3": TMA-fed 3xTF32 kernel (gemm3x.cu)."""
    M, K = A.shape
    N = Bm.shape[0]
    assert Bm.shape[1] == K
    if out is None:
        out = torch.zeros(M, N, device=A.device) if accumulate else torch.empty(M, N, device=A.device)
    d = GemmDesc()
    d.A, d.stride_am, d.stride_ak = A.data_ptr(), A.stride(0), A.stride(1)
    d.B, d.stride_bn, d.stride_bk = Bm.data_ptr(), Bm.stride(0), Bm.stride(1)
    d.bias = bias.data_ptr() if bias is not None else None
    if residual is not None:
        d.residual, d.ld_residual = residual.data_ptr(), residual.stride(0)
    d.out, d.ld_out, d.M, d.N, d.K = out.data_ptr(), out.stride(0), M, N, K
    d.split, d.accumulate, d.ksplit, d.tile_n = split, accumulate, ksplit, tile_n
    if aux is not None:
        d.aux, d.ld_aux, d.epilogue = aux.data_ptr(), aux.stride(0), 3
    with torch.cuda.device(A.device):
        if impl == "tf32x3":
            _C.call("msr3d_gemm_tf32x3_f32", ctypes.byref(d), 0, torch.cuda.current_stream().cuda_stream)
        else:
            _C.call("msr3d_gemm_tc_f32", ctypes.byref(d), torch.cuda.current_stream().cuda_stream)
    return out

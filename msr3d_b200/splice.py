"""Scene-token splice into the LLM input — the step right after the prompter.

Replaces the scene branch of `MSR3D.build_embeds` (model/msr3d/msr3d.py:277-287):

    scene_embeds = self.llm_proj(scene_dict['obj_tokens'])                 # Linear 256 -> D
    scene_embeds = scene_embeds.to(dtype=inputs_embeds.dtype)              # fp16 with Vicuna
    scene_embeds_index = torch.where(input_ids == scene_sp_token)          # host sync in the reference
    inputs_embeds[scene_embeds_index] = scene_embeds.reshape(-1, D)
    attention_mask (-> bool)[scene_embeds_index] = scene_dict['obj_masks'].reshape(-1)

Here it is three launches of this library: one CTA computes the placeholder positions (row-major,
the order of `torch.where`) and the spliced bool attention mask; ONE tensor-core GEMM projects the
tokens and its epilogue converts to the embedding dtype and stores every row directly at its
position in `inputs_embeds` (no `scene_embeds` tensor, no index tensors on the host); the backward
gathers the matching rows of dL/d(inputs_embeds) and runs the three gradient GEMMs / column sums.

`SceneSplice` carries `llm_proj` under the reference's parameter names (`llm_proj.weight`,
`llm_proj.bias`, model/msr3d/msr3d.py:84-87) so MSR3D checkpoints load into it.
No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _C
from .fused import GemmDesc

c_int, c_ll, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
_C.register_protos({
    "msr3d_scene_rows_i64": [c_vp, c_int, c_int, c_ll, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp],
    "msr3d_splice_gather_grad": [c_vp, c_int, c_ll, c_vp, c_int, c_int, c_vp, c_vp],
})

_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
SCENE_SP_TOKEN = 31495      # '景' in the Vicuna vocabulary (data/data_utils.py:21-22, msr3d.py:214)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _gemm(A, sa, Bm, sb, M, N, K, out, ldo, bias=None, split=6, accumulate=0, ksplit=1, row_map=None,
          out_dtype=0):
    d = GemmDesc()
    d.A, d.stride_am, d.stride_ak = A, sa[0], sa[1]
    d.B, d.stride_bn, d.stride_bk = Bm, sb[0], sb[1]
    d.bias = bias
    d.out, d.ld_out, d.M, d.N, d.K = out, ldo, M, N, K
    d.split, d.accumulate, d.ksplit = split, accumulate, ksplit
    d.row_map, d.out_dtype = row_map, out_dtype
    _C.call("msr3d_gemm_tc_f32", ctypes.byref(d), _stream())


def scene_rows(input_ids: torch.Tensor, n_rows: int, attention_mask: Optional[torch.Tensor] = None,
               obj_masks: Optional[torch.Tensor] = None, scene_sp_token: int = SCENE_SP_TOKEN):
    """Positions of the scene placeholders.  Returns (rows int32 [n_rows] flat b*T+t in row-major
    order, -1 padded; count int32 [1] on the device; spliced bool attention mask or None)."""
    assert input_ids.is_cuda and input_ids.dtype == torch.int64 and input_ids.dim() == 2, \
        "input_ids must be a 2-D int64 CUDA tensor (msr3d_b200 has no CPU path)"
    ids = input_ids.contiguous()
    B, T = ids.shape
    dev = ids.device
    rows = torch.empty(n_rows, dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    mask_out = am = om = None
    if attention_mask is not None:
        am = attention_mask.to(device=dev, dtype=torch.int64).contiguous()
        assert am.shape == ids.shape
        mask_out = torch.empty(B, T, dtype=torch.bool, device=dev)
        if obj_masks is not None:
            om = obj_masks.to(device=dev, dtype=torch.bool).contiguous().view(torch.uint8)
            assert om.numel() == n_rows
    with torch.cuda.device(dev):
        _C.call("msr3d_scene_rows_i64", ids.data_ptr(), B, T, int(scene_sp_token),
                am.data_ptr() if am is not None else None, om.data_ptr() if om is not None else None,
                rows.data_ptr(), n_rows, count.data_ptr(),
                mask_out.data_ptr() if mask_out is not None else None, _stream())
    return rows, count, mask_out


class _SpliceFn(torch.autograd.Function):
    """inputs_embeds[rows] = (obj_tokens @ W^T + b).to(inputs_embeds.dtype), in place."""

    @staticmethod
    def forward(ctx, inputs_embeds, obj_tokens, weight, bias, rows, split):
        D, K = weight.shape
        tok = obj_tokens.reshape(-1, K)
        if tok.dtype != torch.float32 or tok.stride(-1) != 1:
            tok = tok.float().contiguous()
        M = tok.shape[0]
        assert inputs_embeds.is_cuda and inputs_embeds.is_contiguous() and inputs_embeds.shape[-1] == D
        assert inputs_embeds.dtype in _DTYPES, f"unsupported embedding dtype {inputs_embeds.dtype}"
        assert weight.dtype == torch.float32 and weight.stride(1) == 1
        with torch.cuda.device(inputs_embeds.device):
            _gemm(tok.data_ptr(), (tok.stride(0), 1), weight.data_ptr(), (weight.stride(0), 1), M, D, K,
                  inputs_embeds.data_ptr(), D, bias=bias.data_ptr() if bias is not None else None, split=split,
                  row_map=rows.data_ptr(), out_dtype=_DTYPES[inputs_embeds.dtype])
        ctx.mark_dirty(inputs_embeds)
        ctx.save_for_backward(tok, weight, rows)
        ctx.has_bias = bias is not None
        ctx.tok_shape = obj_tokens.shape
        ctx.split = split
        return inputs_embeds

    @staticmethod
    def backward(ctx, grad):
        tok, weight, rows = ctx.saved_tensors
        D, K = weight.shape
        M = tok.shape[0]
        dev = grad.device
        g = grad.contiguous()
        dY = torch.empty(M, D, dtype=torch.float32, device=dev)
        d_embeds = d_tok = d_w = d_b = None
        with torch.cuda.device(dev):
            _C.call("msr3d_splice_gather_grad", g.data_ptr(), _DTYPES[g.dtype], D, rows.data_ptr(), M, D,
                    dY.data_ptr(), _stream())
            if ctx.needs_input_grad[1]:       # d tokens = dY . W            (contract over D)
                d_tok = torch.zeros(M, K, dtype=torch.float32, device=dev)
                ks = max(1, min(32, D // 128))
                _gemm(dY.data_ptr(), (D, 1), weight.data_ptr(), (1, weight.stride(0)), M, K, D, d_tok.data_ptr(), K,
                      split=ctx.split, accumulate=2, ksplit=ks)
                d_tok = d_tok.view(ctx.tok_shape)
            if ctx.needs_input_grad[2]:       # dW = dY^T . tokens           (contract over the M tokens)
                d_w = torch.empty(D, K, dtype=torch.float32, device=dev)
                _gemm(dY.data_ptr(), (1, D), tok.data_ptr(), (1, tok.stride(0)), D, K, M, d_w.data_ptr(), K, split=6)
            if ctx.has_bias and ctx.needs_input_grad[3]:
                d_b = torch.zeros(D, dtype=torch.float32, device=dev)
                _C.call("msr3d_colsum_add_f32", dY.data_ptr(), D, M, D, d_b.data_ptr(), None, None, _stream())
        if ctx.needs_input_grad[0]:           # overwritten rows pass nothing back to the embedding lookup
            d_embeds = g.clone()
            valid = rows[rows >= 0].long()
            d_embeds.view(-1, D)[valid] = 0
        return d_embeds, d_tok, d_w, d_b, None, None


def splice_scene_tokens(inputs_embeds: torch.Tensor, attention_mask: torch.Tensor, input_ids: torch.Tensor,
                        obj_tokens: torch.Tensor, obj_masks: torch.Tensor, weight: torch.Tensor,
                        bias: Optional[torch.Tensor], scene_sp_token: int = SCENE_SP_TOKEN, split: int = 6,
                        check: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Functional form.  `inputs_embeds` (B, T, D) is modified in place (as in the reference) and
    returned together with the spliced bool attention mask.  `check=True` verifies on the host that
    the prompts hold exactly one placeholder per scene token (the reference fails with a shape error
    in that case); pass False to stay asynchronous."""
    if not inputs_embeds.is_cuda:
        raise RuntimeError("msr3d_b200.splice: CUDA tensors only (no CPU path)")
    n_rows = obj_tokens.shape[0] * obj_tokens.shape[1]
    rows, count, am = scene_rows(input_ids, n_rows, attention_mask, obj_masks, scene_sp_token)
    if check:
        found = int(count.item())
        if found != n_rows:
            raise RuntimeError(f"shape mismatch: {found} scene placeholders (id {scene_sp_token}) in input_ids "
                               f"for {n_rows} scene tokens")
    out = _SpliceFn.apply(inputs_embeds, obj_tokens, weight, bias, rows, split)
    return out, am


class SceneSplice(nn.Module):
    """`llm_proj` + splice as a module; parameter names follow MSR3D (`llm_proj.weight/bias`)."""

    def __init__(self, in_features: int = 256, llm_hidden: int = 4096, scene_sp_token: int = SCENE_SP_TOKEN):
        super().__init__()
        self.llm_proj = nn.Linear(in_features, llm_hidden)      # model/msr3d/msr3d.py:84-87
        self.scene_sp_token = scene_sp_token

    def forward(self, inputs_embeds, attention_mask, input_ids, scene_dict, check: bool = True):
        return splice_scene_tokens(inputs_embeds, attention_mask, input_ids, scene_dict["obj_tokens"],
                                   scene_dict["obj_masks"], self.llm_proj.weight, self.llm_proj.bias,
                                   self.scene_sp_token, check=check)

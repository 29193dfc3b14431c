"""The fused kernels behind the drop-in module call.

`OSE3DSituation.forward` (prompter.py; reference model/ose3d_situation.py:288-454) is what
`MSR3D.forward` calls and what `trainer/leo_trainer.py:189-195` differentiates with `loss.backward()`.
When the configuration is one the fused engine covers (`fused.supported`), that call runs the engine's
forward program (frozen PointNet++ on the tensor cores, TMA-fed 3xTF32 GEMMs, attention / LayerNorm /
FFN-pair kernels) inside a `torch.autograd.Function`; its backward runs the engine's hand-written backward
program from the incoming `grad_output` and hands autograd one gradient per trainable parameter — so DDP
hooks, gradient accumulation, `clip_grad_norm_` and any torch optimiser see exactly what they would see
from the reference module tree, while no cuBLAS / ATen kernel is launched on the way.

Parameters stay `nn.Parameter`s; their storage is re-homed once into the engine's flat buffer (values and
state-dict keys unchanged) because the kernels read the four input projections of an attention block as
one (816 x 256) matrix.  Gradients are NOT aliased to the engine's buffer in this mode: each backward
returns fresh tensors and autograd accumulates them into `p.grad` as usual.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

_KEYS = ("obj_fts", "obj_locs", "obj_masks", "anchor_locs", "anchor_orientation")


def eligible(module, data_dict) -> Optional[str]:
    """None if this call can run on the fused engine, else the reason it cannot."""
    from . import fused
    if os.environ.get("MSR3D_MODULE_FUSED", "1") != "1" or not getattr(module, "fused_kernels", True):
        return "disabled"
    if any(k not in data_dict for k in _KEYS):
        return "inputs missing"
    why = fused.supported(module)
    if why is not None:
        return why
    fts = data_dict["obj_fts"]
    if not fts.is_cuda or fts.dim() != 4 or fts.shape[-1] != 6 or fts.dtype != torch.float32:
        return "obj_fts must be a (B, O, P, 6) fp32 CUDA tensor"
    if fts.shape[1] + (1 if module.situation_type in fused.ANCHOR_MODES else 0) > 128:
        return "more than 128 tokens per scene"
    p0 = next(module.parameters())
    if not p0.is_cuda or p0.dtype != torch.float32:
        return "parameters must be fp32 CUDA tensors"
    return None


def _engine(module):
    from .fused import FusedPrompter
    eng = getattr(module, "_fused_engine", None)
    if eng is None or eng.dev != next(module.parameters()).device:
        eng = FusedPrompter(module, alias_grads=False)
        object.__setattr__(module, "_fused_engine", eng)      # not a submodule / buffer: never in state_dict
    return eng


class _FusedPrompterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, engine, obj_fts, obj_locs, obj_masks, anchor_locs, anchor_ori, emb, *params):
        # emb: None (frozen encoder: computed here, without autograd) or the embeddings of an UNFROZEN encoder, computed
        # by the caller with autograd (encoder_train.py) — backward then hands d(emb) back to it
        B, L, P = obj_fts.shape[:3]
        engine.prepare(B, L, P, training=module.training)
        if module.training:
            engine.seed_counter.add_(1)            # fresh dropout masks per call (the engine's own optimiser is not used)
        w = engine.ws
        if emb is None:
            w["obj_fts"].copy_(obj_fts)
        w["obj_locs"].copy_(obj_locs)
        w["obj_masks"].copy_(obj_masks)
        w["anchor_locs"].copy_(anchor_locs)
        w["anchor_ori"].copy_(anchor_ori)
        g = engine.module_graph("fwd", emb is not None)
        if g is None:
            tokens = engine.forward(emb.detach() if emb is not None else None)
        else:                                   # the captured forward program (fused.py: module_graph)
            if emb is not None:
                w["emb"].copy_(emb.detach().reshape(w["emb"].shape))
            g.replay()
            tokens = engine._tokens()
        ctx.need_d_emb = emb is not None and emb.requires_grad
        ctx.external_emb = emb is not None
        ctx.engine = engine
        ctx.n_params = len(params)
        ctx.key = engine._key
        ctx.call_id = engine._module_calls = getattr(engine, "_module_calls", 0) + 1
        return tokens.clone()

    @staticmethod
    def backward(ctx, grad_tokens):
        e = ctx.engine
        if ctx.call_id != e._module_calls or ctx.key != e._key:
            raise RuntimeError("msr3d_b200: the fused prompter keeps ONE set of saved activations; call backward() "
                               "before the next forward of the same module (or set module.fused_kernels = False)")
        e.ws["upstream"].copy_(grad_tokens.reshape(e.ws["upstream"].shape))
        # engine-API engines alias p.grad to flat_g: keep what autograd has accumulated there so far
        saved = e.flat_g[:e.n_opt].clone() if e.alias_grads else None
        g = e.module_graph("bwd", ctx.external_emb)
        if g is None:
            e.flat_g.zero_()
            d_emb = e.backward(need_d_emb=ctx.need_d_emb)
        else:
            g.replay()
            d_emb = e.d_emb() if ctx.need_d_emb else None
        flat = e.flat_g[:e.n_opt].clone()          # fresh storage: autograd may keep / steal what we return
        if saved is not None:
            e.flat_g[:e.n_opt].copy_(saved)
        plan = e.__dict__.get("_grad_plan")
        if plan is None:                           # (offset, numel, shape) per autograd input, None = never touched
            plan = []
            for n in e.autograd_names:
                off, p = e.offs.get(n), e.by_name[n]
                plan.append(None if off is None or off >= e.n_opt else (off, p.numel(), p.shape))
            e.__dict__["_grad_plan"] = plan
        grads = [None if t is None else flat[t[0]:t[0] + t[1]].view(t[2]) for t in plan]
        if d_emb is not None:
            B, L = e.ws["obj_locs"].shape[:2]
            d_emb = d_emb.view(B, L, -1)
        return (None,) * 7 + (d_emb,) + tuple(grads)


def fused_forward(module, data_dict) -> torch.Tensor:
    """obj_tokens (B, O, 256) of `module(data_dict)` through the fused engine, differentiable w.r.t. the
    module's trainable parameters."""
    from .fused import encoder_frozen
    eng = _engine(module)
    params = [eng.by_name[n] for n in eng.autograd_names]
    emb = None
    if not encoder_frozen(module):
        # vision.args.freeze False (pcd_pointnet_encoder.py:58-62): the encoder's own autograd Function (training- or
        # eval-mode BatchNorm, encoder_train.py) in front of the transformer's
        emb = module.obj_encoder.encode(data_dict["obj_fts"])
    return _FusedPrompterFn.apply(module, eng, data_dict["obj_fts"], data_dict["obj_locs"], data_dict["obj_masks"],
                                  data_dict["anchor_locs"], data_dict["anchor_orientation"], emb, *params)


def token_masks(module, data_dict) -> torch.Tensor:
    """obj_masks of the call's output (ose3d_situation.py:452): the input's, with the agent token in front in the
    as_object / as_object_add_loc modes."""
    eng = _engine(module)
    return eng.token_masks().clone() if eng.n_anchor else data_dict["obj_masks"].clone()

"""Data-parallel plumbing for the hot path: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch on the GPU box, gloo in CPU tests).

The path shards by scenes — every cloud and every scene is independent — so the only
collective is the gradient exchange of the *trainable prompter parameters*: ONE all-reduce of the
flat fp32 gradient buffer per optimisation step (16.9 MB for the shipped configuration), replacing
DDP's bucketed reducer with find_unused_parameters=True (trainer/leo_trainer.py:50-52,135).
Parameters the shipped mode never touches live in a tail segment that is neither reduced nor
updated, which is what DDP's unused-parameter detection achieves with a graph walk per step."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist

UNUSED_IN_SHIPPED_MODE = ("anchor_feat", "orientation_encoder.", "loc_layers.")


def flat_order(names: Sequence[str], num_layers: int,
               unused: Sequence[str] = UNUSED_IN_SHIPPED_MODE) -> Tuple[List[str], List[str], List[str]]:
    """Order of the trainable parameters in the flat buffers: (weights, biases, never-used).  The four
    input projections of each attention block are adjacent so QKV+cond is a single GEMM.  Weights and
    biases are kept in separate ranges so that the OPT-IN no-decay rule (names containing 'bias',
    optim/utils.py:1-18 — a helper the reference imports but never calls) is one range too."""
    names = list(names)
    order_w, order_b, order_u, taken = [], [], [], set()
    for i in range(num_layers):
        pre = f"spatial_encoder.{i}.self_attn."
        for kind, dst in (("weight", order_w), ("bias", order_b)):
            for nm in ("w_qs", "w_ks", "w_vs", "lang_cond_fc"):
                key = f"{pre}{nm}.{kind}"
                if key in names:
                    dst.append(key)
                    taken.add(key)
    for n in names:
        if n in taken:
            continue
        if unused and n.startswith(tuple(unused)):
            order_u.append(n)
        elif "bias" in n:
            order_b.append(n)
        else:
            order_w.append(n)
    return order_w, order_b, order_u


def is_late(name: str) -> bool:
    """Parameters whose gradient is complete only at the very end of the backward pass: layer 0 of the
    spatial encoder and everything in front of it (projection, location / size encoders, embeddings)."""
    if name.startswith("ext."):     # parameters of the caller attached to the flat buffer (llm_proj, ...): their
        return False                # gradients come out of the LLM's backward, before the prompter's begins
    return (not name.startswith("spatial_encoder.")) or name.startswith("spatial_encoder.0.")


def flat_layout(names: Sequence[str], numel: Dict[str, int], num_layers: int, no_decay_bias: bool = False,
                unused: Sequence[str] = UNUSED_IN_SHIPPED_MODE) -> dict:
    """Layout of the flat parameter / gradient buffers:

        [ late biases | late weights | early weights | early biases | never-used ]

    so that (a) weight decay applies to ONE range, (b) the gradients that the backward pass finishes
    first (layers >= 1: "early") are ONE contiguous bucket whose all-reduce can run under the backward
    of layer 0, and the rest ("late") is one more, (c) the QKV+cond projections of a layer stay adjacent.
    Returns order, offs, total and the element boundaries decay=(d0,d1), late=(0,e0), early=(e0,n_opt).

    decay: the reference optimises ONE uniform parameter list (MSR3D.get_opt_params,
    model/msr3d/msr3d.py:165-172 -> AdamW(weight_decay=0.05), optim/build.py:7-17), i.e. biases and
    LayerNorm parameters decay too: decay = (0, n_opt).  no_decay_bias=True restricts it to the weight
    range (what optim/utils.py:no_decay_param_group would do if the reference called it)."""
    w, b, u = flat_order(names, num_layers, unused)
    b_late, b_early = [n for n in b if is_late(n)], [n for n in b if not is_late(n)]
    w_late, w_early = [n for n in w if is_late(n)], [n for n in w if not is_late(n)]
    order = b_late + w_late + w_early + b_early + u
    offs, total = flat_offsets(order, numel)

    def start(group, default):
        return offs[group[0]] if group else default
    n_opt = start(u, total)
    e1 = start(b_early, n_opt)           # end of the decayed range
    e0 = start(w_early, e1)              # start of the early bucket
    d0 = start(w_late, e0)
    decay = (d0, e1) if no_decay_bias else (0, n_opt)
    return {"order": order, "offs": offs, "total": total, "decay": decay, "weights": (d0, e1), "late": (0, e0),
            "early": (e0, n_opt), "n_opt": n_opt}


def flat_offsets(order: Sequence[str], numel: Dict[str, int], align: int = 4) -> Tuple[Dict[str, int], int]:
    """Element offsets of each tensor (starts aligned to 16 B) and the total length."""
    offs, off = {}, 0
    for n in order:
        offs[n] = off
        off += (numel[n] + align - 1) // align * align
    return offs, off


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def broadcast_(flat: torch.Tensor, src: int = 0) -> None:
    """Identical replicas at start (what DDP's constructor does)."""
    if world()[1] > 1:
        dist.broadcast(flat, src=src)


def allreduce_sum_(flat: torch.Tensor) -> None:
    """The gradient exchange: a single collective over the flat buffer.  The 1/world averaging is
    folded into the optimiser kernel's `grad_scale`."""
    if world()[1] > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)


def rank_seed(cfg_id: int, rank: int) -> int:
    """Synthetic-data seed per rank (cfg_id*1000 + rank)."""
    return cfg_id * 1000 + rank

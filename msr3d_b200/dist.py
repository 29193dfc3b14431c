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


def flat_order(names: Sequence[str], num_layers: int) -> Tuple[List[str], List[str], List[str]]:
    """Order of the trainable parameters in the flat buffers: (decayed weights, non-decayed biases,
    never-used).  The four input projections of each attention block are adjacent so QKV+cond is a
    single GEMM; the no-decay rule is the reference's (optim/utils.py:1-18: names containing 'bias')."""
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
        if n.startswith(UNUSED_IN_SHIPPED_MODE):
            order_u.append(n)
        elif "bias" in n:
            order_b.append(n)
        else:
            order_w.append(n)
    return order_w, order_b, order_u


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

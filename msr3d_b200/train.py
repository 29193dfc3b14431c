"""Hot-path training step: prompter forward/backward, gradient all-reduce of the trainable
prompter parameters over NCCL, global-norm clip and AdamW.

Mirrors what the reference trainer does around the path (trainer/leo_trainer.py:189-247:
zero_grad -> backward (DDP all-reduce, find_unused_parameters=True) -> clip_grad_norm_(5.0)
-> AdamW(lr 3e-5, wd 0.05, betas (0.9, 0.999)) per configs/msr3d.yaml:38-51), with the LLM
replaced by a fixed upstream gradient dL/d(obj_tokens): the frozen LLM is out of scope and
only ever hands the prompter that tensor.

Execution is the fused engine of msr3d_b200/fused.py: ~70 launches of this repo's kernels per
step, replayed as one CUDA graph.  Gradients live in ONE flat fp32 buffer, so the data-parallel
exchange is a single all-reduce and parameters that receive no gradient in the shipped mode
(anchor_feat, orientation_encoder.*, loc_layers.* — SURVEY.md §8a) need no unused-parameter
graph walk: they sit in a tail segment that neither the collective nor the optimiser touches."""
from __future__ import annotations

from typing import Optional

import torch

from . import synthetic
from .fused import FusedPrompter


class HotPathTrainer:
    """pipeline=True: the frozen encoder of batch i+1 runs concurrently with the transformer
    forward/backward/optimiser of batch i (one call = one encoder pass + one update, identical results,
    the update for a batch lands one call later; `flush()` drains the last one).
    pipeline="prefetch": additionally the point clouds of the batch handed to `step` are copied to the
    device on a copy stream while the current step runs (update lands two calls later)."""

    def __init__(self, model: torch.nn.Module, B: int, O: int, P: int, lr: float = 3e-5,
                 weight_decay: float = 0.05, betas=(0.9, 0.999), grad_norm: float = 5.0,
                 world_size: int = 1, split: int = 6, use_graph: bool = True, pipeline: bool = True,
                 chain_sms: Optional[int] = None, extra_params=None, accum_steps: int = 1):
        self.engine = FusedPrompter(model, lr=lr, weight_decay=weight_decay, betas=betas,
                                    grad_norm=grad_norm, world_size=world_size, split=split, chain_sms=chain_sms,
                                    extra_params=extra_params, accum_steps=accum_steps)
        assert accum_steps == 1 or pipeline == "prefetch", "gradient accumulation is built into the prefetch schedule"
        self.engine.prepare(B, O, P, training=model.training)
        self.use_graph = use_graph
        self.pipeline = pipeline
        self._calls = 0
        self.n_trainable = self.engine.n_trainable

    def warmup(self, batch: synthetic.SceneBatch, upstream: torch.Tensor):
        """First eager step (lazy kernel attribute setup), then graph capture (and pipeline priming)."""
        e = self.engine
        e.load_inputs(batch, upstream)
        # the eager step only warms the launch paths up: parameters, optimiser state and the dropout counter are put back
        # (the warm-up batch is not trained on twice)
        saved = [t.clone() for t in (e.flat_p, e.m, e.v, e.opt_state, e.seed_counter)]
        e.train_step_eager()
        for dst, src in zip((e.flat_p, e.m, e.v, e.opt_state, e.seed_counter), saved):
            dst.copy_(src)
        torch.cuda.synchronize()
        if self.pipeline == "prefetch":
            e.prime2(batch, upstream, batch, upstream)
            if self.use_graph:
                e.capture_prefetch()
            self._calls = 2
        elif self.pipeline:
            e.prime(batch, upstream)
            if self.use_graph:
                e.capture_pipelined()
        elif self.use_graph:
            e.capture()

    def step(self, batch: synthetic.SceneBatch, upstream: torch.Tensor,
             loss_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One micro-step fed with `batch` (device or pinned-host tensors).  Returns the surrogate loss
        <obj_tokens, upstream> (device scalar) of the batch whose update was applied in this call;
        `loss_out` (pinned host scalar) additionally receives it through an asynchronous D2H copy."""
        loss = self._step(batch, upstream)
        if loss_out is not None:
            if self.pipeline == "prefetch":
                # its own stream: on the H2D copy stream the 4-byte read-back (which waits for the end of this step)
                # sat in front of the next batch's point-cloud copy and pushed it a whole step back (+12 us per step,
                # tools/e2e_gap.py)
                ev = torch.cuda.Event()
                ev.record()
                cs = self.engine.__dict__.setdefault("d2h_stream", None) or torch.cuda.Stream(device=self.engine.dev)
                self.engine.d2h_stream = cs
                cs.wait_event(ev)
                with torch.cuda.stream(cs):
                    loss_out.copy_(loss, non_blocking=True)
            else:
                loss_out.copy_(loss, non_blocking=True)
        return loss

    def _step(self, batch, upstream):
        e = self.engine
        if not self.pipeline:
            e.load_inputs(batch, upstream)
            e.train_step()
            return e.loss()
        if self.pipeline == "prefetch":
            i = self._calls
            e.prefetch(batch, upstream, i % 2)              # async H2D, consumed by the NEXT call's step
            e.train_step_prefetch((i - 1) % 2)              # encode batch i-1, update with batch i-2
            self._calls += 1
            return e.loss()
        e.stage_points(batch)               # H2D of the point clouds: needed by the overlapped encoder
        e.train_step_pipelined()            # update with the previously staged batch
        loss = e.loss()
        e.stage_scene(batch, upstream)      # small per-scene inputs of `batch`, consumed by the next call
        return loss

    def flush(self):
        """Apply the updates still in the pipeline.  Returns the loss of the last one (pipeline=True) or
        the list of the last two (pipeline="prefetch")."""
        e = self.engine
        if self.pipeline == "prefetch":
            i = self._calls
            e.train_step_prefetch((i - 1) % 2, flush=1)     # encode the last batch, update with the one before
            l1 = e.loss().clone()
            e.train_step_prefetch(i % 2, flush=2)           # update with the last batch
            self._calls += 2
            return [l1, e.loss()]
        e.train_step_pipelined(flush=True)
        return e.loss()

"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of one hot-path training micro-step
(trainer/leo_trainer.py:189-247 around OSE3DSituation.forward): the oracle prompter
(modules_oracle.py + the C oracle for the native ops), torch autograd, global-norm clip 5.0 and
AdamW with the reference's hyper-parameters (configs/msr3d.yaml:38-51).  bench.py times this as
`cpu_baseline` (kind "port") and as the `--impl reference` arm; tests use it as the checker of
the CUDA training step."""
from __future__ import annotations

import torch

from . import modules_oracle as mo

FROZEN_PREFIXES = ("obj_encoder.",)


class CpuHotPath:
    def __init__(self, sd: dict, situation_type="as_transform_for_objects", lr=3e-5, weight_decay=0.05,
                 betas=(0.9, 0.999), grad_norm=5.0, no_decay_bias=False, device="cpu"):
        """device="cuda": the same restatement as stock eager torch ops on the GPU with the reference's own
        CUDA kernels (oracle/_ref) behind FPS / ball query — the GPU baseline of BASELINE.md §4."""
        self.sd = {}
        for k, v in sd.items():
            frozen = k.startswith(FROZEN_PREFIXES) or k == "anchor_size" or not v.is_floating_point()
            v = v.to(device)
            self.sd[k] = v.clone() if frozen else v.clone().requires_grad_(True)
        self.situation_type = situation_type
        self.grad_norm = grad_norm
        named = [(k, v) for k, v in self.sd.items() if v.requires_grad]
        if no_decay_bias:   # optim/utils.py:1-18 (dead code in the reference: imported, never called)
            nd = ("bias", "LayerNorm.bias", "LayerNorm.weight")
            groups = [{"params": [v for k, v in named if not any(s in k for s in nd)], "weight_decay": weight_decay},
                      {"params": [v for k, v in named if any(s in k for s in nd)], "weight_decay": 0.0}]
        else:               # MSR3D.get_opt_params (model/msr3d/msr3d.py:165-172): one uniform list
            groups = [{"params": [v for _, v in named], "weight_decay": weight_decay}]
        self.opt = torch.optim.AdamW(groups, lr=lr, betas=betas)
        self.trainable = [v for _, v in named]

    def forward(self, data: dict):
        return mo.ose3d_situation(self.sd, data, situation_type=self.situation_type)

    def step(self, data: dict, upstream: torch.Tensor, sync: bool = True):
        self.opt.zero_grad(set_to_none=True)
        tokens, _ = self.forward(data)
        loss = (tokens * upstream).sum()
        loss.backward()
        grads = [p.grad for p in self.trainable if p.grad is not None]
        if self.grad_norm is not None:
            total = torch.linalg.vector_norm(torch.stack([g.norm() for g in grads]))
            scale = torch.clamp(self.grad_norm / (total + 1e-6), max=1.0)
            for g in grads:
                g.mul_(scale)
        self.opt.step()
        return float(loss.detach()) if sync else loss.detach()

"""Generates tests/golden/encoder_train.npz from the UNMODIFIED reference `PcdObjEncoder(None, freeze=False)` in
training mode (modules/vision/pcd_pointnet_encoder.py:58-62): training-mode BatchNorm (batch statistics, running
statistics updated), gradients of every PointNet++ parameter through the reference's own autograd Functions
(pointnet2_utils.py GroupingOperation.backward -> _ext.group_points_grad, served by the C oracle).

Authoring-container only (needs /root/reference):  python tests/golden/gen_golden_train_encoder.py
Fixtures hold the seeds, the reference's outputs, gradients and updated running statistics — not the weights
(msr3d_b200.synthetic.make_state_dict regenerates them by name)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ext_oracle, refshim  # noqa: E402
from msr3d_b200 import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)
ext_oracle.set_threads(8)


def main():
    ref = refshim.load()
    fix = {}
    for case, (B, O, P, seed, bn_train) in {"train": (2, 5, 256, 6006, True), "evalbn": (1, 6, 128, 6007, False)}.items():
        # explicit lists: the reference mutates its default `sa_mlps` argument on every construction (pointnet2_modules.py:121)
        enc = ref.pcd_pointnet_encoder.PcdObjEncoder(
            None, sa_mlps=[[3, 64, 64, 128], [128, 128, 128, 256], [256, 256, 512, 768]], freeze=False)
        spec = [(k, list(v.shape), "int64" if v.dtype == torch.int64 else "float32") for k, v in enc.state_dict().items()]
        enc.load_state_dict(synthetic.make_state_dict(spec, seed), strict=True)
        enc.train(bn_train)          # evalbn: unfrozen parameters, BatchNorm on running statistics
        batch = synthetic.make_batch(B, O, P, seed=seed)
        emb, _ = enc(batch.obj_fts)
        G = torch.randn(emb.shape, generator=torch.Generator().manual_seed(seed + 77))
        (emb * G).sum().backward()
        fix[f"{case}:meta"] = np.array([B, O, P, seed, int(bn_train)])
        fix[f"{case}:emb"] = emb.detach().numpy()
        for n, p in enc.pcd_net.named_parameters():
            g = p.grad.numpy()
            if g.size <= 16384:
                fix[f"{case}:grad:{n}"] = g.copy()
            else:            # large tensors: every 37th element, the norm and the sum (keeps the fixture small)
                fix[f"{case}:gradsub:{n}"] = g.reshape(-1)[::37].copy()
                fix[f"{case}:gradstat:{n}"] = np.array([np.linalg.norm(g.astype(np.float64)), g.astype(np.float64).sum()])
        for n, b in enc.pcd_net.named_buffers():
            fix[f"{case}:buf:{n}"] = b.detach().numpy().copy()
        print(case, "emb", tuple(emb.shape), "abs-mean %.4f" % emb.abs().mean().item(),
              "params with grad", sum(1 for _ in enc.pcd_net.parameters()))
    np.savez_compressed(os.path.join(HERE, "encoder_train.npz"), **fix)


if __name__ == "__main__":
    main()

"""Kernel-time breakdown of one forward + backward of the unfrozen encoder at the bench configuration's clouds
(torch.profiler / CUPTI; development aid for DESIGN 5.2)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msr3d_b200 import synthetic  # noqa: E402
from msr3d_b200.encoder import PcdObjEncoder  # noqa: E402


def main():
    nums = [a for a in sys.argv[1:] if a.isdigit()]
    B, O, P = (int(a) for a in nums[:3]) if len(nums) >= 3 else (4, 60, 1024)
    enc = PcdObjEncoder(freeze=False)
    spec = [(k, list(v.shape), "int64" if v.dtype == torch.int64 else "float32") for k, v in enc.state_dict().items()]
    enc.load_state_dict(synthetic.make_state_dict(spec, 7))
    enc = enc.cuda().train()
    batch = synthetic.make_batch(B, O, P, seed=7).to("cuda")
    G = torch.randn(B, O, 768, device="cuda")
    if "--plain" in sys.argv:            # one forward + backward without CUPTI (to be wrapped in ncu)
        enc.encode(batch.obj_fts).backward(G)
        torch.cuda.synchronize()
        return
    for _ in range(3):
        enc.encode(batch.obj_fts).backward(G)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            enc.encode(batch.obj_fts).backward(G)
        torch.cuda.synchronize()
    rows = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(e.name[:90], [0, 0.0])
            r[0] += 1
            r[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    tot = sum(v[1] for v in rows.values())
    print(f"total device time per fwd+bwd: {tot / 3 / 1e3:.3f} ms over {sum(v[0] for v in rows.values()) // 3} launches")
    for k, (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{t / 3:9.1f} us  x{n // 3:3d}  {k}")


if __name__ == "__main__":
    main()

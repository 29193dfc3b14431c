"""Timing of the fused clip + AdamW launch with / without in-kernel gradient zeroing, and of the staging launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from msr3d_b200 import configs, synthetic  # noqa: E402
from msr3d_b200.fused import FusedPrompter  # noqa: E402

B, O, P = synthetic.CONFIGS["cfg2_60x1024"]
model = configs.build_prompter(0, device="cuda").train()
eng = FusedPrompter(model)
eng.prepare(B, O, P, training=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, n=20, cold=True):
    ts = []
    for _ in range(n):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for cold in (True, False):
    for zg in (False, True):
        eng.flat_g.normal_()
        print(f"optimizer_step zero_grads={zg} cold={cold}: {t(lambda: eng.optimizer_step(zero_grads=zg), cold=cold):.1f} us")
    print(f"optimizer_step split=True cold={cold}: {t(lambda: eng.optimizer_step(split=True), cold=cold):.1f} us")
    import ctypes
    from msr3d_b200 import _C
    d = eng._stage_desc(0)
    print(f"stage_step cold={cold}: {t(lambda: _C.lib().msr3d_stage_step(ctypes.byref(d), torch.cuda.current_stream().cuda_stream), cold=cold):.1f} us")
    print(f"flat_g.zero_ cold={cold}: {t(lambda: eng.flat_g.zero_(), cold=cold):.1f} us")

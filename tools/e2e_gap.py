"""Where the gap between `value` (inputs resident) and `e2e` (pinned host inputs + loss read-back) of bench.py comes
from: the same timed loop with each of the two host-facing pieces switched on separately (development aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import msr3d_b200  # noqa: E402
from msr3d_b200 import synthetic  # noqa: E402
from msr3d_b200.train import HotPathTrainer  # noqa: E402


def main():
    dev = torch.device("cuda")
    B, O, P = synthetic.CONFIGS["cfg2_60x1024"]
    model = msr3d_b200.build_model(bench.prompter_cfg())
    model.load_state_dict(synthetic.make_state_dict(bench.load_spec(), 0), strict=True)
    model = model.to(dev).train()
    tr = HotPathTrainer(model, B, O, P, pipeline="prefetch", chain_sms=80)
    host = [synthetic.make_batch(B, O, P, seed=2000 + 17 * i).pin() for i in range(4)]
    devb = [h.to(dev) for h in host]
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(1)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    tr.warmup(devb[0], up)
    for i in range(10):
        tr.step(devb[i % 4], up)

    def timed(n, from_host, with_loss):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for i in range(n):
            flush.zero_()
            ev[i][0].record()
            tr.step((host if from_host else devb)[i % 4], up, loss_out=loss_host if with_loss else None)
            ev[i][1].record()
        torch.cuda.synchronize()
        ts = [a.elapsed_time(b) for a, b in ev]
        return sum(ts) / n

    # how long the point-cloud copy itself takes (the prefetch has one step of time for it)
    cs = torch.cuda.Stream()
    dst = torch.empty_like(devb[0].obj_fts)
    ts = []
    for i in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(cs):
            a.record()
            dst.copy_(host[i % 4].obj_fts, non_blocking=True)
            b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    nb = dst.numel() * 4
    print(f"pinned H2D of obj_fts ({nb / 1e6:.1f} MB), GPU idle: {sorted(ts)[len(ts) // 2]:.1f} us = {nb / sorted(ts)[len(ts) // 2] / 1e3:.1f} GB/s")
    for rep in range(2):
        for name, fh, wl in (("resident, no loss copy", False, False), ("resident + loss copy", False, True),
                             ("pinned host, no loss copy", True, False), ("pinned host + loss copy (= e2e)", True, True)):
            timed(10, fh, wl)
            print(f"{name:34s} {timed(100, fh, wl) * 1e3:8.1f} us / step")


if __name__ == "__main__":
    main()

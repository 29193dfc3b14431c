"""Timings of the "next row" kernels that are not part of the bench step (on the GPU box):
device-side object preprocessing and the deterministic scatter gradients (8f-2), with the
CPU oracle / reference arithmetic timed beside them.  CUDA events, L2 flushed before every launch, median of 15.
    python tools/next_rows_bench.py > gpurun_out/next_rows.json"""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from msr3d_b200 import data, ext  # noqa: E402
from oracle import data_oracle, ext_oracle  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=15):
    ts = []
    for i in range(n + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


res = {}
# ---- preprocessing: one scene of 54 objects, 500 .. 40000 points each -> 60 x 1024
rng = np.random.RandomState(0)
sizes = list(rng.randint(500, 40000, size=54))
points = np.concatenate([np.concatenate([rng.randn(n, 3), rng.uniform(-1, 1, (n, 3))], 1) for n in sizes]).astype(np.float32)
offsets = np.cumsum([0] + sizes).astype(np.int32)
idx = np.stack([rng.choice(n, size=1024, replace=n < 1024) for n in sizes]).astype(np.int32)
rot = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], dtype=np.float32)
p_d, o_d, i_d, r_d = (torch.from_numpy(a).to(dev) for a in (points, offsets, idx, rot))
us = timed(lambda: data.preprocess_objects(p_d, o_d, i_d, r_d, 60))
bytes_alg = points.nbytes + 60 * 1024 * 24 + idx.nbytes + 54 * 1024 * 24
t0 = time.perf_counter()
for _ in range(3):
    data_oracle.preprocess_objects(points, offsets, idx, rot, 60)
cpu_us = (time.perf_counter() - t0) / 3 * 1e6
res["preprocess_objects (54 objects, %.1f M raw points -> 60 x 1024)" % (points.shape[0] / 1e6)] = {
    "us": round(us, 1), "algorithmic_bytes": int(bytes_alg), "gbs": round(bytes_alg / us / 1e3, 1),
    "cpu_oracle_us": round(cpu_us, 1), "note": "eager call through the Python binding (allocations + one launch)"}
# ---- scatter gradients at SA1 / SA2 shapes of config 2
for name, (B, C, N, P, S) in {"group_points_grad SA1 (240 x 6 x 1024, 32 x 32)": (240, 6, 1024, 32, 32),
                              "group_points_grad SA2 (240 x 131 x 32, 16 x 32)": (240, 131, 32, 16, 32)}.items():
    g = torch.Generator().manual_seed(1)
    idx3 = torch.randint(0, N, (B, P, S), generator=g, dtype=torch.int32)
    go = torch.randn(B, C, P, S, generator=g)
    idx_d, go_d = idx3.to(dev), go.to(dev)
    us = timed(lambda: ext.group_points_grad(go_d, idx_d, N))
    b_alg = go.numel() * 4 + idx3.numel() * 4 + B * C * N * 4
    t0 = time.perf_counter()
    ext_oracle.group_points_grad(go, idx3, N)
    cpu_us = (time.perf_counter() - t0) * 1e6
    res[name] = {"us": round(us, 1), "algorithmic_bytes": int(b_alg), "gbs": round(b_alg / us / 1e3, 1),
                 "cpu_oracle_us": round(cpu_us, 1)}
print(json.dumps(res, indent=1))

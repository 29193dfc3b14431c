"""Seeded inputs shared by tests/golden/gen_ext_refcuda.py (reference CUDA, on the GPU box)
and tests/test_oracle_refcuda.py (C oracle, CPU)."""
import torch

from msr3d_b200 import synthetic


def _cloud(B, N, seed, dups=True, tiny=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, 3, generator=g) * 0.35
    if dups:
        for b in range(0, B, 2):
            uniq = min(N, max(2, N // 4))
            x[b] = x[b, :uniq][torch.randint(0, uniq, (N,), generator=g)]
    if tiny and B >= 3:
        x[-1] = 1.0
        x[-2, : N // 3] *= 0.01
    return x.contiguous()


def fps_cases():
    yield "n1024_m32", _cloud(8, 1024, 1), 32
    yield "n32_m16", _cloud(8, 32, 2), 16
    yield "n256_m32", _cloud(4, 256, 3), 32
    yield "n2048_m32", _cloud(4, 2048, 4), 32
    yield "n100_m20", _cloud(4, 100, 5), 20
    yield "n700_m64", _cloud(3, 700, 6), 64
    b = synthetic.make_batch(1, 20, 1024, seed=7)
    yield "scene_n1024_m32", b.obj_fts[0, :, :, :3].contiguous(), 32


def ball_cases():
    for name, x, m in fps_cases():
        N = x.shape[1]
        M = min(m, N)
        perm = torch.randperm(N, generator=torch.Generator().manual_seed(N + M))[:M]
        c = x[:, perm].contiguous()
        c[:, 0] += 5.0
        yield name + "_r0.2", c, x, 0.2, 32
    g = torch.Generator().manual_seed(5)
    N, M, r = 4096, 8, 0.2
    d = torch.randn(1, N, 3, generator=g)
    d = d / d.norm(dim=2, keepdim=True) * r * (1 + (torch.rand(1, N, 1, generator=g) - 0.5) * 4e-7)
    c = torch.rand(1, M, 3, generator=g)
    yield "on_sphere", c.contiguous(), (c[:, :1] + d).contiguous(), r, 64


def nn_cases():
    g = torch.Generator().manual_seed(13)
    u = torch.randn(2, 150, 3, generator=g)
    k = torch.randn(2, 40, 3, generator=g)
    k[:, 5] = k[:, 4]
    yield "n150_m40", u, k

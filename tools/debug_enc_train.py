"""Per-parameter comparison of the kernel training path of the unfrozen encoder (msr3d_b200/encoder_train.py) with the
layer-wise torch module tree on the same GPU (development aid)."""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from msr3d_b200 import synthetic  # noqa: E402
from msr3d_b200.encoder import PcdObjEncoder  # noqa: E402


def build(seed, train=True):
    enc = PcdObjEncoder(freeze=False)
    spec = [(k, list(v.shape), "int64" if v.dtype == torch.int64 else "float32") for k, v in enc.state_dict().items()]
    enc.load_state_dict(synthetic.make_state_dict(spec, seed), strict=True)
    return enc.cuda().train(train)


def main():
    # the module tree is the comparison: keep cuDNN / cuBLAS off TF32 (default on for convolutions), or arg-max /
    # ReLU decisions flip on ITS side
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, O, P = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 10, 256)
    train = (sys.argv[4] != "eval") if len(sys.argv) > 4 else True
    a, b = build(7, train), build(7, train)
    b.pcd_net.train_kernels = False
    batch = synthetic.make_batch(B, O, P, seed=7).to("cuda")
    G = torch.randn(B, O, 768, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    ea, eb = a.encode(batch.obj_fts), b.encode(batch.obj_fts)
    (ea * G).sum().backward()
    (eb * G).sum().backward()
    print("emb rel err", float((ea - eb).abs().max() / eb.abs().max()))
    for (n, p), (_, q) in zip(a.pcd_net.named_parameters(), b.pcd_net.named_parameters()):
        err = float((p.grad - q.grad).abs().max() / q.grad.abs().max().clamp_min(1e-12))
        nerr = float((p.grad - q.grad).norm() / q.grad.norm().clamp_min(1e-20))
        print(f"{n:45s} |g| {float(q.grad.abs().max()):10.3e}  max err {err:9.2e}  norm err {nerr:9.2e}")
    import time
    for name, enc in (("kernels", a), ("module tree", b)):
        ts = []
        for i in range(5):
            for p in enc.parameters():
                p.grad = None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            (enc.encode(batch.obj_fts) * G).sum().backward()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"{name}: fwd + bwd {min(ts):.3f} ms (best of 5, wall clock)")
    for (n, p), (_, q) in zip(a.pcd_net.named_buffers(), b.pcd_net.named_buffers()):
        if p.is_floating_point():
            print(f"{n:45s} buffer rel err {float((p - q).abs().max() / q.abs().max().clamp_min(1e-12)):9.2e}")


if __name__ == "__main__":
    main()

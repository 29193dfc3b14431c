"""Probe: can the SMs be partitioned between the transformer chain and the frozen encoder with CUDA green
contexts (driver API through cuda-python), with torch streams wrapped around them and the whole thing captured
in one CUDA graph?   python tools/green_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from cuda.bindings import driver as drv  # noqa: E402


def chk(r):
    err = r[0]
    if err != drv.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"driver error {err}")
    return r[1] if len(r) == 2 else r[1:]


def make_partition(dev_index: int, n_a: int):
    """Two green contexts: `n_a` SMs (rounded by the driver) and the remainder."""
    torch.cuda.init()
    torch.zeros(1, device=f"cuda:{dev_index}")
    dev = chk(drv.cuDeviceGet(dev_index))
    res = chk(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    print("total SMs", res.sm.smCount)
    out = drv.cuDevSmResourceSplitByCount(1, res, 0, n_a)
    err, groups, nb, remaining = out
    assert err == drv.CUresult.CUDA_SUCCESS, err
    ga = groups[0]
    print("partition A", ga.sm.smCount, "remaining", remaining.sm.smCount)
    ctxs, streams = [], []
    for r in (ga, remaining):
        desc = chk(drv.cuDevResourceGenerateDesc([r], 1))
        g = chk(drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        s = chk(drv.cuGreenCtxStreamCreate(g, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        ctxs.append(g)
        streams.append(torch.cuda.ExternalStream(int(s), device=f"cuda:{dev_index}"))
    return ctxs, streams, (ga.sm.smCount, remaining.sm.smCount)


if __name__ == "__main__":
    ctxs, (sa, sb), counts = make_partition(0, 64)
    x = torch.randn(8192, 8192, device="cuda")
    y = torch.randn(8192, 8192, device="cuda")

    def work(n=4):
        z = x
        for _ in range(n):
            z = z @ y
        return z

    def timeit(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    work()
    print("full GPU        : %.2f ms" % timeit(work))
    for name, s in (("partition A", sa), ("partition B", sb)):
        def run():
            cur = torch.cuda.current_stream()
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                work()
            cur.wait_stream(s)
        run()
        print(f"{name} ({counts[0] if s is sa else counts[1]} SMs): %.2f ms" % timeit(run))

    def both():
        cur = torch.cuda.current_stream()
        sa.wait_stream(cur); sb.wait_stream(cur)
        with torch.cuda.stream(sa):
            work()
        with torch.cuda.stream(sb):
            work()
        cur.wait_stream(sa); cur.wait_stream(sb)
    both()
    print("A and B together: %.2f ms" % timeit(both))
    # graph capture across the two partitions
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    try:
        with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
            both()
        g.replay()
        print("graph replay    : %.2f ms" % timeit(g.replay))
        # our own kernels inside a partition
        from msr3d_b200.fused import gemm_tc
        A = torch.randn(240, 256, device="cuda"); W = torch.randn(816, 256, device="cuda")
        with torch.cuda.stream(sa):
            o = gemm_tc(A, W, impl="tf32x3", tile_n=64)
        torch.cuda.synchronize()
        print("gemm3x in partition A rel err", float((o.double() - A.double() @ W.double().t()).abs().max()))
    except Exception as e:  # noqa: BLE001
        print("FAILED:", type(e).__name__, e)

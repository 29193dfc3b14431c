"""Launch one kernel configuration a few times (to be run under ncu on the GPU box).
    ncu --set full --import-source on -k regex:linear_tc -s 2 -c 1 -o gpurun_out/prof python tools/profile_one.py gemm 240 256 4096
    ncu ... -k regex:attn_bwd ... python tools/profile_one.py attn_bwd"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools import gemm_sweep as gs  # noqa: E402
from msr3d_b200 import _C, fused  # noqa: E402

what = sys.argv[1]
st = torch.cuda.current_stream().cuda_stream
if what == "gemm":
    M, N, K = (int(x) for x in sys.argv[2:5])
    bn = int(sys.argv[5]) if len(sys.argv) > 5 else 64
    ks = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    gs.tune(bn=bn)
    la = gs.gemm_launcher(M, N, K, 6, accumulate=2 if ks > 1 else 0, ksplit=ks)
    for _ in range(4):
        la(st)
else:
    M, H, Q = 240, 256, fused.QKVC
    B, L = 4, 60
    dev = gs.dev
    qkvc = torch.randn(M, Q, device=dev)
    pw = torch.rand(B, L, L, 5, device=dev)
    pad = torch.zeros(B, L, dtype=torch.uint8, device=dev)
    P = torch.zeros(B, 8, L, L, device=dev)
    LOC = torch.zeros(B, 8, L, L, device=dev)
    ctx = torch.zeros(M, H, device=dev)
    dqkvc = torch.zeros(M, Q, device=dev)
    lib = _C.lib()
    for _ in range(4):
        lib.msr3d_attn_fwd_f32(qkvc.data_ptr(), Q, pw.data_ptr(), pad.data_ptr(), B, L, 8, P.data_ptr(),
                               LOC.data_ptr(), ctx.data_ptr(), H, st)
        lib.msr3d_attn_bwd_f32(ctx.data_ptr(), H, qkvc.data_ptr(), Q, pw.data_ptr(), P.data_ptr(), LOC.data_ptr(),
                               B, L, 8, dqkvc.data_ptr(), st)
torch.cuda.synchronize()
print("done")

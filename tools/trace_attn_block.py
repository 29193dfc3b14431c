"""Phase timeline (SM clocks written by the kernel) and in-chain latency of the one-launch attention block
(csrc/attn_block.cu).   python tools/trace_attn_block.py   (on the GPU box)"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from msr3d_b200 import _C, configs, synthetic  # noqa: E402
from msr3d_b200.fused import FusedPrompter  # noqa: E402
from tools.trace_kernels import chain_us  # noqa: E402

MHZ = 1965.0
SLOTS = ["entry", "setup", "pdl_wait", "chunk0_landed", "conv_done", "acc1_seen", "qk_tiles_written", "softmax_done", "acc3_seen",
         "pushed", "cluster_sync2", "tail_done", "end", "pair_done", "loc_table_seen", "accS_seen"]
lib = _C.lib()
lib.msr3d_set_trace_buffer.argtypes = [ctypes.c_void_p]
B, O, P = synthetic.CONFIGS["cfg2_60x1024"]
model = configs.build_prompter(0, device="cuda").train()
eng = FusedPrompter(model)
eng.prepare(B, O, P, training=True)
batch = synthetic.make_batch(B, O, P, seed=1).to("cuda")
eng.load_inputs(batch, torch.randn(B, O, 256, device="cuda"))
eng.train_step_eager()
torch.cuda.synchronize()
name, fn, args, tag, side = next(e for e in eng.fwd_prog if e[0].startswith("msr3d_attn_block_fwd"))


def launch(st):
    assert fn(*args, st) == 0, lib.msr3d_last_error()


print(f"{chain_us(launch):7.2f} us in-chain  attn_block fwd  [{8 * B} CTAs]")
n_cta = 8 * B
buf = torch.zeros(n_cta * 16, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    launch(st)
torch.cuda.synchronize()
lib.msr3d_set_trace_buffer(buf.data_ptr())
launch(st)
torch.cuda.synchronize()
lib.msr3d_set_trace_buffer(None)
t = buf.view(n_cta, 16).cpu().double()
rel = (t - t[:, :1]) / MHZ
for i, nm in enumerate(SLOTS):
    col = rel[:, i][t[:, i] > 0]
    if len(col):
        print(f"  {nm:14s} mean {col.mean():6.2f}  min {col.min():6.2f}  max {col.max():6.2f}")

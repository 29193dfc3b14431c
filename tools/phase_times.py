"""Phase-level timing of the fused step: each phase captured as its own CUDA graph and replayed.
    python tools/phase_times.py            (on the GPU box)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msr3d_b200
from msr3d_b200 import synthetic
from msr3d_b200.fused import FusedPrompter
from msr3d_b200 import configs as helpers

B, O, P = synthetic.CONFIGS["cfg2_60x1024"]
model = helpers.build_prompter(0, device="cuda").train()
eng = FusedPrompter(model)
eng.prepare(B, O, P, training=True)
batch = synthetic.make_batch(B, O, P, seed=1).to("cuda")
up = torch.randn(B, O, 256, device="cuda")
eng.load_inputs(batch, up)
eng.train_step_eager()
torch.cuda.synchronize()


def graph_of(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        fn()
    return g


def timeit(g, n=200):
    for _ in range(10):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def fwd_tr():
    torch.logical_not(eng.ws["obj_masks"], out=eng.ws["key_pad"].view(torch.bool))
    eng._run(eng.pre_prog)
    eng._run(eng.fwd_prog)


phases = {
    "encoder": lambda: eng._encode(),
    "transformer_fwd": fwd_tr,
    "backward": lambda: (eng.flat_g.zero_(), eng.backward()),
    "optimizer": eng.optimizer_step,
    "inputs+loss": lambda: (eng.load_inputs(batch, up), eng.loss()),
}
res = {}
for use_side, small in ((True, True), (True, False), (False, True)):
    eng.use_side, eng.wgrad_small = use_side, small
    eng._key = None
    eng.prepare(B, O, P, training=True)          # rebuild the launch lists with the new policy
    for k, fn in phases.items():
        label = k + ("" if use_side else " (no side stream)") + ("" if small else " (wgrad split-K/64-wide)")
        res[label] = round(timeit(graph_of(fn)), 1)
print(json.dumps(res, indent=1))

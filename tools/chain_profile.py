"""Timeline of one graph-replayed training step from CUPTI (torch.profiler): start offset and duration of every
kernel, i.e. the real chain including inter-kernel gaps.   python tools/chain_profile.py [fwd|bwd|step]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from msr3d_b200 import configs, synthetic  # noqa: E402
from msr3d_b200.fused import FusedPrompter  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "step"
B, O, P = synthetic.CONFIGS["cfg2_60x1024"]
model = configs.build_prompter(0, device="cuda").train()
eng = FusedPrompter(model)
eng.prepare(B, O, P, training=True)
batch = synthetic.make_batch(B, O, P, seed=1).to("cuda")
up = torch.randn(B, O, 256, device="cuda")
eng.load_inputs(batch, up)
eng.train_step_eager()
torch.cuda.synchronize()


def fwd_tr():
    torch.logical_not(eng.ws["obj_masks"], out=eng.ws["key_pad"].view(torch.bool))
    eng._run(eng.pre_prog)
    eng._run(eng.fwd_prog)


fn = {"fwd": fwd_tr, "bwd": lambda: (eng.flat_g.zero_(), eng.backward()),
      "step": lambda: (eng.flat_g.zero_(), eng.forward(), eng.backward(), eng.optimizer_step())}[what]
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    fn()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, capture_error_mode="thread_local"):
    fn()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
end_prev = t0
print(f"{'start':>8} {'dur':>7} {'gap':>6}  kernel")
for e in ev:
    st, en = e.time_range.start, e.time_range.end
    print(f"{st - t0:8.1f} {en - st:7.1f} {st - end_prev:6.1f}  {e.name[:90]}")
    end_prev = max(end_prev, en)
print(f"total {end_prev - t0:.1f} us, kernel time {sum(e.time_range.end - e.time_range.start for e in ev):.1f} us")

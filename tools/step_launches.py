"""One eager training step between cudaProfilerStart/Stop (run under ncu on the GPU box):
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/step_launches.py
Every launch of the step with its device time (serialised, cold cache: compare shares, not absolutes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from msr3d_b200 import synthetic  # noqa: E402
from msr3d_b200.fused import FusedPrompter  # noqa: E402
from msr3d_b200.configs import build_prompter as build_default_prompter  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2_60x1024"
B, O, P = synthetic.CONFIGS[cfg]
model = build_default_prompter(0, device="cuda").train()
eng = FusedPrompter(model)
eng.use_side = False          # one stream: the launch list is the program order
eng.prepare(B, O, P, training=True)
batch = synthetic.make_batch(B, O, P, seed=1).to("cuda")
up = torch.randn(B, O, 256, device="cuda")
eng.load_inputs(batch, up)
for _ in range(3):
    eng.train_step_eager()
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.train_step_eager()
torch.cuda.synchronize()
torch.cuda.profiler.stop()

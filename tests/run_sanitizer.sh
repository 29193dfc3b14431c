#!/bin/bash
# compute-sanitizer pass over one tiny hot-path training step (run on the GPU box):
#   gpurun -- bash tests/run_sanitizer.sh
# memcheck finds out-of-bounds / misaligned accesses in every kernel of the step; racecheck covers
# the shared-memory hand-offs (operand images, FPS partials, attention staging).
set -u
mkdir -p gpurun_out
SCRIPT='
import torch, msr3d_b200
from msr3d_b200 import synthetic
from msr3d_b200.fused import FusedPrompter
from tests import helpers
model = helpers.build_prompter(3, device="cuda"); model.train()
eng = FusedPrompter(model); eng.prepare(1, 8, 256, training=True)
b = synthetic.make_batch(1, 8, 256, seed=3).to("cuda")
eng.load_inputs(b, torch.randn(1, 8, 256, device="cuda"))
eng.train_step_eager(); torch.cuda.synchronize(); print("step ok", float(eng.loss()))
'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "$SCRIPT" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|step ok|Error|hazard" gpurun_out/sanitizer_$tool.log | head -8
done

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
# prefetch schedule, eager (staging launch, geometry beside the encoder, AdamW writing the tf32 weight images)
from msr3d_b200.train import HotPathTrainer
m2 = helpers.build_prompter(4, device="cuda"); m2.train()
tr = HotPathTrainer(m2, 1, 8, 256, use_graph=False, pipeline="prefetch", chain_sms=0)
up = torch.randn(1, 8, 256, device="cuda")
tr.warmup(b, up)
for _ in range(3):
    tr.step(b, up)
tr.flush(); torch.cuda.synchronize(); print("prefetch steps ok", float(tr.engine.loss()))
# scene-token splice: index scan, scattered fp16 epilogue, backward gather + GEMMs
from msr3d_b200 import splice
ids = torch.randint(0, 30000, (1, 40), device="cuda"); ids[0, 5:13] = splice.SCENE_SP_TOKEN
W = torch.randn(96, 256, device="cuda", requires_grad=True); bias = torch.zeros(96, device="cuda", requires_grad=True)
tok = torch.randn(1, 8, 256, device="cuda", requires_grad=True)
out, am = splice.splice_scene_tokens(torch.zeros(1, 40, 96, device="cuda", dtype=torch.float16),
                                     torch.ones(1, 40, dtype=torch.long, device="cuda"), ids, tok,
                                     torch.ones(1, 8, dtype=torch.bool, device="cuda"), W, bias)
out.float().sum().backward(); torch.cuda.synchronize(); print("splice ok", float(W.grad.abs().sum()))
# deterministic scatter gradients (shared-memory CSR) and device-side object preprocessing
from msr3d_b200 import ext, data
idx3 = torch.randint(0, 33, (2, 5, 3), device="cuda", dtype=torch.int32)
g = ext.group_points_grad(torch.randn(2, 7, 5, 3, device="cuda"), idx3, 33)
pts = torch.randn(300, 6, device="cuda")
fts, locs, masks = data.preprocess_objects(pts, torch.tensor([0, 100, 300], dtype=torch.int32),
                                           torch.randint(0, 100, (2, 64), dtype=torch.int32), torch.eye(3), 4)
torch.cuda.synchronize(); print("next rows ok", float(g.abs().sum()), float(fts.abs().sum()))
'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "$SCRIPT" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|step ok|Error|hazard" gpurun_out/sanitizer_$tool.log | head -8
done

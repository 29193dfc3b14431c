#!/bin/bash
# compute-sanitizer over the kernels added for the "next" rows in round 2 (run on the GPU box):
#   gpurun -- bash tests/run_sanitizer_rows.sh
# fusion-templated attention + spatial-weight kernels (csrc/spatial.cu), agent-token input assembly (csrc/modes.cu),
# unfrozen-encoder training kernels (csrc/sa_train.cu) with the TMA-fed GEMMs around them.
set -u
mkdir -p gpurun_out
SCRIPT='
import torch
from msr3d_b200 import configs, synthetic
from msr3d_b200.fused import FusedPrompter
def run(model, B, O, P, extra=0):
    eng = FusedPrompter(model); eng.prepare(B, O, P, training=False)
    b = synthetic.make_batch(B, O, P, seed=3).to("cuda")
    eng.load_inputs(b, torch.randn(B, O + extra, 256, device="cuda"))
    eng.flat_g.zero_(); eng.forward(); eng.backward(); eng.optimizer_step(); torch.cuda.synchronize()
    return float(eng.flat_g.abs().sum())
for fusion in ("mul", "ctx", "add"):
    print("fusion", fusion, run(configs.build_prompter(3, device="cuda", fusion=fusion), 1, 6, 128))
for st in ("as_object", "as_object_add_loc"):
    print(st, run(configs.build_prompter(3, st, device="cuda"), 2, 5, 128, extra=1))
for st in ("as_embedding", "as_cross_attention"):
    print(st, run(configs.build_prompter(3, st, device="cuda"), 2, 5, 128))
from msr3d_b200.encoder import PcdObjEncoder
enc = PcdObjEncoder(freeze=False).cuda().train()
pc = synthetic.make_batch(1, 3, 128, seed=4).to("cuda")
emb = enc.encode(pc.obj_fts); emb.sum().backward(); torch.cuda.synchronize()
print("unfrozen encoder ok", float(emb.abs().sum()), float(enc.pcd_net.fc.weight.grad.abs().sum()))
'
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -c "$SCRIPT" > gpurun_out/sanitizer_rows_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|fusion|as_|unfrozen|Error|hazard" gpurun_out/sanitizer_rows_$tool.log | head -14
done

"""Run under torchrun on >= 2 GPUs: the bucketed, overlapped gradient exchange of the prefetch schedule gives the
same parameter trajectory as the sequential schedule with one all-reduce after the backward.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_ddp_equiv.py"""
import datetime
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from msr3d_b200 import synthetic  # noqa: E402
from msr3d_b200.train import HotPathTrainer  # noqa: E402
from tests import helpers  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
B, O, P, seed = 2, 12, 512, 21
batches = [synthetic.make_batch(B, O, P, seed=seed + 10 * rank + i).to(dev) for i in range(4)]   # different data per rank
ups = [torch.randn(B, O, 256, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(4)]
res = {}
for mode in ("sequential", "prefetch_graph", "prefetch_eager"):
    print(f"[rank {rank}] {mode}", file=sys.stderr, flush=True)
    model = helpers.build_prompter(seed, device=dev)          # eval mode: no dropout
    pipe = False if mode == "sequential" else "prefetch"
    tr = HotPathTrainer(model, B, O, P, lr=1e-4, world_size=world, use_graph=mode != "prefetch_eager", pipeline=pipe)
    if mode == "sequential":
        for b, u in zip(batches, ups):
            tr.step(b, u)
    else:
        tr.engine.prime2(batches[0], ups[0], batches[1], ups[1])
        if tr.use_graph:
            tr.engine.capture_prefetch()
        tr._calls = 2
        for b, u in zip(batches[2:], ups[2:]):
            tr.step(b, u)
        tr.flush()
    torch.cuda.synchronize()
    print(f"[rank {rank}] {mode} stepped", file=sys.stderr, flush=True)
    res[mode] = tr.engine.flat_p.clone()
    # replicas stay identical
    ref = res[mode].clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, res[mode]), f"{mode}: replicas diverged on rank {rank}"
# gradient accumulation over 2 micro-steps with the gated peer-memory exchange: two updates from four batches, replicas identical
model = helpers.build_prompter(seed, device=dev)
tr = HotPathTrainer(model, B, O, P, lr=1e-4, world_size=world, use_graph=True, pipeline="prefetch", accum_steps=2)
if tr.engine.p2p is not None:
    tr.engine.prime2(batches[0], ups[0], batches[1], ups[1])
    tr.engine.capture_prefetch()
    tr._calls = 2
    for b, u in zip(batches[2:], ups[2:]):
        tr.step(b, u)
    tr.flush()
    torch.cuda.synchronize()
    assert float(tr.engine.opt_state[1]) == 2.0 and int(tr.engine.accum_state[0]) == 0
    acc = tr.engine.flat_p.clone()
    ref = acc.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, acc), f"accumulation: replicas diverged on rank {rank}"
    assert torch.isfinite(acc).all() and not torch.equal(acc, res["sequential"])
    if rank == 0:
        print("accumulation (k = 2) under the peer-memory exchange: 2 updates, replicas identical")
for mode in ("prefetch_graph", "prefetch_eager"):
    e = helpers.rel_err(res[mode], res["sequential"])
    if rank == 0:
        print(f"{mode} vs sequential: rel err {e:.2e}")
    assert e < 5e-4, (mode, e)
if rank == 0:
    print("ddp equivalence ok", flush=True)
# communicators whose collectives were captured in CUDA graphs: destroy_process_group() was seen to block forever here
# (NCCL 2.28 / torch 2.11); every rank is done and in sync, so leave without the teardown
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
sys.stderr.flush()
os._exit(0)

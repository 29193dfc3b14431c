"""CUPTI timeline of ONE pipelined, graph-replayed training step exactly as bench.py runs it (encoder of batch i+1
beside the transformer chain of batch i):  python tools/step_timeline.py   [env MSR3D_CHAIN_SMS=...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from msr3d_b200 import configs, synthetic  # noqa: E402
from msr3d_b200.train import HotPathTrainer  # noqa: E402

B, O, P = synthetic.CONFIGS["cfg2_60x1024"]
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:       # python -m torch.distributed.run --nproc-per-node N tools/step_timeline.py: rank 0 prints its timeline
    import datetime
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
model = configs.build_prompter(0, device=dev).train()
tr = HotPathTrainer(model, B, O, P, pipeline="prefetch", world_size=world,
                    chain_sms=int(os.environ.get("MSR3D_CHAIN_SMS", "80")))
batches = [synthetic.make_batch(B, O, P, seed=10 + i).to(dev) for i in range(4)]
up = torch.randn(B, O, 256, device=dev)
tr.warmup(batches[0], up)
for i in range(8):
    tr.step(batches[i % 4], up)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(batches[0], up)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
streams = {}
if rank != 0:
    ev = []
print(f"{'start':>8} {'end':>8} {'dur':>7} strm  kernel")
for e in ev:
    st, en = e.time_range.start - t0, e.time_range.end - t0
    sid = streams.setdefault(getattr(e, "device_resource_id", 0), len(streams))
    print(f"{st:8.1f} {en:8.1f} {en - st:7.1f} {sid:4d}  {e.name[:70]}")
if ev:
    print(f"total {max(e.time_range.end for e in ev) - t0:.1f} us")
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)

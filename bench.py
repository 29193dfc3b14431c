#!/usr/bin/env python
"""bench.py — MSR3D hot-path train-step throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2_60x1024]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one pass of the hot path over one per-GPU batch of synthetic scenes
(BASELINE.json configs[1]: 60 objects x 1024 points, bs=4 per GPU): OSE3DSituation forward
(frozen PointNet++ on the tensor cores + spatial transformer), backward through the trainable
part from a fixed upstream gradient (what the frozen LLM returns), gradient all-reduce over
NCCL when N > 1, global-norm clip and AdamW.  The LLM itself is library code outside the path.

Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM; `e2e` = the same
step driven from pinned HOST buffers (H2D of the batch and D2H of the loss inside the timed region).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MSR3D hot-path train-step samples/sec (60 obj x 1024 pt, bs=4/GPU; prompter fwd+bwd+allreduce+AdamW, frozen LLM excluded)"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2_60x1024")
    ap.add_argument("--split", type=int, default=6,
                    help="tensor-core operand split of the transformer GEMMs: 6 = three-way bf16 (fp32-grade, default), "
                         "3 = two-way, 1 = single bf16 pass; the frozen encoder uses the two-way split unless 1")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="run the frozen encoder and the transformer step of a batch back to back instead of "
                         "overlapping the encoder of batch i+1 with the step of batch i")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="copy the point clouds on the compute stream in front of the step instead of prefetching "
                         "them on a copy stream during the previous step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def load_spec():
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")) as fh:
        return [tuple(x) for x in json.load(fh)]


def prompter_cfg():
    from tests.helpers import prompter_cfg as cfg
    return cfg()


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the same step on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from msr3d_b200 import synthetic
    from oracle import ext_oracle, train_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ext_oracle.set_threads(cores)
    B, O, P = synthetic.CONFIGS[args.config]
    sd = synthetic.make_state_dict(load_spec(), 0)
    hot = train_oracle.CpuHotPath(sd)
    batch = synthetic.make_batch(B, O, P, seed=2000)
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(1))
    steps = max(1, min(args.steps, 10))
    warm = max(1, min(args.warmup, 3))
    for _ in range(warm):
        hot.step(batch.as_dict(), up)
    t0 = time.perf_counter()
    for _ in range(steps):
        hot.step(batch.as_dict(), up)
    dt = (time.perf_counter() - t0) / steps
    val = B / dt
    sample = f"{steps} full steps of {args.config} (B={B}) after {warm} warm-up, oracle port (torch CPU + C/OpenMP native ops)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.config, "batch_per_gpu": B, "objects": O, "points": P,
                   "note": "CPU arm runs on rank 0 only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(args):
    from msr3d_b200 import synthetic
    from oracle import ext_oracle, train_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ext_oracle.set_threads(cores)
    B, O, P = synthetic.CONFIGS[args.config]
    hot = train_oracle.CpuHotPath(synthetic.make_state_dict(load_spec(), 0))
    batch = synthetic.make_batch(B, O, P, seed=2000)
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(1))
    hot.step(batch.as_dict(), up)
    n = 3
    t0 = time.perf_counter()
    for _ in range(n):
        hot.step(batch.as_dict(), up)
    dt = (time.perf_counter() - t0) / n
    return {"value": B / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} full steps of {args.config} (B={B}) after 1 warm-up; oracle port: torch CPU fp32 + C/OpenMP pointnet2 ops"}


def bench_splice(dev, B, L, flush, cores):
    """SURVEY.md §8f-1 (the step after the path): llm_proj + fp16 scatter into the LLM input and its
    backward, at config-2 shapes (T = 204 prompt positions, D = 4096).  Per-call CUDA-event times with
    L2 flushed, and the oracle (torch CPU) timed beside it."""
    from msr3d_b200 import splice
    from oracle import modules_oracle as mo
    T, D = 204, 4096
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(0, 30000, (B, T), generator=g)
    ids[:, 100:100 + L] = splice.SCENE_SP_TOKEN
    am = torch.ones(B, T, dtype=torch.long)
    tok = torch.randn(B, L, 256, generator=g)
    masks = torch.ones(B, L, dtype=torch.bool)
    W = torch.randn(D, 256, generator=g) * 0.02
    bias = torch.zeros(D)
    emb = (torch.randn(B, T, D, generator=g) * 0.02).half()
    G = torch.randn(B, T, D, generator=g).half()
    d = lambda t: t.to(dev)  # noqa: E731
    ids_d, am_d, masks_d, G_d = d(ids), d(am), d(masks), d(G)
    Wd, bd, tokd = (d(t).requires_grad_(True) for t in (W, bias, tok))
    res = {}
    for mode in ("fwd", "fwd+bwd"):
        ts = []
        for i in range(13):
            e = d(emb)
            for t in (Wd, bd, tokd):
                t.grad = None
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out, _ = splice.splice_scene_tokens(e, am_d, ids_d, tokd, masks_d, Wd, bd, check=False)
            if mode != "fwd":
                out.backward(G_d)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b) * 1e3)
        res[mode + "_us"] = round(statistics.median(ts), 1)
    # algorithmic bytes of the forward: tokens + weight + bias read, B*L rows of D halves written
    fwd_bytes = B * L * 256 * 4 + D * 256 * 4 + D * 4 + B * L * D * 2
    res["fwd_algorithmic_bytes"] = fwd_bytes
    res["fwd_gbs"] = round(fwd_bytes / (res["fwd_us"] * 1e-6) / 1e9, 1)
    res["note"] = ("eager launches through the autograd op (3 kernels forward, 4 backward), host launch gaps "
                   "included; the GEMM in a graph chain is ~10 us (tools/gemm_sweep.py)")
    torch.set_num_threads(cores)
    Wc, bc, tc = (t.clone().requires_grad_(True) for t in (W, bias, tok))
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        o, _ = mo.scene_splice(Wc, bc, emb.clone(), am, ids, tc, masks)
        o.backward(G)
    res["cpu_oracle_fwd+bwd_us"] = round((time.perf_counter() - t0) / n * 1e6, 1)
    res["workload"] = f"B={B} T={T} L={L} D={D} fp16 embeddings"
    return res


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    import msr3d_b200
    from msr3d_b200 import _C, ops, synthetic
    from msr3d_b200.train import HotPathTrainer

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 (NCCL prints its
    # version banner there when the first communicator is created) are sent to stderr until the line is printed
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    ops.DEFAULT_SPLIT = 3 if args.split == 6 else args.split
    B, O, P = synthetic.CONFIGS[args.config]
    model = msr3d_b200.build_model(prompter_cfg())
    sd = synthetic.make_state_dict(load_spec(), 0)          # identical replicas on every rank
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).train()
    trainer = HotPathTrainer(model, B, O, P, world_size=world, split=args.split, use_graph=not args.no_graph,
                             pipeline=False if args.no_pipeline else (True if args.no_prefetch else "prefetch"))
    # four distinct pinned host batches per rank, cycled (seed = cfg_id*1000 + rank)
    host = [synthetic.make_batch(B, O, P, seed=2000 + rank + 17 * i).pin() for i in range(4)]
    devb = [h.to(dev) for h in host]
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(1)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    h2d = host[0].nbytes()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, from_host):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        for i in range(n_steps):
            flush.zero_()                                    # L2 flush between timed iterations
            ev[i][0].record()
            if from_host:
                trainer.step(host[i % 4], up, loss_out=loss_host)   # H2D from pinned memory, loss read back
            else:
                trainer.step(devb[i % 4], up)
            ev[i][1].record()
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in ev]

    trainer.warmup(devb[0], up)
    for i in range(max(args.warmup, 3)):
        trainer.step(devb[i % 4], up)
    timed(3, True)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    from msr3d_b200 import roofline
    barrier()
    ms = timed(args.steps, False)
    barrier()
    launches_per_step = trainer.engine.launches_per_step
    ms_e2e = timed(args.steps, True)
    barrier()
    sampler.stop_flag = True
    sampler.join()

    # per-kernel durations (every rank: the step contains a collective): the same step launched eagerly
    # with CUDA events around every kernel (a graph replay cannot be instrumented from the host).  A
    # spin kernel keeps the stream busy while the host enqueues the whole step, so the events bracket
    # back-to-back kernels and not host launch gaps; L2 is flushed before each step.
    ops.TIMER = ktimer = roofline.KernelTimer()
    for i in range(min(args.steps, 20)):
        trainer.engine.load_inputs(devb[i % 4], up)
        flush.zero_()
        torch.cuda._sleep(40_000_000)                         # ~20 ms of GPU spin
        trainer.engine.train_step_eager()
        torch.cuda.synchronize()
    ops.TIMER = None
    barrier()

    tot = torch.tensor([sum(ms), sum(ms_e2e)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)           # max over ranks
    t_dev, t_e2e = (tot / 1e3).tolist()
    if rank == 0:
        value = B * world * args.steps / t_dev
        e2e = B * world * args.steps / t_e2e
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (split-bf16 tensor-core MMAs, f32 accumulate in TMEM)" if args.split != 1 else "bf16",
            "data": "synthetic",
            "config": {"workload": args.config, "batch_per_gpu": B, "objects": O, "points": P,
                       "l2": "flushed (256 MB memset) between timed iterations", "parallelism": f"dp{world}",
                       "schedule": "sequential" if args.no_pipeline else
                       ("pipelined: frozen encoder of batch i+1 overlaps the transformer step of batch i "
                        "(one encoder pass + one update per step)" +
                        ("" if args.no_prefetch else "; point clouds of batch i+2 prefetched (H2D on a copy stream) "
                         "during the same step")),
                       "timing": "sum of per-step CUDA-event intervals, max over ranks"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": sampler.result(),
        }
        roof, table = roofline.summarize(ktimer, B, O, P, 3 if args.split == 6 else args.split,
                                         steps=min(args.steps, 20))
        line["roofline"] = roof
        line["kernels"] = {k: {kk: (round(vv, 5) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                           for k, v in table.items()}
        # the two side measurements must never cost the bench line itself
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args)
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}
        if world == 1:
            try:
                line["next_rows"] = {"scene_splice (SURVEY 8f-1)": bench_splice(dev, B, O, flush, os.cpu_count() or 1)}
            except Exception as e:  # noqa: BLE001
                line["next_rows"] = {"error": f"{type(e).__name__}: {e}"}
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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

L2_TEXT = "flushed (256 MB memset) between timed iterations"
TIMING_TEXT = "sum of per-step CUDA-event intervals, max over ranks"
SCHEDULE_TEXT = {True: "sequential",
                 False: "pipelined: frozen encoder of batch i+1 overlaps the transformer step of batch i (one encoder pass + "
                        "one update per step); point clouds of batch i+2 prefetched (H2D on a copy stream) during the same step"}
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
    ap.add_argument("--fast", action="store_true",
                    help="headline numbers only: skip the side measurements (CPU / GPU baselines, secondary configs, "
                         "LLM context line, scene-token splice)")
    ap.add_argument("--chain-sms", type=int, default=int(os.environ.get("MSR3D_CHAIN_SMS", "80")),
                    help="SMs of the partition that runs the transformer chain (the frozen encoder of the next batch "
                         "gets the rest); 0 = no partition")
    return ap.parse_args()


def load_spec():
    from msr3d_b200 import configs
    return configs.load_spec()


def prompter_cfg():
    from msr3d_b200 import configs
    return configs.prompter_cfg()


def config_dict(name, B, O, P, world, schedule, l2, timing):
    """The `config` object: identical keys in both arms (the driver compares them)."""
    return {"workload": name, "batch_per_gpu": B, "objects": O, "points": P, "l2": l2, "parallelism": f"dp{world}",
            "schedule": schedule, "timing": timing}


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
    steps = max(1, args.steps)
    warm = max(1, args.warmup)
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
        "config": config_dict(args.config, B, O, P, world, SCHEDULE_TEXT[False], L2_TEXT, TIMING_TEXT),
        "note": "CPU arm: ONE process on rank 0 stepping B=%d samples per step on all %d host cores (whatever --gpus says): "
                "compare with the N=1 line of the GPU arm only" % (B, cores),
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


def bench_unfrozen_encoder(dev, B, O, P, flush):
    """forward + backward of the UNFROZEN PointNet++ encoder (vision.args.freeze False, training-mode
    BatchNorm) at the bench configuration's clouds — this repo's kernel path (msr3d_b200/encoder_train.py) and, beside
    it, the layer-wise torch module tree (cuDNN convolutions / ATen BatchNorm over the same index kernels)."""
    from msr3d_b200 import synthetic
    from msr3d_b200.encoder import PcdObjEncoder
    batch = synthetic.make_batch(B, O, P, seed=2100).to(dev)
    G = torch.randn(B, O, 768, device=dev, generator=torch.Generator(dev).manual_seed(3))
    res = {"workload": f"{B * O} clouds x {P} points, training-mode BatchNorm, fwd + bwd, inputs resident, L2 flushed"}
    tf32_was = torch.backends.cudnn.allow_tf32
    for key, kernels, tf32 in (("kernels_ms", True, tf32_was), ("torch_module_tree_tf32_ms", False, True),
                               ("torch_module_tree_fp32_ms", False, False)):
        torch.backends.cudnn.allow_tf32 = tf32
        enc = PcdObjEncoder(freeze=False)
        spec = [(k, list(v.shape), "int64" if v.dtype == torch.int64 else "float32") for k, v in enc.state_dict().items()]
        enc.load_state_dict(synthetic.make_state_dict(spec, 5))
        enc = enc.to(dev).train()
        enc.pcd_net.train_kernels = kernels
        ts = []
        for i in range(7):
            for p in enc.parameters():
                p.grad = None
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            emb = enc.encode(batch.obj_fts)
            emb.backward(G)
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b))
        res[key] = round(statistics.median(ts), 3)
        del enc
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = tf32_was
    # algorithmic work: 113.5 MFLOP per cloud forward, data + weight gradients twice that
    gflop = 3 * 113.5e-3 * B * O
    res["algorithmic_gflop"] = round(gflop, 1)
    res["kernels_tflops"] = round(gflop / res["kernels_ms"], 2)
    res["note"] = ("fp32-grade 3xTF32 products through the TMA-fed tcgen05 GEMM of the transformer (csrc/gemm3x.cu), "
                   "BatchNorm statistics / apply / pooling as separate HBM passes (about 8 GB per step): not yet fused into "
                   "the GEMM prologues and epilogues; module tree = cuDNN convolutions + ATen BatchNorm over this repo's "
                   "index kernels, with TF32 allowed (torch default) and in fp32")
    return res


def bench_module_call(dev, B, O, P, flush):
    """The drop-in call a trainer makes (north star: same nn.Module signature): OSE3DSituation.forward(data_dict) +
    loss.backward() through msr3d_b200/autograd.py, per micro-batch, inputs resident — with the two programs replayed as
    CUDA graphs (default) and launched eagerly through ctypes."""
    import msr3d_b200
    from msr3d_b200 import synthetic
    model = msr3d_b200.build_model(prompter_cfg())
    model.load_state_dict(synthetic.make_state_dict(load_spec(), 0), strict=True)
    model = model.to(dev).train()
    batch = synthetic.make_batch(B, O, P, seed=2200).to(dev)
    up = torch.randn(B, O, 256, device=dev, generator=torch.Generator(dev).manual_seed(4))
    res = {"workload": f"model(data_dict)['obj_tokens'] + backward, {B} x {O} x {P}, training mode, L2 flushed"}
    for key, flag in (("graphs_ms", "1"), ("eager_ms", "0")):
        os.environ["MSR3D_MODULE_GRAPH"] = flag
        ts = []
        for i in range(13):
            for p in model.parameters():
                p.grad = None
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = model(batch.as_dict())["obj_tokens"]
            out.backward(up)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        res[key] = round(statistics.median(ts), 4)
    os.environ["MSR3D_MODULE_GRAPH"] = "1"
    res["samples_per_s_graphs"] = round(B / (res["graphs_ms"] * 1e-3), 1)
    return res


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
                             pipeline=False if args.no_pipeline else (True if args.no_prefetch else "prefetch"),
                             chain_sms=0 if args.no_pipeline else args.chain_sms)
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

    def timed(tr, n_steps, from_host, hostb, devbb, upp):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        for i in range(n_steps):
            flush.zero_()                                    # L2 flush between timed iterations
            ev[i][0].record()
            if from_host:
                tr.step(hostb[i % 4], upp, loss_out=loss_host)   # H2D from pinned memory, loss read back
            else:
                tr.step(devbb[i % 4], upp)
            ev[i][1].record()
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in ev]

    trainer.warmup(devb[0], up)
    for i in range(max(args.warmup, 3)):
        trainer.step(devb[i % 4], up)
    timed(trainer, 3, True, host, devb, up)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    from msr3d_b200 import roofline
    barrier()
    ms = timed(trainer, args.steps, False, host, devb, up)
    barrier()
    launches_per_step = trainer.engine.launches_per_step
    ms_e2e = timed(trainer, args.steps, True, host, devb, up)
    barrier()
    sampler.stop_flag = True
    sampler.join()

    # per-kernel durations: CUPTI intervals over replays of the whole step as one sequential CUDA graph captured with
    # programmatic dependent launch OFF (msr3d_b200/roofline.py explains why); L2 flushed before each replay
    eng = trainer.engine
    trainer.flush()
    lib = _C.lib()
    lib.msr3d_set_pdl(0)
    eng.load_inputs(devb[0], up)
    eng.graph = None
    eng.capture()
    times = roofline.cupti_kernel_times(eng.graph.replay, 5, before_each=flush.zero_)
    lib.msr3d_set_pdl(1)
    eng.graph = None
    barrier()

    tot = torch.tensor([sum(ms), sum(ms_e2e)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)           # max over ranks
    t_dev, t_e2e = (tot / 1e3).tolist()
    if rank == 0:
        value = B * world * args.steps / t_dev
        e2e = B * world * args.steps / t_e2e
        ms_step = t_dev / args.steps * 1e3
        part = getattr(eng, "partition", None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (split-bf16 / 3xTF32 tensor-core products, f32 accumulate in TMEM)" if args.split != 1 else "bf16",
            "data": "synthetic",
            "config": config_dict(args.config, B, O, P, world, SCHEDULE_TEXT[bool(args.no_pipeline)], L2_TEXT, TIMING_TEXT),
            "sm_partition": ({"chain_sms": part.counts[0], "encoder_sms": part.counts[1]} if part is not None else None),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": sampler.result(),
        }
        roof, table, step_roof = roofline.summarize_cupti(times, B, O, P, ms_step, n_opt=eng.n_opt)
        line["roofline"] = roof
        line["step_roofline"] = step_roof
        line["kernels"] = {k: {kk: (round(vv, 5) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                           for k, v in sorted(table.items(), key=lambda kv: -kv[1]["us_per_step"])}
        side = world == 1 and not args.fast
        # side measurements must never cost the bench line itself
        if side and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args)
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}
        if side:
            for key, fn in (("gpu_baseline", lambda: gpu_baseline(args, dev, flush)),
                            ("secondary", lambda: secondary_configs(args, dev, world, flush)),
                            ("llm_context", lambda: llm_context(dev, B, O, ms_step)),
                            ("next_rows", lambda: {"scene_splice (SURVEY 8f-1)": bench_splice(dev, B, O, flush, os.cpu_count() or 1),
                                                   "unfrozen_encoder (SURVEY 8f-2)": bench_unfrozen_encoder(dev, B, O, P, flush),
                                                   "module_call (drop-in nn.Module + autograd)": bench_module_call(dev, B, O, P, flush)})):
                try:
                    line[key] = fn()
                except Exception as e:  # noqa: BLE001
                    line[key] = {"error": f"{type(e).__name__}: {e}"}
            if isinstance(line.get("gpu_baseline"), dict) and line["gpu_baseline"].get("value"):
                line["gpu_baseline"]["ratio_value_over_gpu_baseline"] = round(value / line["gpu_baseline"]["value"], 2)
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL collectives were captured in CUDA graphs: tearing the communicator down can block (seen with
        # tools/check_ddp_equiv.py); all ranks are in sync and the line is out, so leave directly
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def gpu_baseline(args, dev, flush):
    """BASELINE.md §4 bar (i): the reference's own pointnet2 kernels recompiled for sm_100 (oracle/_ref) + stock eager
    torch ops, same step definition, on this GPU.  Test infrastructure timed beside the product, like cpu_baseline."""
    from msr3d_b200 import synthetic
    from oracle import train_oracle
    B, O, P = synthetic.CONFIGS[args.config]
    hot = train_oracle.CpuHotPath(synthetic.make_state_dict(load_spec(), 0), device="cuda")
    batch = synthetic.make_batch(B, O, P, seed=2000).to(dev)
    data = batch.as_dict()
    up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(1)).to(dev)
    for _ in range(3):
        hot.step(data, up, sync=False)
    torch.cuda.synchronize()
    n = 10
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.zero_()
        a.record()
        hot.step(data, up, sync=False)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    med = ms[len(ms) // 2]
    return {"value": B / (med * 1e-3), "unit": UNIT, "ms_per_step": med, "kind": "reference pointnet2 .cu for sm_100 + eager torch ops (fp32, TF32 off)",
            "sample": f"median of {n} eager steps of {args.config} (B={B}) after 3 warm-ups, inputs resident, L2 flushed, host launch gaps included"}


def secondary_configs(args, dev, world, flush):
    """Throughput of the other BASELINE.json configurations that fit one GPU (parity cases in tests/test_configs_gpu.py):
    config 4 shapes (80 objects x 2048 points, bs 2: the attention block falls back to the four-launch path, L > 64) and
    config 5 shapes (bs 8, 60 x 1024).  Extra keys; the headline stays config 2."""
    import msr3d_b200
    from msr3d_b200 import synthetic
    from msr3d_b200.train import HotPathTrainer
    out = {}
    for name in ("cfg4_80x2048", "cfg5_msnn_60x1024", "cfg2_60x1024+all_trainable_13.67M"):
        B, O, P = synthetic.CONFIGS[name.split("+")[0]]
        model = msr3d_b200.build_model(prompter_cfg())
        model.load_state_dict(synthetic.make_state_dict(load_spec(), 0), strict=True)
        model = model.to(dev).train()
        extra = None
        if "+" in name:
            # SURVEY.md 8e / BASELINE config 3: the reference's single gradient exchange + clip + AdamW covers the prompter AND
            # llm_proj (1 052 672), llm_proj_img (4 198 400), clip_proj (4 198 400): attached to the engine's flat buffers
            # (their forward / backward belong to the LLM side: zero gradients here, the exchange / optimiser cost is real)
            g = torch.Generator().manual_seed(11)
            shapes = {"llm_proj.weight": (4096, 256), "llm_proj.bias": (4096,), "llm_proj_img.weight": (4096, 1024),
                      "llm_proj_img.bias": (4096,), "clip_proj.weight": (4096, 1024), "clip_proj.bias": (4096,)}
            extra = {k: torch.nn.Parameter((torch.randn(*v, generator=g) * 0.02).to(dev)) for k, v in shapes.items()}
        tr = HotPathTrainer(model, B, O, P, world_size=1, split=args.split, pipeline="prefetch", chain_sms=args.chain_sms,
                            extra_params=extra)
        devb = [synthetic.make_batch(B, O, P, seed=3000 + 17 * i).to(dev) for i in range(4)]
        up = torch.randn(B, O, 256, generator=torch.Generator().manual_seed(1)).to(dev)
        tr.warmup(devb[0], up)
        for i in range(5):
            tr.step(devb[i % 4], up)
        n = 20
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for i, (a, b) in enumerate(ev):
            flush.zero_()
            a.record()
            tr.step(devb[i % 4], up)
            b.record()
        torch.cuda.synchronize()
        t = sum(a.elapsed_time(b) for a, b in ev) / n
        from msr3d_b200 import roofline
        work = roofline.function_work(B, O, P, n_opt=tr.engine.n_opt)
        flops = sum(w["flops"] for _, w in work if "flops" in w)
        peaks = roofline.measured_peaks()
        out[name] = {"value": B / (t * 1e-3), "unit": UNIT, "ms_per_step": t, "batch_per_gpu": B, "objects": O, "points": P,
                     "optimised_parameters": int(tr.engine.n_opt),
                     "attn_block_kernel": bool(getattr(tr.engine, "attn_block_active", False)),
                     "step_frac_tensor": flops / (t * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]}
        del tr, model
        torch.cuda.empty_cache()
    return out


def llm_context(dev, B, L, hot_ms):
    """SURVEY.md §8(d) context: the frozen LLM around the path — HF Llama, Vicuna-7B geometry (32 layers, hidden 4096,
    32 heads, FFN 11008), random fp16 weights, forward + backward to the inputs_embeds of one micro-batch (B prompts of
    128 text tokens + L scene tokens).  Library code outside the path: timed only to put the hot path's share of a
    training micro-step on record."""
    import time as _t
    from transformers import LlamaConfig, LlamaForCausalLM
    T = 128 + L
    cfg = LlamaConfig(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                      num_attention_heads=32, num_key_value_heads=32, max_position_embeddings=2048)
    t0 = _t.perf_counter()
    with torch.device(dev):
        llm = LlamaForCausalLM(cfg).to(torch.float16)
    for p_ in llm.parameters():
        p_.requires_grad_(False)
    llm.gradient_checkpointing_enable()
    llm.train()
    emb = (torch.randn(B, T, 4096, device=dev) * 0.02).half().requires_grad_(True)
    labels = torch.randint(0, 32000, (B, T), device=dev)

    def step():
        emb.grad = None
        out = llm(inputs_embeds=emb, labels=labels, use_cache=False)
        out.loss.backward()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    n = 3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        step()
    b.record()
    torch.cuda.synchronize()
    llm_ms = a.elapsed_time(b) / n
    del llm
    torch.cuda.empty_cache()
    return {"llm": "HF LlamaForCausalLM, Vicuna-7B geometry, random fp16 weights, frozen, gradient checkpointing, eager",
            "tokens_per_prompt": T, "batch": B, "llm_fwd_bwd_ms": llm_ms, "hot_path_ms": hot_ms,
            "train_microstep_samples_per_s": B / ((llm_ms + hot_ms) * 1e-3),
            "hot_path_share_of_microstep": hot_ms / (llm_ms + hot_ms), "setup_s": round(_t.perf_counter() - t0, 1)}


if __name__ == "__main__":
    main()

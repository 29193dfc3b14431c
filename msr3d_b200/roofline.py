"""Per-kernel live timing (CUDA events on the launching stream) and the roofline arithmetic of
DESIGN.md §Kernels.  bench.py switches the timer on for the timed region; ops.* wrap their
launches in `timer.range(name)` when a timer is installed."""
from __future__ import annotations

import json
import os
from contextlib import contextmanager

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written) or the profiling guide's stated fallback.  Any surprise in the
    file (missing keys, other spellings, unreadable) falls back instead of failing the bench."""
    fallback = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            p = json.load(fh)

        def pick(*names):
            for n in names:
                v = p.get(n)
                if isinstance(v, (int, float)) and v > 0:
                    return float(v)
            return None
        hbm = pick("hbm_gbs", "hbm_gb_s", "hbm_GBps", "hbm_bandwidth_gbs")
        tf = pick("bf16_tflops", "bf16_tf_s", "bf16_dense_tflops")
        if hbm is None or tf is None:
            return fallback
        return {"hbm_gbs": hbm, "bf16_tflops": tf,
                "bf16_tflops_sustained": pick("bf16_tflops_sustained", "bf16_sustained_tflops") or tf, "source": "measured"}
    except (OSError, ValueError, AttributeError):
        return fallback


class KernelTimer:
    def __init__(self):
        self.events = {}

    @contextmanager
    def range(self, name):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        try:
            yield
        finally:
            b.record()
            self.events.setdefault(name, []).append((a, b))

    def mean_ms(self):
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in self.events.items()}

    def counts(self):
        return {k: len(v) for k, v in self.events.items()}


def sa_level_flops(rows: int, widths) -> float:
    """2*M*K*N over the layers of one SharedMLP applied to `rows` (centre,sample) positions."""
    return float(sum(2.0 * rows * widths[i] * widths[i + 1] for i in range(len(widths) - 1)))


def algorithmic_work(B: int, O: int, P: int):
    """Algorithmic FLOPs / bytes per launch of each encoder kernel for (B,O,P) scenes
    (25.95 / 67.50 / 18.90 MFLOP per cloud for SA1 / SA2 / SA3, 1.18 for fc)."""
    Bc = B * O
    return {
        "sa1_mlp": {"flops": Bc * sa_level_flops(32 * 32, (6, 64, 64, 128)),
                    "bytes": Bc * (P * 24 + 32 * 32 * 4 + 32 * 128 * 4)},
        "sa2_mlp": {"flops": Bc * sa_level_flops(16 * 32, (131, 128, 128, 256)),
                    "bytes": Bc * (32 * 128 * 4 + 16 * 32 * 4 + 16 * 256 * 4)},
        "sa3_l1": {"flops": Bc * sa_level_flops(16, (259, 256))}, "sa3_l2": {"flops": Bc * sa_level_flops(16, (256, 512))},
        "sa3_l3": {"flops": Bc * sa_level_flops(16, (512, 768))}, "fc": {"flops": Bc * 2.0 * 768 * 768},
        # spatial transformer GEMMs (per launch; M = B*O tokens)
        "qkvc": {"flops": 2.0 * B * O * 256 * 816}, "out_proj": {"flops": 2.0 * B * O * 256 * 256},
        "ffn1": {"flops": 2.0 * B * O * 256 * 2048}, "ffn2": {"flops": 2.0 * B * O * 256 * 2048},
        "obj_proj": {"flops": 2.0 * B * O * 768 * 256},
        "dW2": {"flops": 2.0 * B * O * 256 * 2048}, "du": {"flops": 2.0 * B * O * 256 * 2048},
        "dW1": {"flops": 2.0 * B * O * 256 * 2048}, "dy2": {"flops": 2.0 * B * O * 256 * 2048},
        "dWo": {"flops": 2.0 * B * O * 256 * 256}, "dctx": {"flops": 2.0 * B * O * 256 * 256},
        "dWqkvc": {"flops": 2.0 * B * O * 256 * 816}, "dx": {"flops": 2.0 * B * O * 256 * 816},
        "dWproj": {"flops": 2.0 * B * O * 768 * 256},
        # index kernels: compulsory traffic = xyz read + index/centre writes
        "sa1_group": {"bytes": Bc * (P * 12 + 32 * 4 + 32 * 12 + 32 * 32 * 4)},
        "sa2_group": {"bytes": Bc * (32 * 12 + 16 * 4 + 16 * 12 + 16 * 32 * 4)},
    }


#: bench tag -> kernel function
_GEMM_TAGS = ("sa3_l1", "sa3_l2", "sa3_l3", "fc", "loc_fc", "obj_proj", "qkvc", "out_proj", "ffn1", "ffn2", "dW2",
              "du", "dW1", "dy2", "dWo", "dctx", "dWqkvc", "dx", "dWproj", "dWloc")
KERNEL_OF_TAG = {t: "linear_tc_kernel" for t in _GEMM_TAGS}
KERNEL_OF_TAG.update({"sa1_mlp": "sa_mlp_kernel<16,64,64,128>", "sa2_mlp": "sa_mlp_kernel<144,128,128,256>"})
#: dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures (cold cache:
#: ncu flushes L2 between replays).  sa_mlp: profiles/r1_ncu_sa_mlp.md.  linear_tc: mean over 26
#: consecutive launches of the transformer step, profiles/r1_ncu_linear_tc_raw.csv (0.8 - 4.4 MB per
#: launch = the fp32 weight matrix + the activations + the zero-filled split-K output lines).
TRAFFIC_NCU = {"sa_mlp_kernel<16,64,64,128>": 5.93e6, "sa_mlp_kernel<144,128,128,256>": 4.47e6,
               "linear_tc_kernel": 2.73e6}


def summarize(timer: KernelTimer, B: int, O: int, P: int, split: int, in_long_step: bool = True, steps: int = 1):
    """roofline object for the dominant kernel + a per-kernel table."""
    ms = timer.mean_ms()
    work = algorithmic_work(B, O, P)
    peaks = measured_peaks()
    tf_peak = peaks["bf16_tflops_sustained"] if in_long_step else peaks["bf16_tflops"]
    table = {}
    for name, t in ms.items():
        w = work.get(name, {})
        row = {"ms": t, "launches": timer.counts()[name]}
        if "flops" in w:
            row["tflops"] = w["flops"] / (t * 1e-3) / 1e12
            row["frac_tensor"] = row["tflops"] / tf_peak
        if "bytes" in w:
            row["gbs"] = w["bytes"] / (t * 1e-3) / 1e9
            row["frac_hbm"] = row["gbs"] / peaks["hbm_gbs"]
        table[name] = row
    # the dominant kernel = the kernel FUNCTION with the largest total time in the step, aggregated over
    # its launches (all transformer / SA3 / fc GEMMs are launches of linear_tc_kernel)
    groups = {}
    for name, row in table.items():
        fn = KERNEL_OF_TAG.get(name, name)
        g = groups.setdefault(fn, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0, "tags": []})
        per_step = row["launches"] / max(1, steps)
        g["ms"] += row["ms"] * per_step
        g["launches"] += per_step
        g["tags"].append(name)
        w = work.get(name, {})
        g["flops"] += w.get("flops", 0.0) * per_step
        g["bytes"] += w.get("bytes", 0.0) * per_step
    rated = {k: v for k, v in groups.items() if v["flops"] > 0 or v["bytes"] > 0}
    if not rated:
        return {"error": "no kernels timed"}, table
    top = max(rated, key=lambda k: rated[k]["ms"])
    g = rated[top]
    total_ms = sum(v["ms"] for v in groups.values())
    common = {"kernel": top, "launches_per_step": round(g["launches"], 2), "ms_per_step": round(g["ms"], 5),
              "share_of_kernel_time": round(g["ms"] / total_ms, 4), "traffic": TRAFFIC_NCU.get(top)}
    if g["flops"] > 0:
        ach = g["flops"] / (g["ms"] * 1e-3) / 1e12
        roof = {**common, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                "peak_source": f"{peaks['source']} bf16 {'sustained' if in_long_step else 'burst'} (MEASURED_PEAKS.json)",
                "note": "algorithmic fp32-equivalent FLOPs summed over the kernel's launches / summed live CUDA-event "
                        "durations (each bracketed launch carries ~7 us of event overhead); the kernel issues 3-6x "
                        "that in bf16 MMAs (split-bf16). These GEMMs are latency bound (M = 240 tokens: 5.5 us fixed "
                        "+ ~1.2 us per 64-wide K chunk in a graph chain), see profiles/r1_ncu_linear_tc.md and "
                        "profiles/r1_gemm_sweep.md"}
    else:
        ach = g["bytes"] / (g["ms"] * 1e-3) / 1e9
        roof = {**common, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "peak_source": f"{peaks['source']} (MEASURED_PEAKS.json)"}
    return roof, table


# ---------------------------------------------------------------------------------------------------------------
# Round 2: per-kernel durations from CUPTI (torch.profiler) over replays of the whole step as ONE CUDA graph that was
# captured with programmatic dependent launch switched OFF — under PDL a kernel's CUPTI interval starts when its
# first CTA is scheduled (inside the previous kernel), so intervals overlap and sum to more than the step.  With PDL
# off every interval is execution time, the intervals of one stream do not overlap, and the dominant kernel's
# per-step time is necessarily <= the step.  (Round 1 bracketed eager launches with CUDA events: ~7 us of event
# overhead per launch, 0.83 ms of "kernel time" in a 0.65 ms step.)
# ---------------------------------------------------------------------------------------------------------------
import re


def cupti_kernel_times(replay, n_replays: int = 5, before_each=None):
    """{kernel name: (us per replay, launches per replay)} of `replay()` (a CUDA-graph replay)."""
    from torch.profiler import ProfilerActivity, profile
    for _ in range(2):
        replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n_replays):
            if before_each is not None:
                before_each()
            replay()
        torch.cuda.synchronize()
    acc = {}
    t_first, t_last = None, None
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = e.name
        if "FillFunctor<unsigned char>" in name:      # the 256 MB L2 flush in front of each replay
            continue
        if name.startswith("Memcpy") or name.startswith("Memset"):
            name = name.split(" ")[0]
        a = acc.setdefault(name, [0.0, 0])
        a[0] += e.time_range.end - e.time_range.start
        a[1] += 1
    return {k: (v[0] / n_replays, v[1] / n_replays) for k, v in acc.items()}


def _short(name: str) -> str:
    """'void msr3d::sa_mlp_kernel<144, 128, ...>(msr3d::SaParams)' -> 'sa_mlp_kernel<144,128,128,256,false>'"""
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("msr3d::", "").replace(" ", "")
    return name.replace("false", "0").replace("true", "1")        # ncu prints bool template arguments as 0 / 1


def function_work(B: int, O: int, P: int, nl: int = 3, n_opt: int = 4224144, F: int = 2048):
    """Algorithmic work per STEP of each kernel function, keyed by a regex on the short kernel name (SURVEY.md §8d
    per-unit figures x units of config (B, O, P)).  flops are fp32-equivalent (the tensor cores issue
    3x that as split-bf16 / 3xTF32 products)."""
    M = B * O
    w = algorithmic_work(B, O, P)
    attn_mm = 2.0 * B * 8 * O * O * 32 * 2                 # q k^T and P v of one layer
    gemm3x = (w["obj_proj"]["flops"] + w["dWproj"]["flops"]
              + nl * (w["dW2"]["flops"] + w["dW1"]["flops"] + w["dWo"]["flops"] + w["dctx"]["flops"] + w["dWqkvc"]["flops"] + w["dx"]["flops"]))
    return [
        (r"^sa_mlp_kernel<16,", {"flops": w["sa1_mlp"]["flops"], "bytes": w["sa1_mlp"]["bytes"], "what": "SA1 MLP 6-64-64-128 + max-pool, gather fused"}),
        (r"^sa_mlp_kernel<144,", {"flops": w["sa2_mlp"]["flops"], "bytes": w["sa2_mlp"]["bytes"], "what": "SA2 MLP 131-128-128-256 + max-pool, gather fused"}),
        (r"^linear_t(ma|c)_kernel", {"flops": w["sa3_l1"]["flops"] + w["sa3_l2"]["flops"] + w["sa3_l3"]["flops"] + w["fc"]["flops"] + 2 * 2.0 * M * 63 * 256,
                                     "what": "SA3 (GroupAll MLP 259-256-512-768) + fc 768-768 + loc_fc / dWloc"}),
        (r"^gemm3x_kernel", {"flops": gemm3x, "what": "object projection, data / weight gradients of the transformer (3xTF32)"}),
        (r"^ffn_pair_kernel<\d+,0", {"flops": nl * (w["ffn1"]["flops"] + w["ffn2"]["flops"]), "what": "FFN1 + GELU + dropout + FFN2 (forward)"}),
        (r"^ffn_pair_kernel<\d+,1", {"flops": nl * (w["du"]["flops"] + w["dy2"]["flops"]), "what": "du + dy2 (backward of the FFN pair)"}),
        (r"^attn_block_fwd_kernel", {"flops": nl * (w["qkvc"]["flops"] + w["out_proj"]["flops"] + attn_mm), "what": "QKV+cond, attention, out-projection, LayerNorm tail"}),
        (r"^adamw_kernel", {"bytes": 28.0 * n_opt, "what": "clip + AdamW: 16 B read + 12 B written per parameter"}),
        (r"^sumsq_kernel", {"bytes": 4.0 * n_opt, "what": "gradient norm"}),
        (r"^split_tf32_kernel", {"bytes": 12.0 * n_opt, "what": "tf32 big / small images of the parameters"}),
        (r"^fps_kernel<256", {"bytes": w["sa1_group"]["bytes"], "what": "FPS (+ centre gather) of SA1"}),
        (r"^ball_query", {"bytes": w["sa1_group"]["bytes"] + w["sa2_group"]["bytes"], "what": "ball query of SA1 and SA2"}),
    ]


def ncu_traffic():
    """{kernel regex: dram bytes per launch} from the committed ncu capture (profiles/r2_ncu_traffic.json, produced by
    tools/ncu_traffic.py from the raw CSV next to it); {} if absent."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh)
    except (OSError, ValueError):
        return {}


def summarize_cupti(times, B: int, O: int, P: int, ms_per_step: float, nl: int = 3, n_opt: int = 4224144):
    """(roofline object of the dominant kernel function, per-function table, whole-step object).
    `times` = cupti_kernel_times() of the sequential full-step graph (PDL off)."""
    peaks = measured_peaks()
    tf_peak, hbm_peak = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
    work = function_work(B, O, P, nl, n_opt)
    traffic = ncu_traffic()
    table = {}
    for name, (us, n) in times.items():
        sh = _short(name)
        row = table.setdefault(sh, {"us_per_step": 0.0, "launches_per_step": 0.0})
        row["us_per_step"] += us
        row["launches_per_step"] += n
    total_us = sum(r["us_per_step"] for r in table.values())
    for sh, row in table.items():
        for rx, wk in work:
            if re.search(rx, sh):
                row["what"] = wk["what"]
                if "flops" in wk:
                    row["tflops"] = wk["flops"] / (row["us_per_step"] * 1e-6) / 1e12
                    row["frac_tensor"] = row["tflops"] / tf_peak
                if "bytes" in wk:
                    row["gbs"] = wk["bytes"] / (row["us_per_step"] * 1e-6) / 1e9
                    row["frac_hbm"] = row["gbs"] / hbm_peak
                break
        for rx, tb in traffic.items():
            if re.search(rx, sh):
                row["dram_bytes_per_launch_ncu"] = tb
        row["share_of_kernel_time"] = row["us_per_step"] / total_us if total_us else 0.0
    rated = {k: v for k, v in table.items() if "tflops" in v or "gbs" in v}
    if not rated:
        return {"error": "no kernels profiled"}, table, {}
    top = max(rated, key=lambda k: rated[k]["us_per_step"])
    r = rated[top]
    per_launch_us = r["us_per_step"] / max(r["launches_per_step"], 1e-9)
    common = {"kernel": top, "what": r.get("what"), "launches_per_step": round(r["launches_per_step"], 2),
              "us_per_launch": round(per_launch_us, 2), "ms_per_step": round(r["us_per_step"] * 1e-3, 5),
              "share_of_kernel_time": round(r["share_of_kernel_time"], 4), "traffic": r.get("dram_bytes_per_launch_ncu"),
              "time_base": "CUPTI kernel intervals over 5 replays of the whole step as one CUDA graph captured with PDL off "
                           "(intervals = execution time; L2 flushed before each replay)"}
    if "tflops" in r:
        roof = {**common, "bound": "tensor", "achieved": r["tflops"], "peak": tf_peak, "unit": "TFLOP/s", "frac": r["frac_tensor"],
                "peak_source": f"{peaks['source']} bf16 sustained (MEASURED_PEAKS.json)",
                "note": "achieved = algorithmic fp32-equivalent FLOPs per launch / launch duration; the kernel issues 3x that "
                        "as split-bf16 products (fp32-grade results from bf16 tensor cores)"}
    else:
        roof = {**common, "bound": "hbm", "achieved": r["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": r["frac_hbm"],
                "peak_source": f"{peaks['source']} (MEASURED_PEAKS.json)"}
    step_flops = sum(wk["flops"] for _, wk in work if "flops" in wk)
    step = {"algorithmic_gflop": round(step_flops / 1e9, 2), "ms_per_step": ms_per_step,
            "achieved_tflops": step_flops / (ms_per_step * 1e-3) / 1e12,
            "frac_tensor": step_flops / (ms_per_step * 1e-3) / 1e12 / tf_peak,
            "sum_of_kernel_us_sequential_graph": round(total_us, 1)}
    return roof, table, step

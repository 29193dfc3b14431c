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

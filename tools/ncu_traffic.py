"""Per-kernel-function summary of an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv`:
mean DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), duration, tensor-pipe activity.
    python tools/ncu_traffic.py profiles/r2_ncu_raw.csv profiles/r2_ncu_traffic.json profiles/r2_ncu_summary.md
The JSON ({regex on the short kernel name: bytes per launch}) is what bench.py reports as roofline.traffic."""
import csv
import json
import re
import sys

src, out_json, out_md = sys.argv[1:4]
rows = list(csv.reader(open(src)))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = rows[hdr]
units = rows[hdr + 1]
data = [r for r in rows[hdr + 2:] if len(r) == len(cols)]
ci = {c: i for i, c in enumerate(cols)}


def num(r, name):
    if name not in ci:
        return None
    v = r[ci[name]].replace(",", "")
    try:
        x = float(v)
    except ValueError:
        return None
    u = units[ci[name]]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
    return x * scale.get(u, 1.0)


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("msr3d::", "").replace(" ", "")


agg = {}
for r in data:
    k = short(r[ci["Kernel Name"]])
    a = agg.setdefault(k, {"n": 0, "dram": 0.0, "us": 0.0, "tensor": 0.0, "grid": r[ci.get("Grid Size", 0)], "regs": None, "smem": None})
    a["n"] += 1
    rd, wr = num(r, "dram__bytes_read.sum") or 0.0, num(r, "dram__bytes_write.sum") or 0.0
    a["dram"] += rd + wr
    a["us"] += num(r, "gpu__time_duration.sum") or 0.0
    t = num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    if t is None:
        t = num(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")
    a["tensor"] += t or 0.0
    a["regs"] = num(r, "launch__registers_per_thread")
    a["smem"] = num(r, "launch__shared_mem_per_block_dynamic")
traffic = {"^" + re.escape(k): a["dram"] / a["n"] for k, a in agg.items()}
json.dump(traffic, open(out_json, "w"), indent=1, sort_keys=True)
with open(out_md, "w") as fh:
    fh.write("| kernel | launches captured | mean duration (us, cold, serialised) | mean DRAM bytes / launch | tensor pipe active % | regs | dyn smem |\n|---|---:|---:|---:|---:|---:|---:|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        fh.write(f"| `{k}` | {a['n']} | {a['us'] / a['n']:.1f} | {a['dram'] / a['n'] / 1e6:.2f} MB | {a['tensor'] / a['n']:.1f} | {a['regs']} | {a['smem']} |\n")
print(open(out_md).read())

"""Summarise an .ncu-rep (ncu --set full) into profiles/: a markdown table (duration, DRAM bytes and throughput, issue
utilisation, FMA pipe, tensor pipe, registers, top stall reasons per kernel) and profiles/ncu_traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum per launch, consumed by bench.py's roofline.traffic).

    python tools/ncu_summary.py gpurun_out/r2f_prof.ncu-rep r2f
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, default=float("nan")):
    i = col.get(name)
    if i is None or r[i] == "":
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def unit(name):
    i = col.get(name)
    return units[i] if i is not None else ""


def to_bytes(r, name):
    v, u = val(r, name), unit(name).lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return v * mult


def to_us(r, name):
    v, u = val(r, name), unit(name).lower()
    mult = {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
    return v * mult


def short(name):
    name = re.sub(r"void <unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    return name


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
lines = [f"# ncu --set full summary `{os.path.basename(rep)}` ({tag})", "",
         "Per launch (cold caches, serialised under the profiler: compare shares and ratios, not absolute times with the bench).", "",
         "| kernel | us | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % of peak | issue active % | FMA pipe % | tensor pipe % | regs | top stalls (warps per issue) |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for r in data:
    name = short(r[col["Kernel Name"]])
    us = to_us(r, "gpu__time_duration.sum")
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    st = sorted(((val(r, h, 0.0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall_cols), reverse=True)[:4]
    lines.append(f"| `{name}` | {us:.1f} | {rd/1e6:.0f} | {wr/1e6:.0f} | {(rd+wr)/us/1e3:.0f} | {val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
                 f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {val(r, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | "
                 f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0):.1f} | {val(r, 'launch__registers_per_thread'):.0f} | "
                 + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |")
    traffic.setdefault(name, []).append({"us": us, "dram_bytes": rd + wr})
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.md"), "w").write("\n".join(lines) + "\n")
out = {"source": os.path.basename(rep), "tag": tag, "per_kernel": {k: {"launches": len(v), "dram_bytes_per_launch": sum(x["dram_bytes"] for x in v) / len(v),
                                                                        "us_per_launch": sum(x["us"] for x in v) / len(v)} for k, v in traffic.items()}}
json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json"), "w"), indent=1)
print("\n".join(lines))

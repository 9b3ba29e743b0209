#!/usr/bin/env python
"""Turn ncu outputs into the small tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches <launches.csv> <out.md> [regex]
    python scripts/summarize_ncu.py full <report.ncu-rep> <out.md> [traffic_key]

`launches` aggregates a `--metrics gpu__time_duration.sum` launch list per
kernel (count, total, share).  `full` extracts the roofline-relevant counters
of every profiled launch from a `--set full` report and, with a traffic key,
updates profiles/dram_traffic.json (bytes per launch, read by bench.py).
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def launches(path, out, pattern=None):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik])[:90]
        if pattern and not re.search(pattern, r[ik]):
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({os.path.basename(path)})\n\n")
        f.write("gpu__time_duration.sum per launch, --clock-control none; cold-cache serialised times: compare shares.\n\n")
        f.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {c} | {t / 1e3:.1f} | {t / 1e3 / c:.2f} | {100 * t / tot:.1f}% |\n")
    print(open(out).read())


def full(rep, out, key=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({os.path.basename(rep)})\n\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"- `{m}` [{units[i]}]: " + ", ".join(r[i] for r in data) + "\n")
        stalls = collections.Counter()
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                for r in data:
                    if r[i]:
                        stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] += float(r[i].replace(",", ""))
        tot = sum(stalls.values()) or 1.0
        f.write("\nwarp stall samples (all profiled launches): " +
                ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in stalls.most_common(8)) + "\n")
        f.write("\nkernels: " + "; ".join(sorted({re.sub(r'\(.*', '', r[ik]) for r in data})) + "\n")
    print(open(out).read())
    if key:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")

        def to_bytes(v, u):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return float(v.replace(",", "")) * scale

        per = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in data]
        tpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "dram_traffic.json")
        d = json.load(open(tpath)) if os.path.isfile(tpath) else {}
        d[key] = sum(per) / len(per)
        json.dump(d, open(tpath, "w"), indent=1, sort_keys=True)
        print("traffic per launch:", d[key])


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)

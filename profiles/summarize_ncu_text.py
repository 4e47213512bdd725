"""Round-2 profile summary from the artefacts of profiles/r02_capture.sh (brought back in gpurun_out/):

    python profiles/summarize_ncu_text.py r02

  * launch lists (`ncu --metrics gpu__time_duration.sum --csv`) -> per-kernel launches / mean us / share
  * `ncu --set full` details pages (text logs r02_<family>.log) -> one row per captured kernel:
    duration, DRAM / memory / compute throughput %, issue slots, executed instructions, registers, occupancy.
Writes profiles/<tag>/summary.md, copies the launch lists and the bench line, and refreshes profiles/knn_traffic.json.
"""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
FIELDS = [("Duration", "us"), ("DRAM Throughput", "DRAM %"), ("Memory Throughput", "mem %"), ("Compute (SM) Throughput", "SM %"),
          ("Issue Slots Busy", "issue %"), ("Executed Ipc Active", "IPC"), ("Executed Instructions", "warp instr"),
          ("Registers Per Thread", "regs"), ("Achieved Occupancy", "occ %"), ("L1/TEX Hit Rate", "L1 hit %"),
          ("L2 Hit Rate", "L2 hit %")]


def launch_table(path, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault(r[ki].split("(")[0], []).append(v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1))
    tot = sum(sum(v) for v in agg.values())
    md = ["### %s" % title, "", "| kernel | launches | mean us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        md.append("| `%s` | %d | %.1f | %.1f %% |" % (k[:88], len(v), sum(v) / len(v), 100 * sum(v) / tot))
    return md + [""]


def parse_details(path):
    out, cur = [], None
    for line in open(path, errors="replace"):
        m = re.match(r"^  (\S.*?) \((\d+), (\d+), (\d+)\)x\((\d+), (\d+), (\d+)\), Context", line)
        if m:
            cur = {"kernel": m.group(1), "grid": "%sx%sx%s" % m.group(2, 3, 4), "block": m.group(5)}
            out.append(cur)
            continue
        if cur is None:
            continue
        for name, _ in FIELDS:
            m = re.match(r"^\s+%s\s+(\S+)\s+([\d.,]+)\s*$" % re.escape(name), line)
            if m and name not in cur:
                cur[name] = (m.group(2).replace(",", ""), m.group(1))
        m = re.search(r"DRAM.*?(\d[\d.,]*)\s*(K|M|G)?byte", line)
        if "Mem Pipes Busy" in line:
            pass
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    outdir = os.path.join(ROOT, "profiles", tag)
    os.makedirs(outdir, exist_ok=True)
    md = ["# ncu summary %s" % tag, "",
          "Produced by `profiles/r02_capture.sh` on one B200 through gpurun.  Launch lists are cold-cache and serialised"
          " (compare SHARES); the `--set full` rows are one replayed launch each (`--clock-control none`).", ""]
    for f, title in (("r02_launches_bench.csv", "`python bench.py --profile --steps 20` (headline kernel + secondary rows)"),
                     ("r02_launches_dcp.csv", "`python profiles/prof_run.py dcp` (two DCP forwards, B=32 N=1024)")):
        p = os.path.join(GO, f)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(outdir, f))
            md += launch_table(p, title)
    md += ["## `ncu --set full` captures", "",
           "| family | kernel | grid x block | " + " | ".join(h for _, h in FIELDS) + " |", "|---|---|---|" + "---|" * len(FIELDS)]
    for fam in ("knn32", "knnstream", "edge", "emd", "attn", "group", "misc", "rpm", "chamfer"):
        p = os.path.join(GO, "r02_%s.log" % fam)
        if not os.path.exists(p):
            continue
        for d in parse_details(p):
            cells = []
            for name, _ in FIELDS:
                v = d.get(name)
                cells.append("" if v is None else ("%s %s" % v if name in ("Duration", "Executed Instructions") else v[0]))
            md.append("| %s | `%s` | %s x %s | %s |" % (fam, d["kernel"].split("(")[0][:70], d["grid"], d["block"], " | ".join(cells)))
    b = os.path.join(GO, "r02_bench.json")
    if os.path.exists(b):
        shutil.copy(b, os.path.join(outdir, "bench_n1.json"))
    open(os.path.join(outdir, "summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md[-40:]))


if __name__ == "__main__":
    main()

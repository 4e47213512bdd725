"""Turn ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize_ncu.py <round-tag> <launches.csv> <full.ncu-rep> [...more .ncu-rep]

Writes profiles/<round-tag>/launches.csv (copy), profiles/<round-tag>/summary.md and, for the kNN
kernel, profiles/knn_traffic.json (dram bytes per launch, read by bench.py's roofline.traffic).
Needs the `ncu` CLI (no GPU) to read the reports.
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def read_rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = (r[hdr.index(k)], units[hdr.index(k)])
        res.append(d)
    return res


def main():
    tag, launches = sys.argv[1], sys.argv[2]
    reps = sys.argv[3:]
    outdir = os.path.join(ROOT, "profiles", tag)
    os.makedirs(outdir, exist_ok=True)
    shutil.copy(launches, os.path.join(outdir, "launches.csv"))
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1)
        agg.setdefault(r[ki].split("(")[0], []).append(ns)
    total = sum(sum(v) for v in agg.values())
    md = ["# ncu summary %s" % tag, "",
          "Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` over "
          "`python bench.py --steps 50 --warmup 5 --profile` (cold-cache, serialised: compare SHARES).", "",
          "| kernel | launches | mean us | share of GPU time |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        md.append("| `%s` | %d | %.2f | %.1f %% |" % (k[:90], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / total))
    for rep in reps:
        md += ["", "## `--set full` capture: %s" % os.path.basename(rep), ""]
        for d in read_rep(rep):
            md.append("### `%s`" % d["kernel"][:100])
            md.append("")
            md.append("| metric | value |")
            md.append("|---|---|")
            for k in KEYS:
                if k in d:
                    md.append("| %s | %s %s |" % (k, d[k][0], d[k][1]))
            md.append("")
            if "knn_kernel<0, 1" in d["kernel"] or "knn_kernel<(int)0, (int)1" in d["kernel"]:
                traffic = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
                json.dump({"dram_bytes_per_launch": traffic, "source": "profiles/%s/%s" % (tag, os.path.basename(rep)),
                           "kernel": d["kernel"]},
                          open(os.path.join(ROOT, "profiles", "knn_traffic.json"), "w"), indent=1)
    open(os.path.join(outdir, "summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md[:40]))


if __name__ == "__main__":
    main()

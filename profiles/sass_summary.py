"""Blackwell-specific SASS mnemonics per kernel of libl3d_b200.so (no GPU needed):

    python profiles/sass_summary.py r02      # writes profiles/r02/sass_mnemonics.md

UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = tensor-map TMA load, UBLKCP = 1-D bulk async copy,
FFMA2/FMUL2/FADD2 = packed fp32 (B200_PROFILING.md "What proves a Blackwell-native kernel").
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA2", "FMUL2",
         "FADD2", "FMNMX3", "MUFU.EX2", "REDUX", "HMMA", "ACQBULK", "ELECT"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    so = os.path.join(ROOT, "learning3d_b200", "libl3d_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    archs = collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            archs[m.group(1)] += 1
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            per[cur]["_total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    per[cur][w] += 1
                    if w == "UTCHMMA" and ".2CTA" in op:
                        per[cur]["UTCHMMA.2CTA"] += 1
    cols = [w for w in WATCH + ["UTCHMMA.2CTA"] if any(c[w] for c in per.values())]
    md = ["# SASS mnemonics of learning3d_b200/libl3d_b200.so (%s)" % tag, "",
          "`cuobjdump -sass`; cubin architectures: %s.  Columns = static instruction counts per kernel." %
          ", ".join("%s x%d" % kv for kv in archs.items()), "",
          "| kernel | instrs | " + " | ".join(cols) + " |", "|---|---|" + "---|" * len(cols)]
    tot = collections.Counter()
    for k, c in per.items():
        if not any(c[w] for w in cols):
            continue
        md.append("| `%s` | %d | %s |" % (k[:80], c["_total"], " | ".join(str(c[w]) if c[w] else "" for w in cols)))
        tot.update(c)
    md.append("| **all %d kernels** | %d | %s |" % (len(per), sum(c["_total"] for c in per.values()),
                                                   " | ".join(str(sum(c[w] for c in per.values())) for w in cols)))
    os.makedirs(os.path.join(ROOT, "profiles", tag), exist_ok=True)
    path = os.path.join(ROOT, "profiles", tag, "sass_mnemonics.md")
    open(path, "w").write("\n".join(md) + "\n")
    print("\n".join(md[-12:]))


if __name__ == "__main__":
    main()

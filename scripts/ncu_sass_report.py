#!/usr/bin/env python
"""Instruction-cache view of an `ncu --set full --import-source on` capture of a dense kernel: per source function, the
number of SASS instructions, how many of them are hot, executed instructions per QP, share of the stall samples and the
share of `no_inst` (instruction fetch) stalls among them.

    ncu -i X.ncu-rep --page source --csv --print-source sass > X_sass.csv          (here, no GPU needed)
    python scripts/ncu_sass_report.py X_sass.csv <QPs in the launch> <lib.so the capture ran> [kernel mangled-name fragment]

The line information comes from `nvdisasm -g` of the cubin inside the .so (same build as the capture)."""
import bisect
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    sass_csv, nq, lib = sys.argv[1], float(sys.argv[2]), os.path.abspath(sys.argv[3])
    frag = sys.argv[4] if len(sys.argv) > 4 else "gi_dense_cta_kernelILi2ELb0ELb0"
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.startswith("capi")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and frag in l)
    recs, cur = [], None
    for ln in dis[start + 1:]:
        if ln.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+\S.*?;", ln):
            recs.append(cur)
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    assert len(data) == len(recs), (len(data), len(recs), "capture and library are not the same build")
    src = open(os.path.join(ROOT, "jrl-qp_b200", "csrc", "gi_dense_cta.cuh")).read().split("\n")
    marks = [i + 1 for i, l in enumerate(src) if ("__device__" in l or "__global__" in l)]

    def fn(c):
        if c is None:
            return "?"
        f, l = c
        if f != "gi_dense_cta.cuh":
            return f
        j = bisect.bisect_right(marks, l) - 1
        return src[marks[j] - 1].strip()[:60] if j >= 0 else "?"

    g = lambda r, k: float(r[ix[k]] or 0)
    tot_s = sum(g(r, "# Samples") for r in data)
    tot_i = sum(g(r, "Instructions Executed") for r in data) / nq
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    mix = {h[6:]: sum(g(r, h) for r in data) for h in st}
    hot = sum(1 for r in data if g(r, "Instructions Executed") / nq >= 5)
    print("SASS instructions %d (%.0f KB), hot (>= 5 executions per QP) %d (%.0f KB); warp-instructions per QP %.0f" % (len(data), len(data) * 16 / 1024, hot, hot * 16 / 1024, tot_i))
    print("stall mix:", {k: round(100 * v / tot_s, 1) for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:9]})
    agg = {}
    for c, r in zip(recs, data):
        a = agg.setdefault(fn(c), [0, 0, 0.0, 0.0, 0.0])
        ex = g(r, "Instructions Executed") / nq
        a[0] += 1
        a[1] += 1 if ex >= 5 else 0
        a[2] += ex
        a[3] += g(r, "# Samples")
        a[4] += g(r, "stall_no_inst")
    print("%-62s %6s %6s %8s %7s %7s" % ("function (innermost frame of the line info)", "sass", "hot", "instr/QP", "samp%", "noinst%"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][3]):
        if a[3] / tot_s > 0.001:
            print("%-62s %6d %6d %8.0f %7.1f %7.1f" % (k, a[0], a[1], a[2], 100 * a[3] / tot_s, 100 * a[4] / max(a[3], 1)))


if __name__ == "__main__":
    main()

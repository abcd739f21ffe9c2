#!/usr/bin/env python
"""Per-function and per-line summary of an `ncu --set full --import-source on` capture of the dense kernel.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_source.csv
    python scripts/ncu_report.py X_source.csv <QPs in the captured launch> [function-name-fragment ...]

Functions are delimited by the __device__ / __global__ lines of jrl-qp_b200/csrc/gi_dense_cta.cuh (the capture must come
from the same revision of that file); for every fragment given, the hottest lines of that function are listed."""
import bisect
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    nq = float(sys.argv[2])
    frags = sys.argv[3:]
    secs, cur, hdr, ix = [], None, None, {}
    for r in rows:
        if r and r[0] == "Line No":
            hdr, cur = r, {}
            secs.append(cur)
            ix = {h: i for i, h in enumerate(hdr)}
            continue
        if cur is None or len(r) < len(hdr) or not r[0]:
            continue
        try:
            ln = int(r[0])
            st = {h[6:]: float(r[i] or 0) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h}
            cur[ln] = (r[1], float(r[ix["Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0), st)
        except ValueError:
            pass
    main_sec = max(secs, key=len)
    tot = sum(v[1] for s in secs for v in s.values())
    tots = sum(v[2] for s in secs for v in s.values())
    print("total warp-instructions per QP %.0f (per source file: %s)" % (tot / nq, [round(sum(v[1] for v in s.values()) / nq) for s in secs]))
    agg = {}
    for s in secs:
        for v in s.values():
            for k, x in v[3].items():
                agg[k] = agg.get(k, 0) + x
    print("stall mix:", {k: round(100 * v / max(1, sum(agg.values())), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    src = open(os.path.join(ROOT, "jrl-qp_b200", "csrc", "gi_dense_cta.cuh")).read().split("\n")
    marks = [i + 1 for i, l in enumerate(src) if "__device__" in l or "__global__" in l]
    fn = {}
    for ln, v in main_sec.items():
        j = bisect.bisect_right(marks, ln) - 1
        key = (marks[j], src[marks[j] - 1].strip()[:70]) if j >= 0 else (0, "?")
        a = fn.setdefault(key, [0, 0])
        a[0] += v[1]
        a[1] += v[2]
    for k, (i, sm) in sorted(fn.items(), key=lambda kv: -kv[1][1]):
        if i / nq > 300:
            print("%5d inst %6.0f/QP %4.1f%%  samples %4.1f%% | %s" % (k[0], i / nq, 100 * i / tot, 100 * sm / tots, k[1]))
    for frag in frags:
        keys = [k for k in fn if frag in k[1]]
        if not keys:
            continue
        lo = keys[0][0]
        hi = min([m for m in marks if m > lo] + [len(src) + 1])
        print("--- hottest lines of", keys[0][1])
        for ln, v in sorted(((ln, v) for ln, v in main_sec.items() if lo <= ln < hi), key=lambda kv: -kv[1][2])[:14]:
            s = sorted(v[3].items(), key=lambda kv: -kv[1])[:2]
            print("%4d inst %5.0f/QP samples %4.2f%% %-30s | %s" % (ln, v[1] / nq, 100 * v[2] / tots, ",".join("%s=%d" % (a, b) for a, b in s), v[0].strip()[:80]))


if __name__ == "__main__":
    main()

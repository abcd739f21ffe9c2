"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        iI, iS = ix["Instructions Executed"], ix["# Samples"]
        iE = ix["L1 Wavefronts Shared Excessive"]
        stall_ix = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue
    try:
        ln = int(r[0])
        inst = float(r[iI])
        samp = float(r[iS])
    except ValueError:
        continue
    st = {k: float(r[i] or 0) for k, i in stall_ix.items()}
    data.append((ln, r[1].strip()[:80], inst, samp, float(r[iE] or 0), st))
tot = sum(d[2] for d in data)
tots = sum(d[3] for d in data)
print("total warp-inst %.4g, samples %d" % (tot, tots))
agg = {}
for d in data:
    for k, v in d[5].items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", {k: round(100 * v / max(1, sum(agg.values())), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for d in sorted(data, key=lambda d: -d[3])[:top]:
    s = sorted(d[5].items(), key=lambda kv: -kv[1])[:2]
    print("%4d inst %5.1f%% samp %5.1f%% excWf %9.0f %-28s | %s" % (d[0], 100 * d[2] / tot, 100 * d[3] / tots, d[4],
          ",".join("%s=%d" % (k[6:], v) for k, v in s), d[1]))

"""Per-function breakdown (instructions executed, samples) from an ncu cuda,sass source CSV."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
src = open(sys.argv[2]).read().split("\n")
nqp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
hdr = None; data = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr) or not r[0]: continue
    try: data.append((int(r[0]), float(r[ix["Instructions Executed"]]), float(r[ix["# Samples"]])))
    except ValueError: pass
marks = [(i + 1, l.strip()[:70]) for i, l in enumerate(src) if "__device__" in l or "__global__" in l]
marks.append((len(src) + 1, "end"))
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
print("total inst/QP %.0f" % (tot / nqp))
for (a, name), (b, _) in zip(marks, marks[1:]):
    ins = sum(d[1] for d in data if a <= d[0] < b); sm = sum(d[2] for d in data if a <= d[0] < b)
    if ins: print("%4d-%4d inst %5.1f%% (%7.0f/QP) samp %5.1f%% | %s" % (a, b, 100 * ins / tot, ins / nqp, 100 * sm / tots, name))

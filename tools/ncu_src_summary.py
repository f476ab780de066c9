"""Summarise an ncu source-page CSV: totals of the stall reasons and the hottest SASS instructions (development aid).
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_src_summary.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(float(r[col[h]] or 0) for r in data) for h in stalls}
alls = sum(tot.values())
print("total samples %d, instructions executed %d" % (alls, sum(float(r[col["Instructions Executed"]] or 0) for r in data)))
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v > 0:
        print("  %-24s %8d  %5.1f %%" % (h, v, 100 * v / alls))
# instruction mix by opcode
mix = {}
for r in data:
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    e = mix.setdefault(op, [0.0, 0.0])
    e[0] += float(r[col["Instructions Executed"]] or 0)
    e[1] += float(r[col["# Samples"]] or 0)
print("opcode mix (executed warp instructions, samples):")
for op, (n, sm) in sorted(mix.items(), key=lambda kv: -kv[1][0])[:25]:
    print("  %-10s %12d %8d" % (op, n, sm))
print("hottest instructions:")
for r in sorted(data, key=lambda r: -float(r[col["# Samples"]] or 0))[:top]:
    why = sorted(((float(r[col[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print("  %6s  %-70s %s" % (r[col["# Samples"]], r[col["Source"]].strip()[:70], ", ".join("%s %d" % (h[6:], v) for v, h in why if v > 0)))

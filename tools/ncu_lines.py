"""Hottest CUDA source lines of an ncu report (development aid).
usage: ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > cs.csv; python tools/ncu_lines.py cs.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":      # a CUDA source line (SASS rows carry an address)
        d = {}
        for h, v in zip(hdr, r):
            d.setdefault(h, v)
        try:
            s = float(d["# Samples"] or 0)
        except ValueError:
            s = 0.0
        stalls = sorted(((float(d[h] or 0), h[6:]) for h in hdr if h.startswith("stall_") and "Not Issued" not in h and d.get(h)), reverse=True)[:2]
        out.append((s, cur, int(r[0]), r[1].strip()[:100], float(d["Instructions Executed"] or 0), stalls))
tot = sum(o[0] for o in out)
print("total samples %d" % tot)
byfile = {}
for o in out:
    byfile[o[1]] = byfile.get(o[1], 0) + o[0]
print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.items()})
for s, f, l, src, ins, st in sorted(out, key=lambda o: -o[0])[:top]:
    print("%5.2f%% %-14s %4d %9.0f  %-100s %s" % (100 * s / tot, f, l, ins, src, " ".join("%s:%d" % (n, v) for v, n in st if v > 0)))

"""Per-source-line instruction counts / stall samples / shared wavefronts from an ncu report with imported source:
  ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
  python tools/ncu_src_lines.py src.csv [top_n] [launch_index]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
cur, hdr, launch, seen = None, None, -1, set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        if not seen or (r[1] in seen and r[1] == first):
            launch += 1
            seen = set()
            first = r[1]
        seen.add(r[1])
        continue
    if r[0] == "Line No":
        hdr = r
        iI, iS, iW = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
        continue
    if hdr is None or len(r) < len(hdr) - 1 or launch != want:
        continue
    if r[0] != "" and r[2] == "-":
        try:
            e, s, w = int(r[iI] or 0), int(r[iS] or 0), int(r[iW] or 0)
        except ValueError:
            continue
        k = (cur, int(r[0]), r[1].strip()[:90])
        agg[k][0] += e
        agg[k][1] += s
        agg[k][2] += w
        for i, v in enumerate((e, s, w)):
            tot[i] += v
print("total: inst %d samples %d shared wavefronts %d" % tuple(tot))
byf = collections.defaultdict(lambda: [0, 0, 0])
for (f, l, t), v in agg.items():
    for i in range(3):
        byf[f][i] += v[i]
for f, v in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print("%-34s inst %.3f samples %.3f wavefronts %.3f" % (f, v[0] / tot[0], v[1] / max(1, tot[1]), v[2] / max(1, tot[2])))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s %4d inst %9d (%.3f) samp %6d (%.3f) wf %8d | %s"
          % (k[0], k[1], v[0], v[0] / tot[0], v[1], v[1] / max(1, tot[1]), v[2], k[2]))

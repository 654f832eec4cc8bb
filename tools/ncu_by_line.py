"""Join an ncu SASS source page (csv) with nvdisasm line info to get per-source-line instruction
counts and stall samples.  Usage:
  ncu -i rep.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all lib.so ; nvdisasm -g -c fbank_kernel.sm_100a.cubin > dis.txt
  python tools/ncu_by_line.py src.csv dis.txt '<mangled kernel name>' [kernel index]
"""
import collections
import csv
import re
import sys

src_csv, dis_txt, mangled = sys.argv[1:4]
kidx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
# --- disassembly: ordered list of (opcode text, file, line)
lines = open(dis_txt).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + mangled + ":"))
ins = []
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2), cur))
# --- ncu csv
rows = list(csv.reader(open(src_csv)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[his[kidx]]
end = his[kidx + 1] - 1 if kidx + 1 < len(his) else len(rows)
body = [r for r in rows[his[kidx] + 1:end] if len(r) == len(h)]
iE, iS, iSrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
print("sass in dis: %d, in ncu: %d" % (len(ins), len(body)))
agg = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
for k, r in enumerate(body):
    loc = ins[k][1] if k < len(ins) else ("?", 0)
    e, s = int(r[iE] or 0), int(r[iS] or 0)
    agg[loc][0] += e
    agg[loc][1] += s
    tot[0] += e
    tot[1] += s
print("total inst %d samples %d" % tuple(tot))
for loc, (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    print("%-22s:%4d  inst %9d (%.3f)  samples %6d (%.3f)" % (loc[0], loc[1], e, e / tot[0], s, s / max(tot[1], 1)))

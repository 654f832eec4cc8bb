"""Per-phase / per-line / per-opcode view of an ncu SASS source page joined with nvdisasm line info."""
import csv, re, collections, sys
src_csv, dis_txt, mangled, srcfile = sys.argv[1:5]
marks = [tuple(x.split(':')) for x in sys.argv[5:]]  # name:line boundaries
marks = [(int(l), n) for n, l in marks]
lines = open(dis_txt).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + mangled + ":"))
ins=[]; cur=("?",0)
for l in lines[start+1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append((m.group(2), cur))
rows=list(csv.reader(open(src_csv)))
his=[i for i,r in enumerate(rows) if r and r[0]=="Address"]
h=rows[his[0]]; body=[r for r in rows[his[0]+1:(his[1]-1 if len(his)>1 else None)] if len(r)==len(h)]
iE,iS,iSrc=h.index("Instructions Executed"),h.index("# Samples"),h.index("Source")
assert len(ins)==len(body), (len(ins),len(body))
agg=collections.defaultdict(lambda:[0,0]); tot=[0,0]; opagg=collections.defaultdict(lambda:[0,0])
for k,r in enumerate(body):
    loc=ins[k][1]; e=int(r[iE] or 0); sm=int(r[iS] or 0)
    agg[loc][0]+=e; agg[loc][1]+=sm; tot[0]+=e; tot[1]+=sm
    t=r[iSrc].split(); op=t[1] if t[0].startswith('@') else t[0]
    opagg[op.split('.')[0]][0]+=e; opagg[op.split('.')[0]][1]+=sm
print("total inst %d samples %d"%tuple(tot))
base=srcfile.split('/')[-1]
def phase(loc):
    f,l=loc
    if f!=base: return f
    name='?'
    for a,n in sorted(marks):
        if l>=a: name=n
    return name
ph=collections.defaultdict(lambda:[0,0])
for loc,(e,sm) in agg.items():
    ph[phase(loc)][0]+=e; ph[phase(loc)][1]+=sm
for k,(e,sm) in sorted(ph.items(), key=lambda kv:-kv[1][1]):
    print("%-28s inst %9d (%.3f)  samples %5d (%.3f)"%(k,e,e/tot[0],sm,sm/max(tot[1],1)))
print()
src=open(srcfile).read().splitlines()
for loc,(e,sm) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:18]:
    txt = src[loc[1]-1].strip()[:80] if loc[0]==base and loc[1]<=len(src) else ''
    print("%-20s:%4d inst %8d samples %5d | %s"%(loc[0],loc[1],e,sm,txt))

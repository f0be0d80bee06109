"""Aggregate the warp-stall samples of an ncu source page per source line.
usage: ncu -i x.ncu-rep --page source --csv --print-source cuda,sass --launch-skip N --launch-count 1 > k.csv; python tools/ncu_source_lines.py k.csv lfa_cl_bwd.cu"""
import csv,sys,collections
def toi(x):
    try: return int(x)
    except Exception: return 0
path=sys.argv[1]
rows=list(csv.reader(open(path)))
cur=None; hdr=None
# attribute each SASS instruction to the (file,line); additionally we want "inlined-into" context: unknown. Just per file:line.
agg=collections.defaultdict(lambda: collections.Counter()); src={}
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0] in("Function Name",): continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    if not r[0].isdigit(): continue
    d=dict(zip(hdr,r))
    key=(cur,int(r[0])); src[key]=r[1]
    a=agg[key]; a['samples']+=toi(d['# Samples']); a['inst']+=toi(d['Instructions Executed'])
    for k in hdr:
        if k.startswith('stall_') and 'Not' not in k: a[k]+=toi(d[k])
tot=sum(a['samples'] for a in agg.values()); toti=sum(a['inst'] for a in agg.values())
print('total',tot,toti)
byfile=collections.Counter()
for (f,l),a in agg.items(): byfile[f]+=a['samples']
print({k:round(v/tot,3) for k,v in byfile.most_common()})
want=sys.argv[2]
for key in sorted(k for k in agg if k[0]==want):
    a=agg[key]
    if a['samples']<0.002*tot: continue
    tops=', '.join(f"{k[6:]}={v/max(a['samples'],1):.2f}" for k,v in a.most_common(5) if k.startswith('stall') and v>0.12*a['samples'])
    print(f"{key[1]:4d} {a['samples']/tot:6.2%} inst {a['inst']/toti:6.2%} | {tops} | {src[key].strip()[:80]}")

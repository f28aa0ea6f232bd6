"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line."""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
topn=int(sys.argv[2]) if len(sys.argv)>2 else 30
hi=[i for i,r in enumerate(rows) if r and r[0]=='Line No'][0]
hdr=rows[hi]
idx={}
for i,h in enumerate(hdr):
    idx.setdefault(h,i)
def num(v):
    try: return int(float(v))
    except: return 0
samples=collections.Counter(); insts=collections.Counter(); tinsts=collections.Counter(); text={}
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
stalls=collections.defaultdict(collections.Counter)
tot=0
for r in rows[hi+1:]:
    if len(r)<len(hdr): continue
    try: ln=int(r[0])
    except: continue
    s=num(r[idx['# Samples']]); ie=num(r[idx['Instructions Executed']]); te=num(r[idx['Thread Instructions Executed']])
    samples[ln]+=s; insts[ln]+=ie; tinsts[ln]+=te; text[ln]=r[1][:100]; tot+=s
    for c in stall_cols:
        v=num(r[idx[c]])
        if v: stalls[ln][c]+=v
ti=sum(insts.values())
print('total samples',tot,'total warp insts',ti,'thread insts',sum(tinsts.values()),'avg util',sum(tinsts.values())/max(ti,1))
allst=collections.Counter()
for ln in stalls:
    allst.update(stalls[ln])
print('stall mix:',', '.join(f"{k[6:]}:{100*v/max(tot,1):.1f}%" for k,v in allst.most_common(10)))
for ln,s in samples.most_common(topn):
    top=', '.join(f"{k[6:]}:{v}" for k,v in stalls[ln].most_common(3))
    print(f"{ln:4d} {100*s/max(tot,1):5.1f}% inst={100*insts[ln]/max(ti,1):5.1f}% util={tinsts[ln]/max(insts[ln],1):4.1f} [{top}] {text[ln]}")

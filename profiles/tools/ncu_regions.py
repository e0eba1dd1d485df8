import csv, sys, collections
fn=sys.argv[1]
rows=csv.reader(open(fn)); hdr=None; cur=None; sec="?"
agg=collections.defaultdict(lambda:[0,0])
for r in rows:
    if not r: continue
    if r[0] in ("File Name","File Path"): sec=r[1].split('/')[-1]; continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None: continue
    if r[0]!="":
        if r[0].isdigit(): cur=(sec,int(r[0]))
        continue
    if cur is None or len(r)<len(hdr) or r[2] in ("","..."): continue
    try: inst=int(r[hdr.index("Instructions Executed")]); smp=int(r[hdr.index("# Samples")])
    except ValueError: continue
    agg[cur][0]+=inst; agg[cur][1]+=smp
# need mapping from line->function region: read source files
import re
def region(sec,line):
    return sec,line
ti=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
# group by file and coarse line buckets given on the command line: file:lo-hi:name
groups=[g.split(':') for g in sys.argv[2:]]
out=collections.defaultdict(lambda:[0,0])
for (sec,line),v in agg.items():
    name=f"{sec} (other)"
    for f,rng,nm in groups:
        lo,hi=map(int,rng.split('-'))
        if f in sec and lo<=line<=hi: name=nm; break
    out[name][0]+=v[0]; out[name][1]+=v[1]
for k,v in sorted(out.items(), key=lambda kv:-kv[1][1]): print(f"{k:40s} inst {100*v[0]/ti:5.1f}%  samples {100*v[1]/ts:5.1f}%")

import csv, sys, collections
fn=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(fn)))
# sections start with a "File Name" row
sec="?"; cur=None; kcount=0
agg=collections.defaultdict(lambda:[0,0,0,""])  # (file,line)->[inst,samples,smemwf,src]
hdr=None
for r in rows:
    if not r: continue
    if r[0]=="Kernel Name":
        kcount+=1
        if kcount>1: break
        continue
    if r[0] in ("File Name","File Path"): sec=r[1].split('/')[-1]; continue
    if r[0]=="Line No": hdr=r; continue
    if r[0] in ("Kernel Name","File Path") or not r[0].isdigit() and r[0]!="": continue
    if hdr is None: continue
    if r[0]!="":
        cur=(sec,int(r[0])); agg[cur][3]=r[1]
        continue
    if cur is None or r[2] in ("","..."): continue
    try:
        inst=int(r[hdr.index("Instructions Executed")]); smp=int(r[hdr.index("# Samples")]); wf=int(r[hdr.index("L1 Wavefronts Shared")])
    except ValueError: continue
    agg[cur][0]+=inst; agg[cur][1]+=smp; agg[cur][2]+=wf
tot=sum(v[0] for v in agg.values()); tots=sum(v[1] for v in agg.values()); totw=sum(v[2] for v in agg.values())
print("total inst",tot,"samples",tots,"smem wavefronts",totw)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:top]:
    print(f"{k[0][:18]:18s}:{k[1]:4d} inst {100*v[0]/tot:5.1f}% samp {100*v[1]/tots:5.1f}% smemwf {100*v[2]/max(totw,1):5.1f}%  {v[3].strip()[:90]}")

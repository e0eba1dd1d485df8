"""Per-function share of executed instructions, stall samples and shared-memory wavefronts from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (source lines are attributed to the enclosing
top-level function of the CUDA headers)."""
import csv, sys, collections, re, os
fn = sys.argv[1]
src_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "..", "av-simulation-at-intersections_b200", "csrc")
rows = list(csv.reader(open(fn))); sec = "?"; cur = None; hdr = None; kc = 0
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        kc += 1
        if kc > 1: break
        continue
    if r[0] in ("File Name", "File Path"): sec = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "":
        if r[0].isdigit(): cur = (sec, int(r[0]))
        continue
    if cur is None or r[2] in ("", "..."): continue
    try:
        inst = int(r[hdr.index("Instructions Executed")]); smp = int(r[hdr.index("# Samples")]); wf = int(r[hdr.index("L1 Wavefronts Shared")])
    except ValueError: continue
    agg[cur][0] += inst; agg[cur][1] += smp; agg[cur][2] += wf
files = {}
out = collections.defaultdict(lambda: [0, 0, 0])
for (sec, line), v in agg.items():
    if sec not in files:
        try: files[sec] = open(os.path.join(src_dir, sec)).read().split('\n')
        except Exception: files[sec] = None
    name = sec + " (other)"
    if files[sec]:
        for l in range(min(line, len(files[sec])) - 1, -1, -1):
            t = files[sec][l]
            m = re.match(r'^(?:__device__|__global__|template|inline|static).*?\b(\w+)\s*\(', t)
            if m and not t.startswith(' '): name = m.group(1); break
    for k in range(3): out[name][k] += v[k]
ti = sum(v[0] for v in out.values()); ts = sum(v[1] for v in out.values()); tw = sum(v[2] for v in out.values())
print(f"total thread-level inst {ti} samples {ts} smem wavefronts {tw}")
for k, v in sorted(out.items(), key=lambda kv: -kv[1][1]):
    if v[1] * 200 < ts: continue
    print(f"{k:28s} inst {100*v[0]/ti:5.1f}%  samples {100*v[1]/ts:5.1f}%  smem wavefronts {100*v[2]/max(tw,1):5.1f}%")

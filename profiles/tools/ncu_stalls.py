"""Where one stall reason is sampled: SASS instructions of `ncu --page source --csv --print-source cuda,sass` ranked by
a stall column (default stall_no_inst), with the instruction in front of each (a stall is charged to the instruction
that could not issue, so the one before it is usually the cause).  python ncu_stalls.py src.csv [column] [N]"""
import csv
import sys

fn = sys.argv[1]
col = sys.argv[2] if len(sys.argv) > 2 else "stall_no_inst"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(fn)))
hdr = None
sass = []          # (address, text, samples of col, all samples, executed)
seen_kernel = 0
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1:
            break
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] != "" or len(r) < len(hdr) or r[2] in ("", "..."):
        continue
    try:
        sass.append((r[2], r[3], int(r[hdr.index(col)] or 0), int(r[hdr.index("# Samples")] or 0),
                     int(r[hdr.index("Instructions Executed")] or 0)))
    except ValueError:
        continue
# the source view lists every SASS line once per source line it belongs to; keep first occurrence by address
uniq = {}
order = []
for s in sass:
    if s[0] not in uniq:
        uniq[s[0]] = s
        order.append(s[0])
order.sort(key=lambda a: int(a, 16) if not a.startswith("0x") else int(a, 16))
total = sum(uniq[a][2] for a in order)
print(f"{col}: {total} samples over {len(order)} instructions")
rank = sorted(range(len(order)), key=lambda i: -uniq[order[i]][2])[:top]
for i in rank:
    a, text, c, allc, ex = uniq[order[i]]
    prev = uniq[order[i - 1]][1] if i > 0 else ""
    print(f"{c:6d} {100.0 * c / max(total, 1):5.1f}%  exec {ex:9d}  {a:>8}  {text[:60]:60s} <- {prev[:50]}")

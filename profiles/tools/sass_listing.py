"""cuobjdump -sass listing of one step kernel of the built library, encodings stripped, with a mnemonic summary in the
header (evidence for: TMA bulk copy UBLKCP, mbarrier SYNCS, no tensor-core instructions).
    python profiles/tools/sass_listing.py [T] [G] > profiles/r2_step_kernel_T20.sass"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = os.path.join(ROOT, "av-simulation-at-intersections_b200", "junction_mpc", "libjmpc.so")
T = sys.argv[1] if len(sys.argv) > 1 else "20"
G = sys.argv[2] if len(sys.argv) > 2 else "32"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name = f"mpc_step_kernelILi{T}ELi{G}E"
ins, inside = [], False
for line in out.split("\n"):
    if "Function :" in line:
        inside = name in line
        continue
    if not inside:
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?;)", line)
    if m:
        ins.append((m.group(1), m.group(2)))
ops = collections.Counter()
for _, t in ins:
    tok = t.split()
    ops[(tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0].rstrip(";")] += 1
want = ["UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "SHFL", "LDL", "STL", "LDG", "STG", "ATOMG",
        "WARPSYNC", "BAR"]
tensor = sum(c for o, c in ops.items() if re.search(r"MMA|UTC", o))
print(f"// profiles/r2_step_kernel_T{T}.sass -- cuobjdump -sass of jmpc::mpc_step_kernel<{T}, {G}> (sm_100a), encodings stripped.")
print("// Built by __graft_entry__.build(): nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17.")
print(f"// {len(ins)} instructions.  grep summary: " + ", ".join(f"{o} {ops.get(o, 0)}" for o in want))
print("// UBLKCP = cp.async.bulk (TMA 1-D bulk copy of the Hessian scratch into shared memory), SYNCS = mbarrier arrive/try_wait;")
print(f"// tensor-core mnemonics (HMMA / DMMA / UTCMMA / UTCHMMA ...): {tensor} -- the per-instance systems are 16x16 .. 50x50 fp64, CUDA-core work.")
print()
for a, t in ins:
    print(f"/*{a}*/  {t}")

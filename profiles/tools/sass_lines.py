"""Static SASS instruction counts per source line for one kernel of a cubin.

    cuobjdump -xelf all libjmpc.so && nvdisasm -g -c jmpc.sm_100a.cubin > all.dis
    python profiles/tools/sass_lines.py all.dis mpc_step_kernelILi20ELi32 [lo_hex hi_hex]

Prints the instruction count per (file, line) inside [lo, hi) and an opcode histogram -- used to see where the
solver loop's instructions come from without a GPU (the dynamic counterpart is ncu's source page, ncu_lines.py)."""
import collections
import re
import sys


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
    inside = False
    cur = ("?", 0)
    per_line = collections.Counter()
    ops = collections.Counter()
    total = 0
    for line in open(path):
        if line.startswith(".text."):
            inside = kernel in line
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        if not (lo <= addr < hi):
            continue
        toks = m.group(2).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[op.split(".")[0]] += 1
        per_line[cur] += 1
        total += 1
    print("instructions", total)
    print("opcodes", ops.most_common(24))
    for (f, l), c in per_line.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 60):
        print(f"{c:5d} {100.0 * c / total:5.1f}%  {f}:{l}")


if __name__ == "__main__":
    main()

"""Static size of every step kernel of a built libjmpc.so and of its solver loop (its largest loop): `python profiles/tools/sass_loops.py path/to/lib.so [--lines N]`.  Needs cuobjdump and nvdisasm (no GPU).
With --lines the per-source-line instruction counts of each solver loop are printed too (sass_lines.py)."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def disassemble(lib):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    return subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout


def kernels(text):
    cur, out = None, collections.OrderedDict()
    src = ("?", 0)
    labels_pending = []
    for line in text.split("\n"):
        if line.startswith(".text."):
            cur = line[6:].rstrip(":")
            out[cur] = {"ins": [], "labels": {}}
            continue
        if cur is None:
            continue
        m = re.match(r"^(\.L_x_\d+):", line)
        if m:
            labels_pending.append(m.group(1))
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            src = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m:
            a = int(m.group(1), 16)
            for l in labels_pending:
                out[cur]["labels"][l] = a
            labels_pending = []
            out[cur]["ins"].append((a, m.group(2), src))
    return out


def main():
    lib = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
    ks = kernels(disassemble(lib))
    for name, k in ks.items():
        if "mpc_step_kernel" not in name:
            continue
        m = re.search(r"mpc_step_kernelILi(\d+)ELi(\d+)", name)
        loops = []
        for a, text, _ in k["ins"]:
            b = re.search(r"BRA.*`\((\.L_x_\d+)\)", text)
            if b and k["labels"].get(b.group(1), 1 << 60) < a:
                loops.append((k["labels"][b.group(1)], a))
        # a real back edge has no RET between target and source (the out-of-line divergence fall-backs of the warp
        # collectives sit behind the function's RET and jump back into it); the interior-point loop is the largest
        rets = [a for a, text, _ in k["ins"] if text.startswith("RET")]
        real = [l for l in loops if not any(l[0] < r < l[1] for r in rets)]
        lo, hi = max(real, key=lambda l: l[1] - l[0]) if real else (0, 0)
        body = [i for i in k["ins"] if lo <= i[0] <= hi]
        ops = collections.Counter()
        for _, text, _ in body:
            t = text.split()
            ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += 1
        print(f"T={m.group(1):>2} G={m.group(2):>2}: kernel {len(k['ins'])} instructions, solver loop {len(body)} "
              f"({len(body) * 16 / 1024:.1f} KB)  " + " ".join(f"{o}:{c}" for o, c in ops.most_common(10)))
        if nlines:
            per = collections.Counter(s for _, _, s in body)
            for (f, l), c in per.most_common(nlines):
                print(f"      {c:5d}  {f}:{l}")


if __name__ == "__main__":
    main()

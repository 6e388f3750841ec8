#!/usr/bin/env python
"""Static view of a kernel's main loop (no GPU): disassemble one function of the built library with
nvdisasm -g, find the longest backward branch (the steady-state loop) and count its instructions per
CUDA source line and per opcode class.
usage: python tools/sass_loop.py <cubin-or-so> <mangled-substring> [--dump]"""
import collections
import re
import subprocess
import sys
import tempfile
import os


def disasm(path, sym):
    if path.endswith(".so"):
        d = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, capture_output=True, check=True)
        outs = []
        for f in sorted(os.listdir(d)):
            if f.endswith(".cubin"):
                outs.append(subprocess.run(["nvdisasm", "-g", os.path.join(d, f)], capture_output=True, text=True).stdout)
        text = "\n".join(outs)
    else:
        text = subprocess.run(["nvdisasm", "-g", path], capture_output=True, text=True).stdout
    lines = text.splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith(".text.") and sym in l]
    assert start, f"{sym} not found"
    s = start[0]
    e = next((i for i in range(s + 1, len(lines)) if lines[i].startswith(".text.") or lines[i].startswith(".section")), len(lines))
    return lines[s:e]


def main():
    path, sym = sys.argv[1], sys.argv[2]
    body = disasm(path, sym)
    ins, labels, cur = [], {}, ("?", 0)
    for l in body:
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"^(\.L_x_\d+):", l)
        if m:
            labels[m.group(1)] = len(ins)
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
    best = None
    for i, (a, s, _) in enumerate(ins):
        m = re.search(r"BRA(\.U)\s+(!?U?P\d+,\s*)?`\((\.L_x_\d+)\)", s)   # uniform backward branch = a warp-uniform loop
        if m and m.group(3) in labels and labels[m.group(3)] < i:
            span = i - labels[m.group(3)]
            if best is None or span > best[0]:
                best = (span, labels[m.group(3)], i)
    span, lo, hi = best
    loop = ins[lo:hi + 1]
    print(f"function has {len(ins)} instructions; main loop = [{ins[lo][0]:#x}, {ins[hi][0]:#x}] with {len(loop)} instructions (static, all paths)")
    by_line = collections.Counter((c[0], c[1]) for _, _, c in loop)
    by_op = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0] for _, s, _ in loop)
    print("per opcode:", ", ".join(f"{k} {v}" for k, v in by_op.most_common(40)))
    print("per source line (>= 3):")
    for (f, ln), n in sorted(by_line.items(), key=lambda x: (x[0][0], x[0][1])):
        if n >= 3:
            print(f"  {f}:{ln:<5d} {n}")
    if "--dump" in sys.argv:
        last = None
        for a, s, c in loop:
            if c != last:
                print(f"   ## {c[0]}:{c[1]}")
                last = c
            print(f"  {a:05x}  {s}")


if __name__ == "__main__":
    main()

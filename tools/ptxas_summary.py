#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output (stdin or file): kernel name, registers, spills, smem."""
import re, subprocess, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
pat = sys.argv[2] if len(sys.argv) > 2 else ""
blocks = re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]
for b in blocks:
    name = b.split("'")[0]
    try:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        dem = name
    dem = dem.replace("vitk::(anonymous namespace)::", "").replace("void ", "")
    dem = re.sub(r"\((CUtensorMap|float|int|void|unsigned|__nv|vitk|long|char|const|uint).*", "", dem)
    if pat and pat not in dem:
        continue
    regs = re.search(r"Used (\d+) registers", b)
    spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
    stack = re.search(r"(\d+) bytes stack frame", b)
    print(f"{dem:70s} regs={regs.group(1) if regs else '?':>3s} stack={stack.group(1) if stack else '?'} "
          f"spill={spill.group(1) if spill else '?'}/{spill.group(2) if spill else '?'}")

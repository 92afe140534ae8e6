#!/usr/bin/env python
"""Static SASS instruction mix of one kernel in the built library (no GPU needed):
    python tools/sass_mix.py corrected_fused_kernelILi11 [lib]
Prints the instruction count per opcode class, so that a change of the arithmetic (packed FFMA2 / FADD2 / FMUL2, integer
overhead) can be judged here before GPU time is spent.  Static counts: loops count once."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "phase-vocoder_b200", "libpv_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ops = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", b):
        ops[m.group(1).split(".")[0]] += 1
    tot = sum(ops.values())
    fp = sum(v for k, v in ops.items() if k in ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2"))
    packed = sum(v for k, v in ops.items() if k in ("FADD2", "FMUL2", "FFMA2"))
    print(f"{name[:110]}\n  total {tot}  fp {fp} (packed {packed})  " +
          "  ".join(f"{k} {v}" for k, v in ops.most_common(24)))

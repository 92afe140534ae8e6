#!/usr/bin/env python
"""Dynamic instruction mix of the profiled kernel in an ncu report (executed warp instructions and stall samples per
opcode), from `ncu -i REPORT --page source --csv`:
    python tools/ncu_opmix.py gpurun_out/r02_corrected.ncu-rep [frames_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
ex, sm, wav = collections.Counter(), collections.Counter(), collections.Counter()
for r in rd:
    src = r["Source"].strip()
    parts = src.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.rstrip(";")
    key = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("LDS", "STS", "LDG", "STG", "MUFU", "BAR") else op.split(".")[0]
    try:
        ex[key] += int(r["Instructions Executed"])
        sm[key] += int(r["# Samples"])
        wav[key] += int(r.get("L1 Wavefronts Shared", "0") or 0)
    except ValueError:
        pass
tot, ts = sum(ex.values()), sum(sm.values())
print(f"total executed {tot:.4g}" + (f" = {tot / frames:.0f} per frame" if frames else "") + f", samples {ts}")
print("| opcode | executed | share | " + ("per frame | " if frames else "") + "stall samples share | smem wavefronts |\n|---|---|---|---|---|" + ("---|" if frames else ""))
for k, v in ex.most_common(40):
    pf = f"{v / frames:.1f} | " if frames else ""
    print(f"| {k} | {v:.4g} | {100 * v / tot:.1f} % | {pf}{100 * sm[k] / ts:.1f} % | {wav[k]:.3g} |")

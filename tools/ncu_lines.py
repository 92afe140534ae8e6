#!/usr/bin/env python
"""Executed warp instructions per CUDA source line of the profiled kernel, every SASS instruction counted ONCE (an
inlined instruction is listed under every line of its inline stack: it goes to the outermost line that is not in a
helper header), split into fp / shared / global / other.  Needs -lineinfo and --import-source on.
    python tools/ncu_lines.py REPORT [frames_per_launch] [top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
HELPERS = ("pv_fft_regs.cuh", "sm_100_rt.hpp", "device_functions.hpp", "sm_100_rt.h")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
path, hdr, cur = None, None, None
owner, info = {}, {}
src_of = {}
for rec in csv.reader(io.StringIO(txt)):
    if not rec:
        continue
    if rec[0] == "File Path":
        path = rec[1].split("/")[-1]
    elif rec[0] == "Line No":
        hdr = rec
        i_ex, i_sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and rec[0].isdigit():
        cur = (path, int(rec[0]))
        src_of[cur] = rec[1].strip()
    elif hdr and rec[0] == "" and rec[2].startswith("0x"):
        a = rec[2]
        try:
            info[a] = (rec[3].strip(), int(rec[i_ex]), int(rec[i_sm]))
        except ValueError:
            continue
        rank = 0 if cur[0] in HELPERS else 1
        if a not in owner or rank > owner[a][0]:
            owner[a] = (rank, cur)


def cls(s):
    p = s.split()
    op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
    if op in ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "MUFU", "FSEL", "FSETP", "FMNMX", "F2I", "I2FP", "F2F", "I2F"):
        return 0
    if op in ("LDS", "STS", "LDSM"):
        return 1
    if op in ("LDG", "STG", "LDGSTS", "UBLKCP", "LDC", "LDCU"):
        return 2
    return 3


agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for a, (s, ex, sm) in info.items():
    r = agg[owner[a][1]]
    r[cls(s)] += ex
    r[4] += sm
tot = sum(sum(r[:4]) for r in agg.values())
ts = sum(r[4] for r in agg.values())
print(f"total {tot:.4g} executed ({tot / frames:.0f} per frame), {ts} samples; per frame: all (fp / shared / global+const / other)")
for k, r in sorted(agg.items(), key=lambda kv: -sum(kv[1][:4]))[:top]:
    a = sum(r[:4])
    print(f"{a / frames:7.1f} {100 * a / tot:5.1f}% ({r[0] / frames:6.1f} /{r[1] / frames:6.1f} /{r[2] / frames:6.1f} /{r[3] / frames:6.1f}) st {100 * r[4] / ts:4.1f}%  {k[0]}:{k[1]}  {src_of.get(k, '')[:90]}")

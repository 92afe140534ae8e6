#!/usr/bin/env python
"""Turns what tools/refresh_profiles.sh left in gpurun_out/ into the tracked files under profiles/ (run here, after
the GPU call): ncu summaries, the launch list, r01_traffic.json (DRAM bytes and warp instructions per launch, which
bench.py reports next to its live timing), SASS listings, bench lines, config tables.  usage: refresh_profiles_post.py [r01]"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
FRAMES = 1184 * 860


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return {h: (rows[2][i], rows[1][i]) for i, h in enumerate(rows[0])}


def num(m, k):
    v, u = m[k]
    return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3}.get(u, 1)


traffic = {}
for mode, kern in (("corrected", "corrected_fused"), ("compat", "compat_fused")):
    rep = os.path.join(G, f"{R}_{mode}.ncu-rep")
    md = os.path.join(P, f"{R}_{mode}_fused_n2048.md")
    with open(md, "w") as f:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, str(FRAMES)], stdout=f, check=True)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hotmap.py"), rep, "512"], capture_output=True, text=True).stdout
    with open(md, "a") as f:
        f.write("\n## Hot/cold layout of the code (tools/ncu_hotmap.py; instruction-cache view)\n\n```\n" + hot + "```\n")
    mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_opmix.py"), rep, str(FRAMES)], capture_output=True, text=True).stdout
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, str(FRAMES), "40"], capture_output=True, text=True).stdout
    with open(md, "a") as f:
        f.write("\n## Executed instructions by opcode (tools/ncu_opmix.py)\n\n" + mix)
        f.write("\n## Executed instructions by source line, every instruction once (tools/ncu_lines.py; top 40)\n\n```\n" + lines + "```\n")
    m = raw_metrics(rep)
    traffic[mode] = {"bytes_per_launch": num(m, "dram__bytes_read.sum") + num(m, "dram__bytes_write.sum"),
                     "frames_per_launch": FRAMES,
                     "warp_instructions_per_frame": num(m, "smsp__inst_executed.sum") / FRAMES,
                     "source": f"profiles/{R}_{mode}_fused_n2048.md (one `ncu --set full` capture; constants, not measured in the bench run)",
                     "workload": "1184 streams x 860 frames, window 2048, hop 512"}
json.dump(traffic, open(os.path.join(P, f"{R}_traffic.json"), "w"), indent=1)

# window 4096 (C5's shape): summaries of the two stream kernels on one batch of 592 streams x 430 frames
F4096 = 592 * 430
for mode in ("compat", "corrected"):
    rep = os.path.join(G, f"{R}_{mode}4096.ncu-rep")
    if os.path.exists(rep):
        with open(os.path.join(P, f"{R}_{mode}_fused_n4096.md"), "w") as f:
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, str(F4096)], stdout=f, check=True)

# launch list
rows = [r for r in csv.reader(open(os.path.join(G, f"{R}_launches.csv"))) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    t = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[iu], 1e-6)
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{R}_launch_list.md"), "w") as f:
    f.write(f"# ncu launch list, round {R[1:].lstrip('0')}\n\ncommand: `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e` under\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare shares).\n"
            f"Raw CSV: `profiles/{R}_launches.csv`.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k[:100]}` | {n} | {t:.3f} | {100*t/tot:.1f} % |\n")
shutil.copy(os.path.join(G, f"{R}_launches.csv"), os.path.join(P, f"{R}_launches.csv"))
for n in ("bench.json", "bench_reference.json", "bench_under_profile_config.json", "fft_bench.md"):
    shutil.copy(os.path.join(G, f"{R}_{n}"), os.path.join(P, f"{R}_{n}"))

# SASS of the two headline kernels
so = os.path.join(ROOT, "phase-vocoder_b200", "libpv_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
# (corrected: ...ELi0E = the normal instantiation, ...ELi1E = processing from the stored analysis)
for name, key in (("corrected_fused_n2048", "corrected_fused_kernelILi11ELi4ELi0E"), ("compat_fused_n2048", "compat_fused_kernelILi11ELi5ELi2"),
                  ("corrected_fused_n4096", "corrected_fused_kernelILi12ELi2ELi0E"), ("compat_fused_n4096", "compat_fused_kernelILi12ELi2ELi3"),
                  ("corrected_stored_n4096", "corrected_fused_kernelILi12ELi2ELi1E")):
    parts = sass.split("\t\tFunction : ")
    body = [p for p in parts if p.startswith("_Z") and key in p.split("\n")[0]]
    open(os.path.join(P, f"{R}_sass_{name}.txt"), "w").write("Function : " + body[0] if body else "not found\n")
for n in ("4096_rates.txt",):
    if os.path.exists(os.path.join(G, f"{R}_{n}")):
        shutil.copy(os.path.join(G, f"{R}_{n}"), os.path.join(P, f"{R}_{n}"))
print(json.dumps(traffic, indent=1))

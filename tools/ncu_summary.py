#!/usr/bin/env python
"""Summarises an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few numbers DESIGN.md and
bench.py quote: duration, DRAM bytes, pipe utilisation, shared-memory wavefronts/conflicts, occupancy,
stall breakdown.  Usage: tools/ncu_summary.py gpurun_out/x.ncu-rep frames_per_launch > profiles/x.md"""
import csv
import io
import subprocess
import sys

rep, frames = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def g(k):
    try:
        return float(m[k][0].replace(",", ""))
    except Exception:
        return float("nan")


name = m.get("Kernel Name", ("?", ""))[0]
dur = g("gpu__time_duration.sum")
unit = m["gpu__time_duration.sum"][1]
dur_ms = dur / 1e6 if unit.startswith("ns") else (dur / 1e3 if unit.startswith("us") else dur)
rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
rd *= scale.get(m["dram__bytes_read.sum"][1], 1)
wr *= scale.get(m["dram__bytes_write.sum"][1], 1)
inst = g("smsp__inst_executed.sum")
wf = g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
bc = g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
print(f"# ncu summary: {name}\n")
print(f"source report: `{rep}` (ncu --set full --clock-control none, 1 launch after warm-up; cold-ish caches, serialised)\n")
print("| metric | value |\n|---|---|")
print(f"| duration | {dur_ms:.3f} ms |")
print(f"| frames per launch (incl. recomputed halo frames) | {frames:.0f} |")
print(f"| dram bytes read / written | {rd/1e9:.3f} GB / {wr/1e9:.3f} GB (sum {((rd+wr)/1e9):.3f} GB = `traffic`) |")
print(f"| achieved DRAM throughput | {(rd+wr)/dur_ms/1e6:.1f} GB/s ({g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} % of ncu peak) |")
print(f"| registers / thread | {g('launch__registers_per_thread'):.0f} |")
print(f"| occupancy limit (regs / smem / warps) blocks | {g('launch__occupancy_limit_registers'):.0f} / {g('launch__occupancy_limit_shared_mem'):.0f} / {g('launch__occupancy_limit_warps'):.0f} |")
print(f"| achieved warps active | {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % |")
print(f"| issue slots active | {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} % |")
print(f"| FMA pipe | {g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.1f} % |")
print(f"| ALU pipe | {g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.1f} % |")
print(f"| LSU pipe | {g('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f} % |")
print(f"| XU (MUFU) pipe | {g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.1f} % |")
print(f"| warp instructions | {inst:.3e} ({inst/frames:.0f} per frame) |")
print(f"| shared-memory wavefronts | {wf:.3e} ({wf/frames:.0f} per frame), bank-conflict excess {bc:.3e} ({100*bc/max(wf,1):.2f} %) |")
print(f"| L1 hit rate | {g('l1tex__t_sector_hit_rate.pct'):.1f} % |")
print(f"| L2 hit rate | {g('lts__t_sector_hit_rate.pct'):.1f} % |")
print("\nWarp stall samples (smsp__pcsamp_warps_issue_stalled_*):\n\n| reason | samples | share |\n|---|---|---|")
st = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: g(h) for h in hdr
      if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
tot = sum(v for v in st.values() if v == v)
for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
    if v > 0:
        print(f"| {k} | {v:.0f} | {100*v/tot:.1f} % |")

"""Hot/cold layout of a kernel from an ncu report with SASS source: run-length encodes the per-instruction
execution counts along the code and prints stall shares per window (to spot instruction-cache problems:
short hot runs separated by cold blocks).  usage: ncu_hotmap.py <rep.ncu-rep> [window]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    win = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    ie, ins, ns = ix["Instructions Executed"], ix["stall_no_inst"], ix["# Samples"]
    ex = [int(r[ie]) for r in data]
    mx = sorted(ex)[int(len(ex) * 0.98)]
    tot_s = sum(int(r[ns]) for r in data)
    tot_n = sum(int(r[ins]) for r in data)
    print(f"{rows[0][1]}\nstatic {len(data)} instr, executed {sum(ex):.3e}, samples {tot_s}, no_inst {tot_n} ({100*tot_n/tot_s:.1f} %)")
    print("hot static instructions (>= 0.5 of full rate):", sum(e >= 0.5 * mx for e in ex),
          " partially hot (0.05..0.5):", sum(0.05 * mx <= e < 0.5 * mx for e in ex))
    prev, start, segs = None, 0, []
    for i, e in enumerate(ex):
        b = round(e / mx, 1)
        if b != prev:
            if prev is not None:
                segs.append((start, i - 1, prev))
            prev, start = b, i
    segs.append((start, len(ex) - 1, prev))
    print("runs (first, last, rate):")
    for s in segs:
        if s[1] - s[0] >= 8:
            print("  ", s, data[s[0]][1].split()[0:2])
    print("window  executed  no_inst  samples  share")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for w in range(0, len(data), win):
        blk = data[w:w + win]
        e = sum(int(r[ie]) for r in blk)
        n = sum(int(r[ins]) for r in blk)
        s = sum(int(r[ns]) for r in blk)
        top = sorted(((sum(int(r[ix[h]]) for r in blk), h) for h in stalls), reverse=True)[:3]
        print(f"{w:6d} {e:10.2e} {n:7d} {s:8d} {100*n/max(1,s):5.0f}%   " + ", ".join(f"{h[6:]} {100*v/max(1,s):.0f}%" for v, h in top))


if __name__ == "__main__":
    main()

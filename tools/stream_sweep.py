#!/usr/bin/env python
"""Stream-count sweep at the headline shape (window 2048, hop 512, 860 frames per stream): device-resident
throughput for batches that are NOT a multiple of the resident segment capacity.
Prints a markdown table (kept in profiles/r02_stream_sweep.md).  Usage on the GPU box: python tools/stream_sweep.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import pvb200

N, H, F = 2048, 512, 860
COUNTS = [int(a) for a in sys.argv[1:]] or [300, 592, 600, 740, 900, 1184, 1200, 1300, 1480, 1500, 1776, 1800, 2000, 2368, 2400]


def timed(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    n_in = N + (F - 1) * H
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    smax = max(COUNTS)
    x = torch.randn((smax, n_in), device="cuda", generator=g) * 0.1
    out = torch.empty((smax, 1, F * H), device="cuda")
    print("| mode | streams | ms/launch | frames/s | relative to 1184 streams |")
    print("|---|---|---|---|---|")
    for mode, wt, name in ((pvb200.MODE_CORRECTED, pvb200.WIN_HANN_PERIODIC, "corrected (+7 st)"),
                           (pvb200.MODE_COMPAT, pvb200.WIN_HAMMING, "compat")):
        pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=mode, window_type=wt,
                                 pitch=(float(np.float32(2 ** (7 / 12))),))
        ref = None
        rows = []
        for S in COUNTS:
            ms = timed(lambda: pv.process(x[:S], F, out=out[:S]))
            fps = S * F / (ms * 1e-3)
            rows.append((S, ms, fps))
            if S == 1184:
                ref = fps
        for S, ms, fps in rows:
            rel = f"{fps / ref:.3f}" if ref else "-"
            print(f"| {name} | {S} | {ms:.3f} | {fps / 1e6:.2f} M | {rel} |", flush=True)
        pv.close()


if __name__ == "__main__":
    main()

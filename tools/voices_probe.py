#!/usr/bin/env python
"""Cost of extra voices: device-resident rate at V = 1, 2, 4 for the C4 shape (window 256, hop 64,
4096 streams) and the headline shape (window 2048, hop 512, 1184 streams).  Usage on the GPU box: python tools/voices_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import torch

import pvb200
from stream_sweep import timed

f32 = lambda b: float(np.float32(b))
BETAS = [1.0, f32(2 ** (4 / 12)), f32(2 ** (7 / 12)), 2.0]
print("| window | streams x frames | voices | ms/launch | frames/s | ns per frame and GPU | cost relative to 1 voice |")
print("|---|---|---|---|---|---|---|")
for N, H, S, F in ((256, 64, 4096, 3445), (2048, 512, 1184, 430)):
    x = torch.randn((S, N + (F - 1) * H), device="cuda") * 0.1
    base = None
    for V in (1, 2, 4):
        pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_CORRECTED, window_type=pvb200.WIN_HANN_PERIODIC, pitch=tuple(BETAS[:V]))
        out = torch.empty((S, V, F * H), device="cuda")
        ms = timed(lambda: pv.process(x, F, out=out))
        ns = ms * 1e6 / (S * F)
        base = base or ns
        print(f"| {N} | {S} x {F} | {V} | {ms:.3f} | {S * F / ms / 1e3:.1f} M | {ns:.2f} | {ns / base:.2f} |", flush=True)
        pv.close()
        del out

"""Frame-range split with the analysis stored ({|X|, D} per frame kept by the analysis pass, PvAggArgs::md) against the same split
recomputing the forward transform (PV_NO_MD_STORE=1 at pv_create): ms per call by window, streams and voices."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "phase-vocoder_b200"), ROOT]
import numpy as np
import torch

import pvb200

PITCH = [1.0, 2 ** (4 / 12), 2 ** (7 / 12), 2.0]


def run(N, H, S, F, V, no_md):
    os.environ["PV_MD_MIN_WINDOW"] = "256"          # the table covers the windows the library leaves out by default
    if no_md:
        os.environ["PV_NO_MD_STORE"] = "1"
    else:
        os.environ.pop("PV_NO_MD_STORE", None)
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_CORRECTED, window_type=pvb200.WIN_HANN_PERIODIC,
                             pitch=tuple(float(np.float32(b)) for b in PITCH[:V]))
    x = torch.randn((S, N + (F - 1) * H), device="cuda") * 0.1
    out = torch.empty((S, V, F * H), device="cuda")
    for _ in range(2):
        pv.process(x, F, out=out)
    torch.cuda.synchronize()
    n0 = pv.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        pv.process(x, F, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 4, (pv.launch_count() - n0) // 4


print("| window | streams x frames | voices | stored analysis, ms (launches) | recomputing, ms (launches) |\n|---|---|---|---|---|")
for N, H, S, F in ((512, 128, 400, 1700), (512, 128, 2, 10000), (256, 64, 544, 3445), (256, 64, 2, 6890), (1024, 256, 64, 2000), (2048, 512, 308, 860), (2048, 512, 2, 20000), (4096, 1024, 2, 168750)):
    for V in (1, 2, 4):
        if N == 4096 and V > 1:
            continue
        a, la = run(N, H, S, F, V, False)
        b, lb = run(N, H, S, F, V, True)
        print(f"| {N} | {S} x {F} | {V} | {a:.3f} ({la}) | {b:.3f} ({lb}) |", flush=True)

"""C5 (one hour of 48 kHz stereo, window 4096, hop 1024, corrected +7 semitones) on one GPU: ms per call, and the same with the
stored analysis switched off (PV_NO_MD_STORE=1 at pv_create) -- the two must agree bit for bit.  python tools/c5_once.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests"), ROOT]
import numpy as np
import torch

import pvb200

N, H, F = 4096, 1024, 168750
x = torch.randn((2, N + (F - 1) * H), device="cuda") * 0.1


def run(no_md):
    if no_md:
        os.environ["PV_NO_MD_STORE"] = "1"
    else:
        os.environ.pop("PV_NO_MD_STORE", None)
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_CORRECTED, window_type=pvb200.WIN_HANN_PERIODIC,
                             pitch=(float(np.float32(2 ** (7 / 12))),))
    out = torch.empty((2, 1, F * H), device="cuda")
    st = torch.zeros((2, pv.state_bytes()), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        pv.process(x, F, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        pv.process(x, F, out=out)
    e1.record()
    torch.cuda.synchronize()
    pv.process(x, F, out=out, state=st, flags=pvb200.CARRY_OUT)
    torch.cuda.synchronize()
    print(f"{'recompute' if no_md else 'stored analysis'}: {e0.elapsed_time(e1) / 5:.3f} ms per call, {2 * F / (e0.elapsed_time(e1) / 5) / 1e3:.2f} M frames/s")
    return out, st


a, sa = run(False)
b, sb = run(True)
print("bit-identical output:", bool(torch.equal(a, b)), " state:", bool(torch.equal(sa, sb)))

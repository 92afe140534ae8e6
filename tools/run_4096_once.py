"""One batch of window-4096 streams per mode (profiling target for the in-place kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests"), ROOT]
import numpy as np, torch, pvb200
S, F, N, H = 600, 430, 4096, 1024
x = torch.randn((S, N + (F - 1) * H), device="cuda") * 0.1
for mode, wt in ((pvb200.MODE_COMPAT, pvb200.WIN_HAMMING), (pvb200.MODE_CORRECTED, pvb200.WIN_HANN_PERIODIC)):
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=mode, window_type=wt, pitch=(float(np.float32(2 ** (7 / 12))),))
    out = torch.empty((S, 1, F * H), device="cuda")
    for _ in range(3):
        pv.process(x, F, out=out)
    torch.cuda.synchronize()

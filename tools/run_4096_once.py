"""One batch of window-4096 streams per mode (profiling target for the window-4096 stream kernels; prints the rate)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests"), ROOT]
import numpy as np, torch, pvb200
S, F, N, H = 592, 430, 4096, 1024      # 592 = 2 waves of 2 CTAs x 148 SMs
x = torch.randn((S, N + (F - 1) * H), device="cuda") * 0.1
for mode, wt in ((pvb200.MODE_COMPAT, pvb200.WIN_HAMMING), (pvb200.MODE_CORRECTED, pvb200.WIN_HANN_PERIODIC)):
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=mode, window_type=wt, pitch=(float(np.float32(2 ** (7 / 12))),))
    out = torch.empty((S, 1, F * H), device="cuda")
    for _ in range(3):
        pv.process(x, F, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        pv.process(x, F, out=out)
    e1.record()
    torch.cuda.synchronize()
    print("mode", mode, "window 4096:", round(S * F * 3 / (e0.elapsed_time(e1) * 1e-3) / 1e6, 2), "M frames/s,", pv.launch_count(), "launches")

"""C4's shape (4096 streams, window 256, hop 64, four voices) a few times: profiling target for the window-256 corrected kernel
(two launches of two voices each per call); prints the rate."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests"), ROOT]
import numpy as np, torch, pvb200
S, F, N, H = 3552, 1720, 256, 64       # 3552 = one wave of 3 CTAs x 148 SMs x 8 groups at two voices per launch
x = torch.randn((S, N + (F - 1) * H), device="cuda") * 0.1
pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_CORRECTED, window_type=pvb200.WIN_HANN_PERIODIC,
                         pitch=tuple(float(np.float32(b)) for b in (1.0, 2 ** (4 / 12), 2 ** (7 / 12), 2.0)))
out = torch.empty((S, 4, F * H), device="cuda")
for _ in range(2):
    pv.process(x, F, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    pv.process(x, F, out=out)
e1.record()
torch.cuda.synchronize()
print("C4 shape:", round(S * F * 3 / (e0.elapsed_time(e1) * 1e-3) / 1e6, 2), "M frames/s,", pv.launch_count() // 5, "launches per call")

import os, sys
sys.path[:0] = ["/root/repo/phase-vocoder_b200", "/root/repo/tests", "/root/repo"]
import numpy as np, torch, pvb200
N, H, F = 4096, 1024, 168750
x = torch.randn((2, N + (F - 1) * H), device="cuda") * 0.1
pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_CORRECTED, window_type=pvb200.WIN_HANN_PERIODIC, pitch=(float(np.float32(2 ** (7 / 12))),))
out = torch.empty((2, 1, F * H), device="cuda")
for _ in range(3):
    pv.process(x, F, out=out)
torch.cuda.synchronize()

"""One batched FFT launch shape, a few times (profiling target): python tools/fft_once.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "phase-vocoder_b200")]
import torch

import pvb200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
batch = (1 << 28) // (8 * n)
pv = pvb200.PhaseVocoder(256)
x = (torch.randn(batch, n, device="cuda") + 1j * torch.randn(batch, n, device="cuda")).to(torch.complex64)
o = torch.empty_like(x)
for _ in range(3):
    pv.fft_batch(x, out=o)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    pv.fft_batch(x, out=o)
e1.record()
torch.cuda.synchronize()
print(f"n={n} batch={batch}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch, {16.0 * n * batch / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9:.0f} GB/s")

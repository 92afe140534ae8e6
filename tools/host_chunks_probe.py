"""How many chunks should the host pipeline cut the headline batch into?  Interleaved repeats to cancel drift."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "phase-vocoder_b200")]
import numpy as np, torch, pvb200
S, F, N, H = 1184, 860, 2048, 512
n_in = N + (F - 1) * H
mode = pvb200.MODE_COMPAT if os.environ.get("PV_PROBE_MODE") == "compat" else pvb200.MODE_CORRECTED
pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=mode, window_type=2 if mode else 0, pitch=(1.4983071,))
xf = torch.randn(S, n_in).mul_(0.1).pin_memory(); of = torch.empty((S, 1, F * H), dtype=torch.float32).pin_memory()
xi = (xf * 20000).to(torch.int16).pin_memory(); oi = torch.empty((S, 1, F * H), dtype=torch.int16).pin_memory()
def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
cfgs = (8, 12, 16, 24, 32, 48, 64)
res = {c: ([], []) for c in cfgs}
for c in cfgs:
    os.environ["PV_HOST_CHUNKS"] = str(c)
    pv.process_host(xf, F, out=of); pv.process_host_pcm16(xi, F, out=oi)
for rep in range(5):
    for c in cfgs:
        os.environ["PV_HOST_CHUNKS"] = str(c)
        res[c][0].append(t(lambda: pv.process_host(xf, F, out=of)))
        res[c][1].append(t(lambda: pv.process_host_pcm16(xi, F, out=oi)))
for c in cfgs:
    print("chunks %2d  float median %.1f ms (min %.1f)   pcm16 median %.1f ms (min %.1f)" % (c, np.median(res[c][0]), min(res[c][0]), np.median(res[c][1]), min(res[c][1])))

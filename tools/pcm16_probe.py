import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch, pvb200
S, F, N, H = 1184, 860, 2048, 512
n_in = N + (F - 1) * H
pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_CORRECTED, window_type=2, pitch=(1.4983071,))
xi = torch.randint(-3000, 3000, (S, n_in), dtype=torch.int16).pin_memory()
oi = torch.empty((S, 1, F * H), dtype=torch.int16).pin_memory()
xf = (xi.float() / 32768).pin_memory(); of = torch.empty((S, 1, F * H), dtype=torch.float32).pin_memory()
def t(fn, n=4):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("float host path  %.1f ms" % t(lambda: pv.process_host(xf, F, out=of)))
print("pcm16 host path  %.1f ms" % t(lambda: pv.process_host_pcm16(xi, F, out=oi)))
for c in (1, 4, 8, 16, 32, 64):
    os.environ["PV_HOST_CHUNKS"] = str(c)
    print("chunks=%d  pcm16 %.1f ms   float %.1f ms" % (c, t(lambda: pv.process_host_pcm16(xi, F, out=oi)), t(lambda: pv.process_host(xf, F, out=of))))
del os.environ["PV_HOST_CHUNKS"]
pc = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_COMPAT, window_type=1)
print("compat default: pcm16 %.1f ms  float %.1f ms" % (t(lambda: pc.process_host_pcm16(xi, F, out=oi)), t(lambda: pc.process_host(xf, F, out=of))))

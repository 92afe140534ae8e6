"""Small invocations of every kernel family in one process: a quick end-to-end exercise of the library, sized so
that it also runs under compute-sanitizer (memcheck / racecheck) where that tool is available."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests"), ROOT]
import numpy as np, torch, pvb200

f32 = lambda v: float(np.float32(v))
rng = np.random.default_rng(0)


def sig(S, n):
    return torch.from_numpy((rng.normal(size=(S, n)) * 0.1).astype(np.float32)).cuda()


def corrected(N, H, betas, S, F):
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_CORRECTED, window_type=pvb200.WIN_HANN_PERIODIC,
                             pitch=tuple(f32(b) for b in betas))
    x = sig(S, N + (F - 1) * H + 3)
    st = torch.zeros((S, pv.state_bytes()), dtype=torch.uint8, device="cuda")
    pv.process(x, F, state=st, flags=pvb200.CARRY_OUT)
    pv.process(x, F, state=st, flags=pvb200.CARRY_IN | pvb200.CARRY_OUT)
    pv.aggregate(x, F)
    return pv, x


def compat(N, H, S, F):
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_COMPAT)
    x = sig(S, N + (F - 1) * H + 1)
    st = torch.zeros((S, pv.state_bytes()), dtype=torch.uint8, device="cuda")
    pv.process(x, F, n_analysed=F - 2, state=st, flags=pvb200.CARRY_OUT)
    pv.process(x, F, state=st, flags=pvb200.CARRY_IN)
    return pv, x


for N in (256, 2048):                                   # tuned fused kernels
    corrected(N, N // 4, (1.0, 1.5), 3, 12)
    compat(N, N // 4, 3, 12)
corrected(2048, 512, (0.8,), 1, 200)                    # few streams: frame-range split on the GPU
for N in (128, 4096):                                   # generic / in-place large-window kernels
    corrected(N, N // 4, (1.5,), 2, 9)
    compat(N, N // 4, 2, 9)
corrected(4096, 1024, (1.0, 1.26, 1.5), 1, 6)           # three voices at 4096: generic kernel, state in global memory
corrected(1024, 100, (1.3,), 2, 10)                     # hop not a multiple of 4
# host paths, chunked
os.environ["PV_HOST_CHUNKS"] = "3"
pv, x = corrected(512, 128, (1.2,), 4, 60)
xh = x.cpu().numpy()
pv.process_host(xh, 60)
pv.process_host_pcm16((xh * 20000).astype(np.int16), 60)
# real-time server and stand-alone FFT
rt = pvb200.RealtimeServer(pv, 4, 2)
for _ in range(3):
    rt.step()
rt.close()
for n in (8, 32, 512, 4096, 8192):
    z = torch.randn((5, n), dtype=torch.complex64, device="cuda")
    pv.fft_batch(z)
    pv.fft_batch(z, inverse=True)
os.environ["PV_FFT_PIPELINE"] = "1"
z = torch.randn((9, 1024), dtype=torch.complex64, device="cuda")
pv.fft_batch(z)
# per-frame reference contract
pvc, xc = compat(256, 128, 1, 4)
spec = torch.empty((512,), dtype=torch.complex64, device="cuda")
torch.cuda.synchronize()
print("sanitize smoke done")

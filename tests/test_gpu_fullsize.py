"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle finishes only bounded
samples of these in seconds; tests/test_gpu_compat.py / test_gpu_corrected.py compare those samples directly).

  headline  1184 streams x 860 frames, window 2048, hop 512 (the bench.py workload, 2.1 GB in / 2.1 GB out)
  C5        2 streams x 168 750 frames, window 4096, hop 1024 (one hour of 48 kHz stereo)

Properties, all evaluated on the GPU over EVERY sample of the full-size result:
  * identity: corrected mode with pitch ratio 1 and Hs = Ha reconstructs its input (WOLA with gain Hs / sum w^2)
  * homogeneity: both pipelines are positively homogeneous of degree 1, and scaling by a power of two is exact in
    fp32 -- out(2x) == 2 out(x) bit for bit (phases do not move, magnitudes double)
  * batch invariance: a stream inside the big batch == the same stream processed on its own, bit for bit
  * shift invariance (compat; frames are independent): out(x delayed by one hop) == out(x) delayed by one hop
  * a sample of streams against the fp64 oracle on the first frames (tolerance 100 dB; the +7 semitone output of the full-size
    run itself, decision-aligned as in tests/aligned.py, and the identity run directly)."""
import numpy as np
import pytest

import pv_oracle as po
from aligned import aligned_parity
from signals import snr_db

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pvb200  # noqa: E402

SEMI7 = float(np.float32(2 ** (7 / 12)))


def make(N, H, mode, betas=(1.0,)):
    return pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=mode, pitch=tuple(betas),
                               window_type=pvb200.WIN_HANN_PERIODIC if mode == pvb200.MODE_CORRECTED else pvb200.WIN_HAMMING)


def signal(S, n, seed):
    """Seeded noisy multitone batch generated on the device (as SURVEY 8d C4's generator: three partials + noise)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.arange(n, device="cuda", dtype=torch.float32)[None, :] / 44100.0
    x = torch.randn((S, n), device="cuda", generator=g) * 1e-3
    for _ in range(3):
        f = 80 + torch.rand((S, 1), device="cuda", generator=g) * 7920
        a = 0.1 + torch.rand((S, 1), device="cuda", generator=g) * 0.2
        x += a * torch.sin(2 * np.pi * f * t)
    return x


def snr_t(want, got):
    num = (want.double() ** 2).sum()
    den = ((want.double() - got.double()) ** 2).sum()
    return float(10 * torch.log10(num / den))


@pytest.mark.parametrize("S,F,N,H", [(1184, 860, 2048, 512), (2, 168750, 4096, 1024)])
def test_corrected_full_size(S, F, N, H):
    n = N + (F - 1) * H
    x = signal(S, n, 7)
    ident = make(N, H, pvb200.MODE_CORRECTED, (1.0,))
    y = ident.process(x, F)
    # identity away from the first window (the OLA needs N/H frames to fill) over every stream
    assert snr_t(x[:, N:F * H], y[:, 0, N:]) > 100
    pv = make(N, H, pvb200.MODE_CORRECTED, (SEMI7,))
    out = pv.process(x, F)
    assert torch.isfinite(out).all()
    # homogeneity, exact for a power of two
    assert torch.equal(pv.process(x * 2, F), out * 2)
    # batch invariance on a few streams
    for s in sorted({0, S // 2, S - 1}):
        alone = pv.process(x[s:s + 1].contiguous(), F)
        assert torch.equal(alone[0], out[s])
    # energy: a pitch shift moves partials, it does not create or lose level (bin remap: within 6 dB)
    ein, eout = float((x[:, N:].double() ** 2).mean()), float((out[:, 0, N:].double() ** 2).mean())
    assert 0.25 < eout / ein < 4.0
    # a sample against the fp64 oracle
    cf = 48
    win = po.window(po.WIN_HANN_PERIODIC, N)
    for s in (0, S - 1):
        xs = x[s, :N + cf * H].cpu().numpy()
        want, _ = po.process_corrected(xs, N, H, H, win, [1.0], cf)
        assert snr_db(want[0], y[s, 0, :cf * H].cpu().numpy()) > 100
        # the pitch-shifted output of the FULL-SIZE launch (not a re-run): its first cf frames depend only on the first
        # cf frames of the input
        D = pv.unwrap_decisions(x[s, :N + cf * H].contiguous(), cf).cpu().numpy()
        r = aligned_parity(xs, N, H, H, win, [SEMI7], cf, D, out[s, :, :cf * H].cpu().numpy())
        assert r["phase_ratio"] < 1.0 and r["frac"] < 1e-3 and min(r["aligned"]) > 100, r


@pytest.mark.parametrize("S,F,N,H", [(1184, 860, 2048, 512), (2, 168750, 4096, 1024)])
def test_compat_full_size(S, F, N, H):
    n = N + (F - 1) * H
    x = signal(S, n + H, 11)
    pv = make(N, H, pvb200.MODE_COMPAT)
    out = pv.process(x[:, :n].contiguous(), F)
    assert torch.isfinite(out).all()
    assert torch.equal(pv.process((x[:, :n] * 2).contiguous(), F), out * 2)
    # delayed by one hop: frame k of the delayed run is frame k+1 of the original; the overlap-add of the first
    # (N-1)//H frames starts from a different history, everything after is identical
    halo = (N - 1) // H
    shifted = pv.process(x[:, H:n + H].contiguous(), F)
    assert torch.equal(shifted[:, :, halo * H:(F - 1) * H], out[:, :, (halo + 1) * H:])
    for s in sorted({0, S // 2, S - 1}):
        alone = pv.process(x[s:s + 1, :n].contiguous(), F)
        assert torch.equal(alone[0], out[s])
    cf = 48
    win = po.window(po.WIN_HAMMING, N)
    for s in (0, S - 1):
        want, _ = po.process_compat(x[s, :N + cf * H].cpu().numpy(), N, H, H, win, cf, cf)
        assert snr_db(want, out[s, 0, :cf * H].cpu().numpy()) > 100


def test_c4_full_size_harmoniser():
    """C4: 4096 streams x 10 s (6890 frames), window 256, hop 64, four voices {1, +4, +7, +12 semitones}: 7.2 GB in,
    28.9 GB out.  Voice 0 (ratio 1) is the identity; homogeneity and batch invariance as above."""
    S, F, N, H = 4096, 6890, 256, 64
    betas = (1.0, float(np.float32(2 ** (4 / 12))), SEMI7, 2.0)
    n = N + (F - 1) * H
    x = signal(S, n, 3)
    pv = make(N, H, pvb200.MODE_CORRECTED, betas)
    out = pv.process(x, F)
    assert out.shape == (S, 4, F * H)
    assert snr_t(x[:, N:F * H], out[:, 0, N:]) > 100
    for s in (0, 2047, 4095):
        alone = pv.process(x[s:s + 1].contiguous(), F)
        assert torch.equal(alone[0], out[s])
    x *= 2
    out2 = pv.process(x, F)
    out *= 2
    assert torch.equal(out2, out)
    del out2
    # every voice keeps the level of its input within the bin-remap tolerance (6 dB)
    ein = float((x[:64, N:].double() ** 2).mean())
    for v in range(4):
        assert 0.25 < float((out[:64, v, N:].double() ** 2).mean()) / ein < 4.0

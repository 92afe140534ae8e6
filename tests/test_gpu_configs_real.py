"""BASELINE.json configs C1 and C2 on the reference's REAL input samples (committed raw-integer slices of
testtones/440sine.wav and testtones/MAT_ZO_24_bit.wav, tests/golden/golden_wav.npz), through the C ABI, both modes.

compat    = the reference's own arithmetic (pinned by output/testout.wav): SNR >= 100 dB against the fp64 oracle.
corrected = pitch x1.5 (C1) / +7 semitones (C2): the stage the reference never implemented (PARITY UNPINNED); SNR >= 100 dB
            against the fp64 oracle with the unwrap decisions aligned (tests/aligned.py), per-bin phase parity
            within the bin's fp32 uncertainty, flips rare.  The window is the one the reference's constructor builds
            (Hamming, src/phaseVocoder.h:85-89) and, separately, the periodic Hann of its one-argument constructor."""
import numpy as np
import pytest

import pv_oracle as po
from aligned import aligned_parity
from signals import snr_db

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pvb200  # noqa: E402

f32 = lambda b: float(np.float32(b))
CONFIGS = {"c1": (256, 64, f32(1.5), 512), "c2": (2048, 512, f32(2 ** (7 / 12)), 64)}


@pytest.mark.parametrize("cfg", ["c1", "c2"])
def test_compat_on_the_reference_wav(golden_wav, cfg):
    N, H, _, nf = CONFIGS[cfg]
    x = golden_wav[cfg]                                       # [2 channels, n]
    assert x.shape[1] == N + (nf - 1) * H
    pv = pvb200.PhaseVocoder(N, effect="t", scale=1, hop=N // H)             # the reference's constructor call, main.cpp:84
    assert (pv.hopSize, pv.outHopSize) == (H, H)
    got = pv.process(torch.from_numpy(x).cuda(), nf).cpu().numpy()
    for c in range(2):
        want, _ = po.process_compat(x[c], N, H, H, pv.imp, nf, nf)
        assert snr_db(want, got[c, 0]) > 100, (cfg, c, snr_db(want, got[c, 0]))


@pytest.mark.parametrize("cfg,wt", [("c1", pvb200.WIN_HAMMING), ("c1", pvb200.WIN_HANN_PERIODIC),
                                    ("c2", pvb200.WIN_HAMMING), ("c2", pvb200.WIN_HANN_PERIODIC)])
def test_pitch_shift_on_the_reference_wav(golden_wav, cfg, wt):
    N, H, beta, nf = CONFIGS[cfg]
    x = golden_wav[cfg]
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_CORRECTED, window_type=wt, pitch=(beta,))
    xd = torch.from_numpy(x).cuda()
    got = pv.process(xd, nf).cpu().numpy()
    for c in range(2):
        D = pv.unwrap_decisions(xd[c], nf).cpu().numpy()
        r = aligned_parity(x[c], N, H, H, pv.imp, [beta], nf, D, got[c])
        print(f"{cfg} window_type={wt} ch{c}: direct {r['direct'][0]:.1f} dB, aligned {r['aligned'][0]:.1f} dB, "
              f"flips {r['flips']} (significant-bin fraction {r['frac']:.2e}), phase ratio {r['phase_ratio']:.3f}")
        assert r["phase_ratio"] < 1.0, r
        assert r["frac"] < 1e-3, r
        assert min(r["aligned"]) > 100, r


def test_process_host_pcm24_applies_audiofile_rules(golden_wav):
    """Packed 24-bit PCM in and out (C2's real samples, as stored in MAT_ZO_24_bit.wav): the conversions on the device ==
    the float path with AudioFile's 24-bit rules applied on the host, byte for byte; both modes; chunked pipeline too."""
    import wav_oracle as wo
    v = golden_wav["c2_raw"]                                 # [2, n] int32, 24 significant bits
    raw = wo.s24_to_bytes(v)                                   # [2, n, 3] uint8, the file's bytes
    assert np.array_equal(wo.bytes_to_s24(raw), v)
    xf = (v.astype(np.float32) / np.float32(8388608))
    N, H = 2048, 512
    nf = (v.shape[1] - N) // H + 1
    for mode, pitch in ((pvb200.MODE_COMPAT, 1.0), (pvb200.MODE_CORRECTED, 2.0 ** (7 / 12))):
        pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=mode, window_type=pvb200.WIN_HANN_PERIODIC, pitch=[pitch])
        got = pv.process_host_pcm24(raw, nf)
        assert got.dtype == np.uint8 and got.shape == (2, 1, nf * H, 3)
        ref = pv.process_host(xf, nf)
        assert np.array_equal(got, wo.s24_to_bytes(wo.float_to_s24(ref)))
        # odd row length (device rows are padded to 4 samples) and an output length that is not a multiple of 4
        pv2 = pvb200.PhaseVocoder(256, hop_in=64, hop_out=62, mode=mode, window_type=pvb200.WIN_HANN_PERIODIC, pitch=[pitch])
        r2 = np.ascontiguousarray(raw[:, :256 + 30 * 64 + 3])
        g2 = pv2.process_host_pcm24(r2, 31)
        f2 = pv2.process_host(np.ascontiguousarray(xf[:, :r2.shape[1]]), 31)
        assert np.array_equal(g2, wo.s24_to_bytes(wo.float_to_s24(f2)))


def test_pcm24_host_pipeline_chunks_bit_exact(monkeypatch, golden_wav):
    import wav_oracle as wo
    raw = wo.s24_to_bytes(golden_wav["c2_raw"])
    pv = pvb200.PhaseVocoder(1024, hop_in=256, hop_out=256, mode=pvb200.MODE_CORRECTED, window_type=pvb200.WIN_HANN_PERIODIC, pitch=[1.25])
    nf = 100
    monkeypatch.setenv("PV_HOST_CHUNKS", "1")
    one = pv.process_host_pcm24(raw, nf)
    monkeypatch.setenv("PV_HOST_CHUNKS", "7")
    assert np.array_equal(one, pv.process_host_pcm24(raw, nf))

"""Properties of the oracle itself (CPU only): FFT vs numpy, frame-level == stream-level,
corrected-mode analytic checks (the only pin corrected mode has -- parity unpinned)."""
import numpy as np
import pytest

import pv_oracle as po
from signals import multitone, snr_db


@pytest.mark.parametrize("n", [8, 64, 512, 4096])
def test_fft_matches_numpy(n):
    rng = np.random.default_rng(n)
    z = rng.normal(size=n) + 1j * rng.normal(size=n)
    assert np.allclose(po.fft(z, -1), np.fft.fft(z), atol=1e-9 * n)
    assert np.allclose(po.fft(z, +1), np.fft.ifft(z) * n, atol=1e-9 * n)


def test_windows():
    w = po.window(po.WIN_HAMMING, 256)
    i = np.arange(256)
    assert np.allclose(w, 0.54 - 0.46 * np.cos(2 * np.pi * i / 255), atol=2e-6)
    assert np.allclose(po.window(po.WIN_HANN_SYM, 256), 0.5 * (1 - np.cos(2 * np.pi * i / 255)), atol=2e-6)
    assert np.allclose(po.window(po.WIN_HANN_PERIODIC, 256), 0.5 * (1 - np.cos(2 * np.pi * i / 256)), atol=2e-6)


def test_schedule_matches_main_cpp():
    assert po.reference_schedule(441000, 128, 128) == (3445, 3445)
    assert po.reference_schedule(1024, 256, 256) == (3, 4)      # last synthesised frame is un-analysed
    assert po.reference_schedule(100, 128, 128) == (0, 0)
    assert po.reference_schedule(441000, 1, 128) == (440999, 3445)


@pytest.mark.parametrize("N,Ha,Hs", [(256, 128, 128), (256, 64, 64), (1024, 102, 512), (64, 16, 32)])
def test_frame_api_equals_stream_api(N, Ha, Hs):
    x = multitone(N + 12 * Ha, seed=N + Ha)
    win = po.window(po.WIN_HAMMING, N)
    nf = 10
    out, back = po.process_compat(x, N, Ha, Hs, win, nf, nf)
    b = np.zeros(N)
    chunks = []
    for k in range(nf):
        spec = po.analysis_frame(x[k * Ha:k * Ha + N], win)
        b = po.resynthesis_frame(b, spec, win, Hs)
        chunks.append(b[:Hs].copy())
    assert np.array_equal(np.concatenate(chunks), out)
    assert np.array_equal(b, back)


def test_analysis_matches_numpy_model():
    """Steps A-D written independently with numpy.fft."""
    N = 256
    x = multitone(N, seed=5)
    win = po.window(po.WIN_HAMMING, N)
    f = x.astype(np.float64) * win
    z = np.zeros(2 * N)
    z[:N // 2] = f[N // 2:]
    z[3 * N // 2:] = f[:N // 2]
    X = np.fft.fft(z)
    got = po.analysis_frame(x, win)
    assert np.allclose(got[:, 0], np.abs(X), atol=1e-10)
    assert np.allclose(got[:, 1], np.arctan(X.imag / X.real), atol=1e-7)


def test_resynthesis_collapses_to_abs_re():
    """D1+D2 collapse: re' = |Re X|, im' = Re X * Im X / |X| (SURVEY 7 'hard parts')."""
    N, Hs = 256, 128
    x = multitone(N, seed=9)
    win = po.window(po.WIN_HAMMING, N)
    f = x.astype(np.float64) * win
    z = np.zeros(2 * N)
    z[:N // 2] = f[N // 2:]
    z[3 * N // 2:] = f[:N // 2]
    X = np.fft.fft(z)[:N // 2 + 1]
    Y = np.abs(X.real) + 1j * (X.real * X.imag / np.abs(X))
    Y[0] = Y[0].real
    Y[-1] = Y[-1].real
    y = np.fft.irfft(Y, N)
    y = np.roll(y, N // 2) * win
    got = po.resynthesis_frame(np.zeros(N), po.analysis_frame(x, win), win, Hs)
    assert np.allclose(got, y, atol=1e-9)


def test_nan_flag():
    N = 64
    win = po.window(po.WIN_HAMMING, N)
    z = np.zeros(N, np.float32)
    assert not np.isnan(po.analysis_frame(z, win)).any()
    assert np.isnan(po.analysis_frame(z, win, flags=po.FLAG_NAN_COMPAT)[:, 1]).all()


# ---------------- corrected mode (specification) ----------------

def test_corrected_identity_is_near_perfect_reconstruction():
    """beta = 1, Ha = Hs: psi_k == P_k exactly, so Y == X and WOLA-normalised OLA returns x."""
    N, H = 1024, 256
    x = multitone(N + 60 * H, seed=3)
    win = po.window(po.WIN_HANN_PERIODIC, N)
    nf = 56
    out, st = po.process_corrected(x, N, H, H, win, [1.0], nf)
    lo, hi = N, nf * H                      # fully overlapped region
    assert snr_db(x[lo:hi], out[0, lo:hi]) > 140
    # the accumulators telescope exactly to the last analysis phase
    assert np.array_equal((st.psi[0] >> np.uint64(32)).astype(np.uint32), st.P_prev)


@pytest.mark.parametrize("beta", [1.5, 2 ** (7 / 12), 0.75])
def test_corrected_pitch_shift_moves_a_sine(beta):
    N, H, fs = 2048, 512, 44100.0
    f0 = 1000.0
    n = N + 80 * H
    x = (0.25 * np.sin(2 * np.pi * f0 * np.arange(n) / fs)).astype(np.float32)
    win = po.window(po.WIN_HANN_PERIODIC, N)
    out, _ = po.process_corrected(x, N, H, H, win, [beta], 80)
    seg = out[0, 8 * H:72 * H]
    spec = np.abs(np.fft.rfft(seg * np.hanning(len(seg))))
    fpk = np.argmax(spec) * fs / len(seg)
    assert abs(fpk - beta * f0) < 0.004 * beta * f0
    # the plain bin-remap scheme (DESIGN.md "corrected mode") keeps the frequency exact but not the
    # lobe shape: level is only preserved to within ~5 dB at 75 % overlap -- a documented limit
    rms = np.sqrt(np.mean(seg ** 2))
    assert abs(20 * np.log10(rms / (0.25 / np.sqrt(2)))) < 6.0


def test_corrected_time_stretch_keeps_pitch():
    N, Ha, Hs, fs = 1024, 128, 256, 44100.0
    f0 = 1234.0
    n = N + 200 * Ha
    x = (0.25 * np.sin(2 * np.pi * f0 * np.arange(n) / fs)).astype(np.float32)
    win = po.window(po.WIN_HANN_PERIODIC, N)
    out, _ = po.process_corrected(x, N, Ha, Hs, win, [1.0], 200)
    assert out.shape[1] == 200 * Hs                      # duration x2
    seg = out[0, 8 * Hs:190 * Hs]
    spec = np.abs(np.fft.rfft(seg * np.hanning(len(seg))))
    fpk = np.argmax(spec) * fs / len(seg)
    assert abs(fpk - f0) < 2.0
    rms = np.sqrt(np.mean(seg ** 2))
    assert abs(20 * np.log10(rms / (0.25 / np.sqrt(2)))) < 0.5


def test_corrected_state_continuation_is_bit_exact():
    """Splitting a stream at any frame and carrying {P_prev, psi, tail} reproduces the
    single-pass result bit for bit (integer phase path => associative)."""
    N, Ha, Hs = 512, 100, 128
    x = multitone(N + 64 * Ha, seed=11)
    win = po.window(po.WIN_HANN_PERIODIC, N)
    betas = [1.0, 2 ** (4 / 12), 1.5]
    full, _ = po.process_corrected(x, N, Ha, Hs, win, betas, 60)
    a, st = po.process_corrected(x, N, Ha, Hs, win, betas, 23)
    b, _ = po.process_corrected(x[23 * Ha:], N, Ha, Hs, win, betas, 37, state=st)
    assert np.array_equal(np.concatenate([a, b], axis=1), full)


def test_corrected_aggregate_reproduces_phase_carry():
    """psi after K frames == P_0[a]<<32 + (K-1)*nomS + Rq*sum(D) -- the frame-range scan identity
    that the segmented GPU path and the multi-GPU carry exchange rely on."""
    N, Ha, Hs, beta = 512, 128, 160, 2 ** (7 / 12)
    x = multitone(N + 50 * Ha, seed=21)
    win = po.window(po.WIN_HANN_PERIODIC, N)
    K = 40
    _, st = po.process_corrected(x, N, Ha, Hs, win, [beta], K)
    t = po.corrected_tables(N, Ha, Hs, beta)
    _, P0 = po.corrected_aggregate(x, N, Ha, win, 1)
    sumD, P_last = po.corrected_aggregate(x[Ha:], N, Ha, win, K - 1, P_prev=P0)
    assert np.array_equal(P_last, st.P_prev)
    ok = t["a_lo"] <= t["a_hi"]
    a = t["a_hi"][ok]
    want = (P0[a].astype(np.uint64) << np.uint64(32)) + np.uint64(K - 1) * t["nomS"][ok] \
        + (sumD[a] * np.int64(t["Rq"])).astype(np.uint64)
    assert np.array_equal(st.psi[0][ok], want)


def test_unwrap_conditioning_f32_vs_f64_oracle():
    """DESIGN.md "conditioning of the phase unwrap": the fp32 and fp64 variants of the SAME oracle agree
    to > 100 dB for integer R but can unwrap weak bins differently when R is fractional."""
    N, Ha, Hs, nf = 1024, 256, 256, 60
    win = po.window(po.WIN_HANN_PERIODIC, N)
    tonal = multitone(N + nf * Ha, seed=7, noise=0.0)
    a, _ = po.process_corrected(tonal, N, Ha, Hs, win, [1.4983], nf, precision=64)
    b, _ = po.process_corrected(tonal, N, Ha, Hs, win, [1.4983], nf, precision=32)
    assert snr_db(a[0], b[0]) > 60           # fractional R: unwrap flips on weak bins are possible
    noisy = multitone(N + nf * Ha, seed=7, noise=1e-3)
    a, _ = po.process_corrected(noisy, N, Ha, Hs, win, [2.0], nf, precision=64)      # integer R
    b, _ = po.process_corrected(noisy, N, Ha, Hs, win, [2.0], nf, precision=32)
    assert snr_db(a[0], b[0]) > 100


@pytest.mark.parametrize("N,H,beta,nf,kind", [(256, 64, 1.5, 2000, "sine"), (1024, 256, 2 ** (7 / 12), 300, "tones")])
def test_decision_aligned_parity_between_the_f32_and_f64_oracle(N, H, beta, nf, kind):
    """The method the GPU parity tests use (tests/aligned.py), demonstrated between the two precisions of the
    oracle on clean tonal inputs, the worst case for the unwrap: most bins hold leakage 80-150 dB below the peak.
    Directly the two disagree (flipped aliases, ~80-95 dB); every disagreement is a boundary case within the
    bin's fp32 phase uncertainty; with those few decisions aligned the outputs agree to > 100 dB."""
    from aligned import aligned_parity
    if kind == "sine":
        x = (0.25 * np.sin(2 * np.pi * 440 * np.arange(N + nf * H) / 44100)).astype(np.float32)
    else:
        x = multitone(N + nf * H, seed=3, noise=0.0)
    win = po.window(po.WIN_HANN_PERIODIC, N)
    out32, D32, _ = po.process_corrected_traced(x, N, H, H, win, [beta], nf, precision=32)
    r = aligned_parity(x, N, H, H, win, [beta], nf, D32, out32)
    assert r["phase_ratio"] < 1.0, r          # per-bin phase parity modulo one turn
    assert r["frac"] < 1e-3, r                # flips are rare among the bins that carry energy
    assert min(r["aligned"]) > 100, r         # and nothing else differs
    assert min(r["direct"]) > 60, r


def test_24_bit_sample_rules_round_trip_and_match_the_wav_decoder():
    """The checker side of pv_process_host_pcm24 (oracle/wav_oracle.py): bytes <-> sign-extended ints round trip, the encode rule
    is AudioFile's `(int32)(x * 8388608.)` (truncation toward zero, no clamp, low three bytes), and the helper decodes a 24-bit
    WAV `data` chunk exactly like decode_wav (AudioFile.h:508-518)."""
    import struct

    import wav_oracle as wo
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.integers(-(1 << 23), 1 << 23, size=4000), [-(1 << 23), (1 << 23) - 1, -1, 0, 1]]).astype(np.int32)
    b = wo.s24_to_bytes(v)
    assert b.shape == (len(v), 3) and b.dtype == np.uint8
    assert np.array_equal(wo.bytes_to_s24(b), v)
    x = v.astype(np.float32) / np.float32(8388608)                     # the decode rule; exact in fp32
    assert np.array_equal(wo.float_to_s24(x), v)                       # ... so encoding gives the integers back
    assert wo.float_to_s24(np.float32([0.9999999, -0.9999999, 0.5 / 8388608, -0.5 / 8388608])).tolist() == [8388607, -8388607, 0, 0]
    over = wo.float_to_s24(np.float32([1.0, 1.5]))                     # no clamp: the low three bytes wrap
    assert wo.bytes_to_s24(wo.s24_to_bytes(over)).tolist() == [-(1 << 23), -(1 << 22)]
    # a mono 24-bit WAV around the same bytes
    pcm = b.tobytes()
    hdr = b"RIFF" + struct.pack("<i", 36 + len(pcm)) + b"WAVE" + b"fmt " + struct.pack("<ihhiihh", 16, 1, 1, 44100, 3 * 44100, 3, 24) + \
        b"data" + struct.pack("<i", len(pcm))
    s, rate, bits = wo.decode_wav(hdr + pcm)
    assert bits == 24 and rate == 44100 and np.array_equal(s[0], x)

"""Pins the compat-mode oracle to the reference's own committed outputs.

Tolerance: +-1 LSB of 16 bit after the reference's own float->int16 rule
(AudioFile.h:1045-1049); the golden WAVs were produced by fp32 cuFFT arithmetic, the oracle
truth is fp64, so the two can straddle a truncation boundary."""
import hashlib
import os

import numpy as np
import pytest

import pv_oracle as po
import wav_oracle as wo

N, H = 256, 128


def _s16_to_float(a):
    return a.astype(np.float32) / np.float32(32768.0)


@pytest.mark.parametrize("precision", [64, 32])
def test_testout_head(golden, precision):
    x = _s16_to_float(golden["testout_head_in"])
    want = golden["testout_head_out"].astype(np.int32)
    win = po.window(po.WIN_HAMMING, N)
    nf = len(want) // H
    out, _ = po.process_compat(x, N, H, H, win, n_analysed=nf, n_frames=nf, precision=precision)
    got = wo.float_to_s16(out.astype(np.float32)).astype(np.int32)
    assert np.abs(got - want).max() <= 1
    # main.cpp:288-290: `if(numChannels = 1)` duplicates channel 0 into channel 1
    assert np.array_equal(golden["testout_head_out"], golden["testout_head_out_ch1"])


def test_testout_tail_zero_fill(golden):
    """Frames near the end read past the input; zero fill reproduces the golden tail and the
    last numSamples - nS*Hs samples stay 0."""
    x_tail = _s16_to_float(golden["testout_tail_in"])
    k0 = int(golden["testout_tail_first_frame"])
    n = int(golden["num_samples"])
    nA, nS = po.reference_schedule(n, H, H)
    assert (nA, nS) == (3445, 3445)
    win = po.window(po.WIN_HAMMING, N)
    nf = nS - k0
    # local frame numbering: frame j here is stream frame k0+j; analysed while k0+j < nA
    out, _ = po.process_compat(x_tail, N, H, H, win, n_analysed=nA - k0, n_frames=nf)
    got = wo.float_to_s16(out.astype(np.float32)).astype(np.int32)[H:]   # first hop lacks its left neighbour
    want = golden["testout_tail_out"].astype(np.int32)
    assert len(want) == n - (k0 + 1) * H
    assert np.abs(got - want[:len(got)]).max() <= 1
    assert not want[len(got):].any()


def test_sine1000_head(golden):
    """1000hzout.wav pins independent analysis/synthesis hops (Ha=1, Hs=128) and the symmetric Hann."""
    x = _s16_to_float(golden["sine1000_head_in"])
    want = golden["sine1000_head_out"].astype(np.int32)
    win = po.window(po.WIN_HANN_SYM, N)
    nf = len(want) // H
    out, _ = po.process_compat(x, N, 1, H, win, n_analysed=nf, n_frames=nf)
    got = wo.float_to_s16(out.astype(np.float32)).astype(np.int32)
    assert np.abs(got - want).max() <= 1


def test_full_files_when_reference_present(golden, reference_dir):
    names = ["testtones/test.wav", "output/testout.wav", "testtones/1000sine.wav", "output/1000hzout.wav"]
    blobs = [open(os.path.join(reference_dir, n), "rb").read() for n in names]
    for b, h in zip(blobs, golden["sha256"]):
        assert hashlib.sha256(b).hexdigest() == str(h)
    for (bi, bo, Ha, wt) in ((blobs[0], blobs[1], 128, po.WIN_HAMMING), (blobs[2], blobs[3], 1, po.WIN_HANN_SYM)):
        x, rate, bits = wo.decode_wav(bi)
        g, _, _ = wo.decode_wav(bo)
        assert rate == 44100 and bits == 16 and g.shape[0] == 2
        want = np.round(g[0].astype(np.float64) * 32768).astype(np.int32)
        nA, nS = po.reference_schedule(x.shape[1], Ha, H)
        out, _ = po.process_compat(x[0], N, Ha, H, po.window(wt, N), nA, nS)
        got = wo.float_to_s16(out.astype(np.float32)).astype(np.int32)
        assert len(got) == 440960
        assert np.abs(got - want[:len(got)]).max() <= 1


def test_wav_roundtrip_rules():
    x = np.array([[0.0, 0.5, -0.5, 1.5, -1.5, 0.99999, 1e-5, -1e-5]], np.float32)
    blob = wo.encode_wav16(np.vstack([x, x]))
    assert len(blob) == 44 + 2 * 2 * x.shape[1]
    y, rate, bits = wo.decode_wav(blob)
    assert rate == 44100 and bits == 16
    want = np.array([0, 16383, -16383, 32767, -32767, 32766, 0, 0], np.float32) / np.float32(32768)
    assert np.array_equal(y[0], want) and np.array_equal(y[1], want)


def test_wav_slices_equal_the_reference_files(golden_wav, reference_dir):
    """tests/golden/golden_wav.npz (C1 / C2 inputs) against the files themselves, decoded by the numpy restatement
    of AudioFile's rules (which tests/test_wav_and_cli.py checks against src/AudioFile.h itself)."""
    for name, key, off, h in (("testtones/440sine.wav", "c1", 0, golden_wav["sha256"][0]),
                              ("testtones/MAT_ZO_24_bit.wav", "c2", golden_wav["c2_offset"], golden_wav["sha256"][1])):
        blob = open(os.path.join(reference_dir, name), "rb").read()
        assert hashlib.sha256(blob).hexdigest() == h
        x, rate, bits = wo.decode_wav(blob)
        assert rate == 44100 and x.shape == (2, golden_wav[key + "_num_samples"])
        sl = golden_wav[key]
        assert np.array_equal(x[:, off:off + sl.shape[1]], sl)
    assert bits == 24

"""Stand-alone batched FFT (pv_fft_batch) against the fp64 oracle FFT and numpy.
Tolerance: fp32 butterflies, error grows ~ sqrt(log n) ulps -> SNR >= 125 dB for n <= 8192."""
import numpy as np
import pytest
import torch

import pvb200
from oracle import pv_oracle as po

pytestmark = pytest.mark.gpu


def snr_db(want, got):
    want, got = np.asarray(want, np.complex128), np.asarray(got, np.complex128)
    num = np.sum(np.abs(want) ** 2)
    den = np.sum(np.abs(want - got) ** 2)
    return np.inf if den == 0 else 10 * np.log10(num / den)


@pytest.fixture(scope="module")
def pv():
    return pvb200.PhaseVocoder(256)


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_forward_and_inverse_match_fp64(pv, n):
    rng = np.random.default_rng(n)
    for batch in (1, 3, 133):
        z = (rng.normal(size=(batch, n)) + 1j * rng.normal(size=(batch, n))).astype(np.complex64)
        zd = torch.from_numpy(z).cuda()
        f = pv.fft_batch(zd).cpu().numpy()
        assert snr_db(np.fft.fft(z.astype(np.complex128), axis=1), f) > 125
        i = pv.fft_batch(zd, inverse=True).cpu().numpy()
        assert snr_db(np.fft.ifft(z.astype(np.complex128), axis=1) * n, i) > 125      # unnormalised, like cuFFT
    # same convention as the oracle's own FFT (the one the golden WAVs pin through steps C and F)
    assert snr_db(po.fft(z[0].astype(np.complex128), -1), f[0]) > 125
    # in place, and forward o inverse = n * identity
    buf = zd.clone()
    pv.fft_batch(buf, out=buf)
    pv.fft_batch(buf, inverse=True, out=buf)
    assert snr_db(z.astype(np.complex128) * n, buf.cpu().numpy()) > 120


@pytest.mark.parametrize("n,batch", [(32, 64 * 150 + 5), (128, 16 * 300 + 3), (512, 4 * 149 + 1), (2048, 743), (4096, 301), (8192, 299)])
def test_large_batches_take_the_pipelined_kernel(pv, n, batch):
    """More tiles than SMs: persistent CTAs with the next tile's input in flight; ragged last tile; in place too."""
    rng = np.random.default_rng(batch)
    z = (rng.normal(size=(batch, n)) + 1j * rng.normal(size=(batch, n))).astype(np.complex64)
    zd = torch.from_numpy(z).cuda()
    want = np.fft.fft(z.astype(np.complex128), axis=1)
    assert snr_db(want, pv.fft_batch(zd).cpu().numpy()) > 125
    pv.fft_batch(zd, out=zd)
    assert snr_db(want, zd.cpu().numpy()) > 125
    # an input that is not 16-byte aligned falls back to the plain kernel
    flat = torch.zeros(batch * n + 1, dtype=torch.complex64, device="cuda")
    flat[1:] = torch.from_numpy(z).cuda().reshape(-1)
    assert snr_db(want, pv.fft_batch(flat[1:].view(batch, n)).cpu().numpy()) > 125


def test_reference_benchmark_signals(pv):
    """The inputs of the reference's FFT timing runs (src/50Hz/*.dat etc.: 0.1 sin(2 pi f t) at 44.1 kHz, 128..65536
    samples) as real sequences in the complex input, sizes the batched kernel covers."""
    t = np.arange(8192) / 44100.0
    sig = 0.1 * np.sin(2 * np.pi * 500 * t) + 0.1 * np.sin(2 * np.pi * 505 * t + 2.345) + 0.1 * np.sin(2 * np.pi * 12000 * t - 0.884)
    for n in (128, 512, 2048, 8192):
        z = sig[:n].astype(np.complex64)[None, :]
        f = pv.fft_batch(torch.from_numpy(z).cuda()).cpu().numpy()[0]
        assert snr_db(np.fft.fft(sig[:n].astype(np.float32).astype(np.float64)), f) > 125
        if n >= 2048:       # the 500 Hz and 12 kHz lines stand out of the leakage floor
            mag = np.abs(f[:n // 2])
            for hz in (500, 12000):
                b = round(hz * n / 44100)
                assert mag[b - 1:b + 2].max() > 20 * np.median(mag)


def test_bad_sizes_are_rejected(pv):
    with pytest.raises(pvb200.PvError):
        pv.fft_batch(torch.zeros((2, 24), dtype=torch.complex64, device="cuda"))
    with pytest.raises(pvb200.PvError):
        pv.fft_batch(torch.zeros((1, 16384), dtype=torch.complex64, device="cuda"))

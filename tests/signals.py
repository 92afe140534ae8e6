"""Seeded synthetic inputs of the shapes BASELINE.json / SURVEY 8d name (test + bench helper)."""
import numpy as np


def multitone(n, fs=44100.0, seed=0, noise=1e-3, n_tones=3):
    """C4 generator: sum of 3 sines, f uniform in [80, 8000] Hz, amp 0.1-0.3, plus N(0, noise)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    x = np.zeros(n)
    for _ in range(n_tones):
        f = rng.uniform(80.0, 8000.0)
        a = rng.uniform(0.1, 0.3)
        ph = rng.uniform(0, 2 * np.pi)
        x += a * np.sin(2 * np.pi * f * t + ph)
    x += rng.normal(0.0, noise, n)
    return x.astype(np.float32)


def c3_multitone(n, fs=44100.0):
    """0.1 sin(2pi 500 t) + 0.1 sin(2pi 505 t + 2.345) + 0.1 sin(2pi 12000 t - 0.884)
    (matches src/500Hz+505Hz+12000Hz/*.dat of the reference)."""
    t = np.arange(n, dtype=np.float64) / fs
    x = 0.1 * np.sin(2 * np.pi * 500 * t) + 0.1 * np.sin(2 * np.pi * 505 * t + 2.345) \
        + 0.1 * np.sin(2 * np.pi * 12000 * t - 0.884)
    return x.astype(np.float32)


def snr_db(ref, got):
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    err = np.sum((ref - got) ** 2)
    sig = np.sum(ref ** 2)
    if err == 0:
        return np.inf
    return 10 * np.log10(sig / err)

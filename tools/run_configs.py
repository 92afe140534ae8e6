#!/usr/bin/env python
"""Runs the five BASELINE.json configs (C1..C5) on one GPU: parity against the oracle on a
bounded sample of each and device-resident throughput.  Prints a markdown table (kept in
profiles/r01_configs.md).  Usage on the GPU box: python tools/run_configs.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import pv_oracle as po
import pvb200
from aligned import aligned_parity
from signals import c3_multitone, multitone, snr_db

f32 = lambda b: float(np.float32(b))


def timed(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def golden_wav():
    """Committed raw-integer slices of the reference's C1 / C2 input WAVs (tests/golden/golden_wav.npz), decoded with
    AudioFile's rules."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_wav.npz"))
    return (g["c1_440sine"].astype(np.float32) / np.float32(32768.0), g["c2_matzo"].astype(np.float32) / np.float32(8388608.0))


def tiled(x):
    """Generator that repeats a real slice to the requested length (timing rows; parity uses the slice itself)."""
    return lambda n, s: np.tile(x[s % len(x)], n // x.shape[1] + 1)[:n]


def run(name, N, Ha, Hs, mode, betas, streams, frames, fs, gen, check_streams=2, check_frames=None, wt=None):
    """Parity column: compat = worst direct SNR against the fp64 oracle; corrected = "direct / decision-aligned" SNR
    (tests/aligned.py: the unwrap is discontinuous, a few boundary decisions may differ between fp32 and fp64)."""
    corrected = mode == "corrected"
    if wt is None:
        wt = pvb200.WIN_HANN_PERIODIC if corrected else pvb200.WIN_HAMMING
    pv = pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs, mode=pvb200.MODE_CORRECTED if corrected else pvb200.MODE_COMPAT,
                             window_type=wt, pitch=tuple(betas))
    n_in = N + (frames - 1) * Ha
    xs = np.stack([gen(n_in, s) for s in range(min(streams, 16))])
    x = torch.from_numpy(np.tile(xs, ((streams + len(xs) - 1) // len(xs), 1))[:streams]).cuda()
    out = torch.empty((streams, len(betas) if corrected else 1, frames * Hs), device="cuda")
    ms = timed(lambda: pv.process(x, frames, out=out))
    cf = frames if check_frames is None else min(frames, check_frames)
    # parity on the first cf frames of the checked streams (a separate short run: same kernels, same arithmetic)
    got = pv.process(x[:check_streams, :N + (cf - 1) * Ha].contiguous(), cf).cpu().numpy()
    worst, worst_al, flips = 1e9, 1e9, 0
    win = pv.imp
    for s in range(check_streams):
        xc = xs[s][:N + (cf - 1) * Ha]
        if corrected:
            D = pv.unwrap_decisions(torch.from_numpy(xc).cuda(), cf).cpu().numpy()
            r = aligned_parity(xc, N, Ha, Hs, win, betas, cf, D, got[s])
            assert r["phase_ratio"] < 1.0, r
            worst, worst_al, flips = min(worst, min(r["direct"])), min(worst_al, min(r["aligned"])), flips + r["flips"]
        else:
            w, _ = po.process_compat(xc, N, Ha, Hs, win, cf, cf)
            worst = min(worst, snr_db(w, got[s, 0]))
    fps = streams * frames / (ms * 1e-3)
    V = len(betas) if corrected else 1
    gbs = fps * (4 * Ha + 4 * V * Hs) / 1e9
    par = f"{worst:.1f} / {worst_al:.1f} ({flips} flips)" if corrected else f"{worst:.1f}"
    print(f"| {name} | {N} | {Ha}/{Hs} | {mode} | {V} | {streams} x {frames} | {ms:.3f} | {fps/1e6:.2f} M | "
          f"{fps*Ha/fs:,.0f} | {gbs:.0f} | {par} |", flush=True)


def realtime_step(streams=4096, N=256, H=64, betas=(1.0, 2 ** (4 / 12), 2 ** (7 / 12), 2.0), calls=300):
    """C4 as a real-time server would run it: ONE hop per stream per call, state carried on the device
    (PV_PROCESS_CARRY_IN | CARRY_OUT).  Reports the device time per call against the hop period."""
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_CORRECTED, window_type=pvb200.WIN_HANN_PERIODIC,
                             pitch=tuple(f32(b) for b in betas))
    x = torch.randn((streams, N), device="cuda") * 0.1
    st = torch.zeros((streams, pv.state_bytes()), dtype=torch.uint8, device="cuda")
    out = torch.empty((streams, len(betas), H), device="cuda")
    pv.process(x, 1, out=out, state=st, flags=pvb200.CARRY_OUT)
    fl = pvb200.CARRY_IN | pvb200.CARRY_OUT
    ms = timed(lambda: pv.process(x, 1, out=out, state=st, flags=fl), n=calls)
    t0 = time.perf_counter()
    for _ in range(calls):
        pv.process(x, 1, out=out, state=st, flags=fl)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / calls * 1e3
    print(f"\nC4 real-time step: {streams} streams x 1 hop ({H} samples) x {len(betas)} voices per call: "
          f"{ms*1e3:.0f} us device time per call, {wall*1e3:.0f} us wall per call (hop period {H/44100*1e3:.2f} ms, "
          f"window budget {N/44100*1e3:.1f} ms) -> {streams/(ms*1e-3)/1e6:.1f} M frames/s")
    # the same through the block server (pv_rt_*): host buffers in, host buffers out, one graph launch per block
    for S, B in ((streams, 1), (streams, 4), (64, 1)):
        rt = pvb200.RealtimeServer(pv, S, B)
        rt.input[:] = np.random.default_rng(0).normal(size=rt.input.shape).astype(np.float32) * 0.1
        for _ in range(50):
            rt.step()
        ts = []
        for _ in range(2000):
            t0 = time.perf_counter()
            rt.step()
            ts.append(time.perf_counter() - t0)
        ts = np.sort(np.array(ts)) * 1e6
        print(f"C4 block server: {S} streams x {B} hop(s) per block, host buffer to host buffer: median {ts[1000]:.0f} us, "
              f"p99 {ts[1980]:.0f} us, max {ts[-1]:.0f} us per block (block period {B*H/44100*1e3:.2f} ms)")
        rt.close()


def sweep():
    """Window sweep at equal total audio per launch (profiles/r01_window_sweep.md)."""
    print("| mode | window | Ha/Hs | streams x frames | ms/launch | frames/s | audio-s/s | algorithmic GB/s | SNR dB |")
    print("|---|---|---|---|---|---|---|---|---|")
    noisy = lambda n, s: multitone(n, seed=s, noise=1e-3)
    for mode, betas, label in (("compat", [1.0], "compat"), ("corrected", [f32(2 ** (7 / 12))], "corrected (+7 st)")):
        for N in (256, 512, 1024, 2048):
            streams, frames, H = 2368 * 2048 // N // 2, 1720, N // 4
            import io, contextlib
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                run("x", N, H, H, mode, betas, streams, frames, 44100, noisy, check_frames=60)
            c = [t.strip() for t in buf.getvalue().strip().split("|")]
            print(f"| {label} | {N} | {H}/{H} | {streams} x {frames} | {c[7]} | {c[8]} | {c[9]} | {c[10]} | {c[11]} |", flush=True)


def main():
    if "--sweep" in sys.argv:
        return sweep()
    print("| config | window | Ha/Hs | mode | voices | streams x frames | ms/launch | frames/s | audio-s/s (input) | "
          "algorithmic GB/s | worst SNR vs fp64 oracle (dB): compat direct; corrected direct / decision-aligned |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    tone = lambda n, s: multitone(n, seed=s, noise=0.0)
    noisy = lambda n, s: multitone(n, seed=s, noise=1e-3)
    semi = lambda k: f32(2 ** (k / 12))
    # C1: testtones/440sine.wav (committed slice, tiled to the file's 10 s = 6890 frames for the timing), window 256 hop 64:
    # compat plumbing + pitch x1.5; window table of the reference's constructor (Hamming) and the periodic Hann
    c1, c2 = golden_wav()
    run("C1 compat (440sine.wav)", 256, 64, 64, "compat", [1.0], 2, 6890, 44100, tiled(c1), check_frames=512)
    run("C1 pitch x1.5 (440sine.wav, Hamming)", 256, 64, 64, "corrected", [1.5], 2, 6890, 44100, tiled(c1), check_frames=512, wt=pvb200.WIN_HAMMING)
    run("C1 pitch x1.5 (440sine.wav, Hann)", 256, 64, 64, "corrected", [1.5], 2, 6890, 44100, tiled(c1), check_frames=512)
    # C2: testtones/MAT_ZO_24_bit.wav (committed 24-bit slice; the file is 5.58 s stereo = 2 x 480 frames), window 2048 hop 512,
    # +7 semitones; and the headline batch
    run("C2 compat (MAT_ZO_24_bit.wav)", 2048, 512, 512, "compat", [1.0], 2, 480, 44100, tiled(c2), check_frames=64)
    run("C2 +7 st (MAT_ZO_24_bit.wav, Hamming)", 2048, 512, 512, "corrected", [semi(7)], 2, 480, 44100, tiled(c2), check_frames=64, wt=pvb200.WIN_HAMMING)
    run("C2 +7 st (MAT_ZO_24_bit.wav, Hann)", 2048, 512, 512, "corrected", [semi(7)], 2, 480, 44100, tiled(c2), check_frames=64)
    run("C2 headline batch", 2048, 512, 512, "corrected", [semi(7)], 1184, 860, 44100, noisy, check_frames=120)
    run("C2 headline batch", 2048, 512, 512, "compat", [1.0], 1184, 860, 44100, noisy, check_frames=120)
    # C3: window 1024, time stretch "in-hop 10 / out-hop 2": hop divisors (102/512) and literal samples (10/2)
    c3 = lambda n, s: c3_multitone(n)
    run("C3 divisors 10/2", 1024, 102, 512, "compat", [1.0], 1, 4300, 44100, c3, check_streams=1, check_frames=300)
    run("C3 divisors 10/2", 1024, 102, 512, "corrected", [1.0], 1, 4300, 44100, c3, check_streams=1, check_frames=300)
    run("C3 literal 10/2", 1024, 10, 2, "compat", [1.0], 1, 44000, 44100, c3, check_streams=1, check_frames=1500)
    # C4: 4096 streams x window 256 hop 64 x 4 voices, 10 s each
    run("C4 harmoniser", 256, 64, 64, "corrected", [1.0, semi(4), semi(7), 2.0], 4096, 6890, 44100, noisy, check_frames=400)
    # C5: 1 h, 48 kHz, stereo, window 4096 hop 1024 (168 750 frames per channel)
    run("C5 long file", 4096, 1024, 1024, "compat", [1.0], 2, 168750, 48000, noisy, check_frames=300)
    run("C5 long file", 4096, 1024, 1024, "corrected", [semi(7)], 2, 168750, 48000, noisy, check_frames=100)
    realtime_step()


if __name__ == "__main__":
    main()

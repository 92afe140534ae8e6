"""Stand-alone FFT micro-benchmark: the reference times single transforms of 32..1024 points with
its hand-written FFTs against cuFFT (milestone deck slide 7: cuFFT 9.2 .. 14.3 us per transform on its GPU).
Here: pv_fft_batch against cuFFT (through torch.fft, which plans once and caches) on this B200 --
(1) latency of ONE transform per launch, (2) throughput of a large batch against the HBM roofline
(16*n bytes per transform: 8n in, 8n out).  cuFFT is the comparison arm only; the product never calls it."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "phase-vocoder_b200")]
import numpy as np
import torch

import pvb200


def timed(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3          # us


def graph_latency(fn, reps=100):
    """Device-side time of one call in a chain of `reps` dependent launches replayed as a CUDA graph (no Python or
    driver launch cost): us per transform."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        return timed(g.replay, n=20, warm=3) / reps


def main():
    peak = 6545.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    pv = pvb200.PhaseVocoder(256)
    print(f"HBM roofline denominator: {peak:.0f} GB/s\n")
    print("| n | 1 transform per call, C ABI entry: ours (us) | cuFFT through torch.fft (us) | in a CUDA graph: ours (us) | cuFFT (us) | batch | ours (us) | cuFFT (us) | ours GB/s | % of HBM peak | cuFFT GB/s | max err vs cuFFT |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for n in (32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):
        one = (torch.randn(1, n, device="cuda") + 1j * torch.randn(1, n, device="cuda")).to(torch.complex64)
        o1 = torch.empty_like(one)
        # the C ABI call itself, arguments resolved once (what a C caller pays; the Python wrapper's own asserts / pointer
        # look-ups cost another 1 - 2 us per call and are no part of the library)
        lib, hdl, pi, po_, st = pvb200.load(), pv._h, one.data_ptr(), o1.data_ptr(), torch.cuda.current_stream().cuda_stream
        t_one = timed(lambda: lib.pv_fft_batch(hdl, pi, po_, n, 1, -1, st))
        c_one = timed(lambda: torch.fft.fft(one, dim=1))      # allocating the result is its faster call (out= adds a copy)
        g_one, gc_one = graph_latency(lambda: pv.fft_batch(one, out=o1)), graph_latency(lambda: torch.fft.fft(one, dim=1, out=o1))
        batch = (1 << 28) // (8 * n)                 # 256 MB in, 256 MB out: larger than L2
        x = (torch.randn(batch, n, device="cuda") + 1j * torch.randn(batch, n, device="cuda")).to(torch.complex64)
        o = torch.empty_like(x)
        t_b = timed(lambda: pv.fft_batch(x, out=o), n=20, warm=3)
        if os.environ.get("PV_FFT_AB"):
            os.environ["PV_FFT_PIPELINE"] = "0"
            t0 = timed(lambda: pv.fft_batch(x, out=o), n=20, warm=3)
            os.environ["PV_FFT_PIPELINE"] = "1"
            t1 = timed(lambda: pv.fft_batch(x, out=o), n=20, warm=3)
            del os.environ["PV_FFT_PIPELINE"]
            print(f"  n={n}: plain {t0:.0f} us, pipelined {t1:.0f} us, default {t_b:.0f} us", file=sys.stderr)
        ours = o.clone()
        c_b = timed(lambda: torch.fft.fft(x, dim=1), n=20, warm=3)      # no out=: that adds a copy
        o = torch.fft.fft(x, dim=1)
        err = (ours - o).abs().max().item() / o.abs().max().item()
        gb = 16.0 * n * batch / 1e9
        print(f"| {n} | {t_one:.1f} | {c_one:.1f} | {g_one:.2f} | {gc_one:.2f} | {batch} | {t_b:.0f} | {c_b:.0f} | {gb/(t_b*1e-6):.0f} | "
              f"{100*gb/(t_b*1e-6)/peak:.0f} % | {gb/(c_b*1e-6):.0f} | {err:.1e} |", flush=True)
        del x, o, ours


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the fused phase-vocoder hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode compat|corrected]

A "step" is one pass of the fused analysis -> processing -> resynthesis/OLA path over one batch of
synthetic audio streams at the headline shape of BASELINE.json (window 2048, hop 512).  The batch
(>= 2 GB of fp32 input per GPU, far larger than the 126 MB L2) is generated with the C4 generator of
SURVEY 8d.  One process per GPU (torchrun for N > 1); streams are sharded over ranks with no
data-path collective ("weak" scaling: every rank gets its own batch).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` goes through the C ABI host entry point pv_process_host with pinned HOST buffers
(H2D + kernel + D2H inside the timed region); `roofline` is algorithmic bytes / measured kernel
time against the measured HBM peak; `cpu_baseline` is the f32 oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

WINDOW, HOP = 2048, 512
FS = 44100.0
FRAMES_PER_STREAM = 860            # ~10 s of audio per stream
STREAMS_PER_GPU = 1184             # 148 SMs x 8; 1184 * 441856 * 4 B = 2.09 GB of input (> 126 MB L2)
SEMITONES_7 = 2.0 ** (7.0 / 12.0)
L2_NOTE = "inputs (2.1 GB/GPU) and outputs exceed the 126 MB L2; no flush needed"
SHARDING_NOTE = "independent streams per rank, no data-path collective"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("PV_BENCH_MODE", "corrected"), choices=["compat", "corrected"])
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU)
    ap.add_argument("--frames", type=int, default=FRAMES_PER_STREAM)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the long file of config 5 (one GPU: record `c5`; --gpus > 1: frame-range sharded, record `c5_sharded`)")
    return ap.parse_args()


def workload_name(mode, streams, frames):
    return (f"synthetic multitone+noise streams (SURVEY 8d C4 generator), {streams} streams/GPU x {frames} frames, "
            f"window {WINDOW}, hop_in=hop_out={HOP}, mode={mode}"
            + (", pitch +7 semitones" if mode == "corrected" else ""))


def _baseline_metric():
    """BASELINE.json's own metric string (frames/s is the `value`; audio-seconds/s and the roofline fractions ride along
    in `audio_s_per_s` and `roofline`)."""
    try:
        with open(os.path.join(ROOT, "BASELINE.json")) as f:
            return json.load(f)["metric"]
    except Exception:
        return "STFT frames/s & audio-sec/s (N=2048,hop=512) at 1/2/4/8 B200; % HBM roofline"


METRIC = _baseline_metric()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the f32 oracle port on the host cores (the reference has no CPU path)
# ------------------------------------------------------------------------------------------------
def _fast_oracle():
    odir = os.path.join(ROOT, "oracle")
    # rebuild on this host so that -march=native matches the cores we time on
    subprocess.run(["make", "-s", "-B", "-C", odir, "libpv_oracle_fast.so"], check=False,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(odir, "libpv_oracle_fast.so"))
    fp = C.POINTER(C.c_float)
    L.pvo_bench_f32.argtypes = [C.c_int, C.c_double, fp, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, fp, C.c_long,
                                fp, C.c_int]
    L.pvo_bench_f32.restype = C.c_int
    L.pvo_bench_threads.restype = C.c_int
    L.pvo_window.argtypes = [C.c_int, C.c_int, fp]
    return L


def host_streams(n_streams, n_in):
    from signals import multitone
    return np.stack([multitone(n_in, fs=FS, seed=s) for s in range(n_streams)])


def cpu_port_run(L, x, frames, threads, mode="compat"):
    fp = C.POINTER(C.c_float)
    corrected = mode == "corrected"
    win = np.empty(WINDOW, np.float32)
    L.pvo_window(2 if corrected else 0, WINDOW, win.ctypes.data_as(fp))
    out = np.empty((x.shape[0], frames * HOP), np.float32)
    t0 = time.perf_counter()
    rc = L.pvo_bench_f32(1 if corrected else 0, float(np.float32(SEMITONES_7)), x.ctypes.data_as(fp), x.shape[0],
                         x.shape[1], WINDOW, HOP, HOP, win.ctypes.data_as(fp), frames, out.ctypes.data_as(fp), threads)
    dt = time.perf_counter() - t0
    assert rc == 0
    return dt


def cpu_baseline(mode, target_s=12.0):
    """Bounded sample of the same workload shape: calibrate, then ~target_s seconds of CPU work."""
    L = _fast_oracle()
    cores = L.pvo_bench_threads()
    frames = 64
    n_in = WINDOW + (frames - 1) * HOP
    x = host_streams(cores, n_in)
    dt = cpu_port_run(L, x, frames, cores, mode)
    rate = cores * frames / dt
    frames2 = FRAMES_PER_STREAM
    streams2 = int(max(cores, min(64 * cores, rate * target_s / frames2)))
    streams2 = (streams2 + cores - 1) // cores * cores
    n_in2 = WINDOW + (frames2 - 1) * HOP
    base = host_streams(min(streams2, 16), n_in2)
    x2 = np.tile(base, ((streams2 + len(base) - 1) // len(base), 1))[:streams2]
    dt2 = cpu_port_run(L, x2, frames2, cores, mode)
    value = streams2 * frames2 / dt2
    # the same port on ONE thread (SURVEY 8d asks for both): a few streams, ~3 s
    s1 = int(max(1, min(8, value / cores * 3.0 / frames2)))
    dt1 = cpu_port_run(L, x2[:s1], frames2, 1, mode)
    return {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{streams2} streams x {frames2} frames (window {WINDOW}, hop {HOP}, {mode}-mode f32 oracle "
                      f"port, {dt2:.1f} s wall, all {cores} host threads)",
            "audio_s_per_s": value * HOP / FS,
            "single_thread": {"value": s1 * frames2 / dt1, "unit": "frames/s",
                              "sample": f"{s1} streams x {frames2} frames, {dt1:.1f} s wall"}}


def reference_gpu_build():
    """The reference's OWN cuFFT pipeline (karnel/*.cu + phaseVocoder.cpp driven like src/main.cpp:204-297) on this
    GPU, at the HEADLINE window 2048 / hop 512.  Comparison point only.  The unmodified sources cannot launch windows
    > 512 (<<<1, 2N>>>, karnel/kernel.cu:337), so this is oracle/_ref/pv_ref_harness_patched: the same sources with only
    the launch geometry of kernel.cu:301,314,337,354,380,393,406,419 rewritten (oracle/ref_harness/Makefile)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "pv_ref_harness_patched")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/pv_ref_harness_patched not built (needs the reference checkout at build time)"}
    import tempfile
    from signals import multitone
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.f32"), os.path.join(td, "out.f32")
        multitone(400 * HOP + WINDOW, fs=FS, seed=0).tofile(fin)
        try:
            r = subprocess.run([exe, fin, fout, str(WINDOW), str(WINDOW // HOP)], capture_output=True, text=True, timeout=180)
            info = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:      # comparison point only: never fail the bench on it
            return {"unavailable": f"harness failed: {e}"}
    return {"value": info["frames_per_s"], "unit": "frames/s", "window": WINDOW, "hop": HOP,
            "frames": info["frames_synth"], "analysis_s": info["analysis_s"], "resynthesis_s": info["resynthesis_s"],
            "note": "reference kernels (launch geometry patched for windows > 512, nothing else) + cuFFT plan per frame + "
                    "managed-memory attach per call, 1 channel"}


def run_reference(args):
    """--impl reference: the reference has no CPU implementation of this path (src/phaseVocoder.cpp only
    launches CUDA), so the timed arm is the oracle PORT of its pipeline on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = _fast_oracle()
    cores = L.pvo_bench_threads()
    frames = 64
    x = host_streams(cores, WINDOW + (frames - 1) * HOP)
    mode = args.mode
    dt = cpu_port_run(L, x, frames, cores, mode)                # calibration (untimed)
    budget = 120.0 / max(1, args.steps + args.warmup)           # whole run within ~2 minutes
    frames = int(max(32, min(FRAMES_PER_STREAM, (cores * frames / dt) * min(budget, 8.0) / cores)))
    x = np.tile(host_streams(min(cores, 16), WINDOW + (frames - 1) * HOP), ((cores + 15) // 16, 1))[:cores]
    for _ in range(args.warmup):
        cpu_port_run(L, x, frames, cores, mode)
    times = [cpu_port_run(L, x, frames, cores, mode) for _ in range(args.steps)]
    total = sum(times)
    value = cores * frames * args.steps / total
    sample = f"{cores} streams x {frames} frames per step (bounded sample of the workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "audio_s_per_s": value * HOP / FS, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        # the same keys and values as our arm's `config` (the driver compares them)
        "config": {"workload": workload_name(mode, args.streams, args.frames), "window": WINDOW, "hop": HOP,
                   "mode": mode, "voices": 1, "streams_per_gpu": args.streams, "frames_per_stream": args.frames,
                   "l2": L2_NOTE, "sharding": SHARDING_NOTE},
        "note": "the reference has no CPU path (src/phaseVocoder.cpp only launches CUDA): this arm is the f32 oracle port of the "
                "same pipeline on all host threads, each step a bounded sample of the workload",
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln.split(", ") for (t, ln) in self.lines if t0 <= t <= t1 + 0.1] or [ln.split(", ") for _, ln in self.lines]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def device_streams(torch, n_streams, n_in, seed):
    """C4 generator on the device: 3 sines (f in [80, 8000] Hz, amp 0.1-0.3) + N(0, 1e-3) per stream."""
    g = torch.Generator(device="cuda")
    g.manual_seed(1234 + seed)
    t = torch.arange(n_in, device="cuda", dtype=torch.float32) / FS
    x = torch.empty((n_streams, n_in), device="cuda", dtype=torch.float32)
    chunk = 64
    for s0 in range(0, n_streams, chunk):
        s1 = min(n_streams, s0 + chunk)
        f = torch.rand((s1 - s0, 3, 1), device="cuda", generator=g) * (8000.0 - 80.0) + 80.0
        a = torch.rand((s1 - s0, 3, 1), device="cuda", generator=g) * 0.2 + 0.1
        ph = torch.rand((s1 - s0, 3, 1), device="cuda", generator=g) * 6.2831853
        xs = (a * torch.sin(6.2831853 * f * t[None, None, :] + ph)).sum(1)
        xs += torch.randn(xs.shape, device="cuda", generator=g) * 1e-3
        x[s0:s1] = xs
    return x


def bind_to_gpu_numa_node(index):
    """Runs this rank on the CPUs next to its GPU, so that the page-locked host buffers of the end-to-end
    leg (first touch) sit on the NUMA node the GPU's PCIe root hangs off.  Plain NVML; harmless if it fails."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception as e:          # noqa: BLE001 -- a placement hint, never a reason to fail the run
        print(f"[bench] no NUMA binding: {e}", file=sys.stderr)
        return None


def c5_file(torch, seconds=3600.0):
    """BASELINE.json config 5's input: a synthetic 48 kHz stereo file (three tones per channel + noise, seed 0), window 4096,
    hop 1024.  Returns (x [2, n] on the GPU, n_frames)."""
    N, H, fs = 4096, 1024, 48000.0
    nf = int(seconds * fs) // H
    n = N + (nf - 1) * H
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    t = torch.arange(n, device="cuda", dtype=torch.float64)
    x = torch.empty((2, n), device="cuda", dtype=torch.float32)
    for c in range(2):
        f = torch.rand(3, generator=g, device="cuda", dtype=torch.float64) * (8000.0 - 80.0) + 80.0
        a = torch.rand(3, generator=g, device="cuda", dtype=torch.float64) * 0.2 + 0.1
        acc = torch.zeros(n, device="cuda", dtype=torch.float32)
        for i in range(3):
            acc += (a[i] * torch.sin(6.283185307179586 * f[i] / fs * t)).float()
        x[c] = acc + torch.randn(n, generator=g, device="cuda") * 1e-3
    return x, nf


def c5_single(torch, pvb200, local, seconds=3600.0):
    """Config 5 on ONE GPU (the record `c5` of the N = 1 line): the two channels are cut into frame-range parts on the device
    (analysis pass that keeps {|X|, D} of every frame -> per-part phase carry -> processing from the stored analysis; compat:
    independent segments).  Device-resident, CUDA events, 3 calls after 2 warm-ups."""
    N, H = 4096, 1024
    x, nf = c5_file(torch, seconds)
    rec = {"workload": f"synthetic {seconds / 3600:g} h 48 kHz stereo file, window {N}, hop {H}: 2 x {nf} frames, one GPU"}
    for mode in ("corrected", "compat"):
        corr = mode == "corrected"
        pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, device=local, mode=pvb200.MODE_CORRECTED if corr else pvb200.MODE_COMPAT,
                                 window_type=pvb200.WIN_HANN_PERIODIC if corr else pvb200.WIN_HAMMING, pitch=(SEMITONES_7,))
        out = torch.empty((2, 1, nf * H), device="cuda")
        for _ in range(2):
            pv.process(x, nf, out=out)
        torch.cuda.synchronize()
        n0 = pv.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            pv.process(x, nf, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        rec[mode] = {"ms": ms, "frames_per_s": 2 * nf / (ms * 1e-3), "audio_s_per_s": 2 * seconds / (ms * 1e-3),
                     "launches_per_call": (pv.launch_count() - n0) // 3, "checksum": float(out[0, 0, :4096].double().abs().sum())}
        pv.close()
        del out
        torch.cuda.empty_cache()
    return rec


def c5_sharded(torch, dist, pvb200, world, rank, local, seconds=3600.0):
    """BASELINE.json config 5: a synthetic 1-hour 48 kHz stereo file, window 4096 / hop 1024 (2 x 168 750 frames), frame-range
    sharded over the ranks behind the C ABI (pv_shard_begin -> ONE NCCL all-gather of the carry records -> pv_shard_finish).
    Every rank hands the library only its own view (range + overlap-add halo + one frame).  Timing: CUDA events, max over
    ranks; rank 0 also runs the whole file alone for the speed-up and checks the gathered result bit for bit."""
    from pvb200 import sharding
    N, H = 4096, 1024
    x, nf = c5_file(torch, seconds)                   # the same file on every rank (seeded)
    n = x.shape[1]
    comm = sharding.TorchComm()
    rec = {"workload": f"synthetic {seconds / 3600:g} h 48 kHz stereo file, window {N}, hop {H}: 2 x {nf} frames, frame ranges over "
                       f"{world} GPUs", "exchange": "one all_gather of pv_shard_carry_elems() int64 per channel and rank (corrected); "
                                                    "none (compat: input halo recomputed)"}
    for mode in ("corrected", "compat"):
        corr = mode == "corrected"
        mk = lambda: pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, device=local, mode=pvb200.MODE_CORRECTED if corr else pvb200.MODE_COMPAT,
                                         window_type=pvb200.WIN_HANN_PERIODIC if corr else pvb200.WIN_HAMMING, pitch=(SEMITONES_7,))
        pv = mk()
        p = pv.shard_plan(nf, world, rank)
        first = max(0, p.ks - 1) if p.k1 > p.k0 else 0
        xr = x[:, first * H:min(n, (max(p.k1, 1) - 1) * H + N)].contiguous() if p.k1 > p.k0 else x[:, :N].contiguous()
        run = lambda: sharding.process_sharded_capi(pv, xr, first, nf, comm)[0]
        for _ in range(2):
            out = run()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            out = run()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        per = (nf + world - 1) // world
        pad = torch.zeros((2, 1, per * H), device="cuda")
        pad[:, :, :out.shape[2]] = out
        parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, parts, dst=0)
        r = {"ms": float(ms.item()), "frames_per_s": 2 * nf / (float(ms.item()) * 1e-3), "carry_bytes_per_rank": 2 * 8 * pv.shard_carry_elems() if corr else 0}
        if rank == 0:
            got = torch.cat(parts, 2)[:, :, :nf * H]
            del parts, pad
            one = mk()
            ref = one.process(x, nf)
            torch.cuda.synchronize()
            r["bit_identical_to_single_gpu"] = bool(torch.equal(got, ref))
            del got
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(reps):
                one.process(x, nf, out=ref)
            t1.record()
            torch.cuda.synchronize()
            r["single_gpu_ms"] = t0.elapsed_time(t1) / reps
            r["speedup_vs_single_gpu"] = r["single_gpu_ms"] / r["ms"]
            one.close()
            del ref
        pv.close()
        del out, xr
        torch.cuda.empty_cache()
        rec[mode] = r
    return rec


def run_ours(args):
    import torch
    import torch.distributed as dist

    import pvb200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    S, F = args.streams, args.frames
    n_in = WINDOW + (F - 1) * HOP
    corrected = args.mode == "corrected"

    def make_pv(is_corrected):
        return pvb200.PhaseVocoder(WINDOW, hop_in=HOP, hop_out=HOP, device=local,
                                   mode=pvb200.MODE_CORRECTED if is_corrected else pvb200.MODE_COMPAT,
                                   window_type=pvb200.WIN_HANN_PERIODIC if is_corrected else pvb200.WIN_HAMMING,
                                   pitch=(SEMITONES_7,))

    pv = make_pv(corrected)
    V = pv.n_voices
    x = device_streams(torch, S, n_in, rank)
    out = torch.empty((S, V, F * HOP), device="cuda", dtype=torch.float32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    for _ in range(max(3, args.warmup)):
        pv.process(x, F, out=out)
    barrier()
    pv.timing(True)
    pv.timing_read()
    l0 = pv.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        pv.process(x, F, out=out)
    e1.record()
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler else None
    launches = pv.launch_count() - l0
    kern_ms, kern_n = pv.timing_read()
    pv.timing(False)
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    frames_step = S * F * world
    value = frames_step * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host entry point of the C ABI ----
    e2e = None
    if not args.no_e2e:
        xh = torch.empty((S, n_in), dtype=torch.float32, pin_memory=True)
        xh.copy_(x)
        oh = torch.empty((S, V, F * HOP), dtype=torch.float32, pin_memory=True)
        for _ in range(2):
            pv.process_host(xh, F, out=oh)
        barrier()
        n_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            pv.process_host(xh, F, out=oh)           # synchronous: H2D + kernel + D2H
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": frames_step * n_e2e / float(dt.item()), "unit": "frames/s",
               "h2d_bytes_per_step": int(S * n_in * 4 * world), "d2h_bytes_per_step": int(S * V * F * HOP * 4 * world),
               "steps": n_e2e, "audio_s_per_s": frames_step * n_e2e / float(dt.item()) * HOP / FS,
               "api": "pv_process_host (C ABI, pinned host buffers)",
               "host_cpus_bound": (len(numa) if numa else None)}
        # keep a checksum so that the D2H result is actually consumed
        e2e["checksum"] = float(oh[0, 0, :4096].double().abs().sum())
        # the ceiling of this number on this box: the SAME bytes copied in both directions at once (two streams, whole
        # buffers, no kernel), on all ranks at the same time
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def copies():
            with torch.cuda.stream(s_in):
                x.copy_(xh, non_blocking=True)
            with torch.cuda.stream(s_out):
                oh.copy_(out, non_blocking=True)
        copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            copies()
        barrier()
        dtc = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
        ceil_v = frames_step * n_e2e / float(dtc.item())
        e2e["copy_ceiling"] = {"value": ceil_v, "unit": "frames/s", "frac_achieved": e2e["value"] / ceil_v,
                               "gb_per_s_each_direction_all_gpus": S * n_in * 4 * world * n_e2e / float(dtc.item()) / 1e9,
                               "method": "pinned host buffers of this run copied H2D and D2H concurrently (two streams, no kernel), all "
                                         "ranks at once, same barriers; what pv_process_host could reach if compute were free"}
        del oh
        # the same call with 16-bit PCM host buffers (what a WAV file holds): AudioFile's int16<->float rules run
        # on the device, half the bytes cross PCIe
        xi = torch.empty((S, n_in), dtype=torch.int16, pin_memory=True)
        xi.copy_((xh * 32767.0).to(torch.int16))
        oi = torch.empty((S, V, F * HOP), dtype=torch.int16, pin_memory=True)
        for _ in range(2):
            pv.process_host_pcm16(xi, F, out=oi)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            pv.process_host_pcm16(xi, F, out=oi)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # its own ceiling: the same int16 buffers copied both ways at once (half the bytes of the float leg)
        x16 = torch.empty((S, n_in), dtype=torch.int16, device="cuda")
        o16 = torch.empty((S, V, F * HOP), dtype=torch.int16, device="cuda")

        def copies16():
            with torch.cuda.stream(s_in):
                x16.copy_(xi, non_blocking=True)
            with torch.cuda.stream(s_out):
                oi.copy_(o16, non_blocking=True)
        copies16()
        barrier()
        t0c = time.perf_counter()
        for _ in range(n_e2e):
            copies16()
        barrier()
        dtc16 = torch.tensor([time.perf_counter() - t0c], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dtc16, op=dist.ReduceOp.MAX)
        ceil16 = frames_step * n_e2e / float(dtc16.item())
        del x16, o16
        e2e["pcm16"] = {"value": frames_step * n_e2e / float(dt.item()), "unit": "frames/s",
                        "copy_ceiling": {"value": ceil16, "unit": "frames/s", "frac_achieved": frames_step * n_e2e / float(dt.item()) / ceil16},
                        "h2d_bytes_per_step": int(S * n_in * 2 * world), "d2h_bytes_per_step": int(S * V * F * HOP * 2 * world),
                        "api": "pv_process_host_pcm16 (16-bit PCM host buffers, AudioFile conversions on the device)",
                        "checksum": float(oi[0, 0, :4096].double().abs().sum())}
        del xh, xi, oi

    # ---- the other mode, device-resident only (reported next to the headline, same workload) ----
    other = None
    if world == 1:
        pv2 = make_pv(not corrected)
        for _ in range(3):
            pv2.process(x, F, out=out)
        torch.cuda.synchronize()
        pv2.timing(True)
        pv2.timing_read()
        n2 = max(3, min(args.steps, 5))
        for _ in range(n2):
            pv2.process(x, F, out=out)
        torch.cuda.synchronize()
        ms2, k2 = pv2.timing_read()
        other = {"mode": "compat" if corrected else "corrected", "value": S * F * k2 / (ms2 * 1e-3), "unit": "frames/s",
                 "kernel_ms": ms2 / k2, "note": "compat = the reference's own arithmetic (golden-WAV pinned); "
                 "corrected = phase-unwrap/pitch pipeline the north star names (+7 semitones)"}
        pv2.close()

    # ---- BASELINE config 5: ONE long stereo file -- on one GPU (record `c5`), or frame-range sharded over the ranks (`c5_sharded`) ----
    c5 = None
    if not args.no_c5:
        del x, out
        torch.cuda.empty_cache()
        try:
            c5 = c5_sharded(torch, dist, pvb200, world, rank, local) if world > 1 else c5_single(torch, pvb200, local)
        except Exception as e:          # noqa: BLE001 -- a side record: never lose the headline line over it
            c5 = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        peak, peak_src, sm_max = peaks()
        bytes_per_frame = 4 * HOP + 4 * V * HOP
        # fused stream-kernel time per step (a step may issue several launches: ragged last wave, frame-range split)
        kern_s = (kern_ms * 1e-3 / args.steps) if kern_n else (ms_total * 1e-3 / args.steps)
        achieved = S * F * bytes_per_frame / kern_s / 1e9
        traffic = None          # dram__bytes_read + dram__bytes_write of one ncu --set full capture of this launch shape
        traffic_source = None
        issue = None            # warp-instruction issue slots (4 per SM and cycle): what bound the round-1 kernels
        for tf in ("r02_traffic.json", "r01_traffic.json"):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", tf)))[args.mode]
            except Exception:
                continue
            if tr["frames_per_launch"] != S * F:
                continue
            traffic = tr["bytes_per_launch"]
            traffic_source = (f"profiles/{tf}: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this "
                              "launch shape (a constant from that capture, NOT measured in this run)")
            wpf = tr.get("warp_instructions_per_frame")
            if wpf:
                sms = torch.cuda.get_device_properties(local).multi_processor_count
                peak_issue = sms * 4 * (sm_max or 1965.0) * 1e6
                issue = {"warp_instructions_per_frame": wpf, "achieved": S * F * wpf / kern_s / 1e9,
                         "peak": peak_issue / 1e9, "unit": "G warp-instructions/s",
                         "frac": S * F * wpf / kern_s / peak_issue,
                         "note": "instruction count from the ncu capture in profiles/, time measured live"}
            break
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_source, "algorithmic_bytes_per_launch": S * F * bytes_per_frame, "peak_source": peak_src, "kernel": "fused stream kernel",
                    "kernel_ms": kern_s * 1e3, "algorithmic_bytes_per_frame": bytes_per_frame,
                    "note": "fully fused, the path is bound by instruction issue / latency at 16-20 resident warps per SM, not by HBM "
                            "(SURVEY fact 5; DESIGN.md 4.4): see `issue`",
                    "issue": issue}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s",
            "audio_s_per_s": value * HOP / FS, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.mode, S, F), "window": WINDOW, "hop": HOP, "mode": args.mode,
                       "voices": V, "streams_per_gpu": S, "frames_per_stream": F,
                       "l2": L2_NOTE, "sharding": SHARDING_NOTE},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        line["other_mode"] = other
        if c5 is not None:
            line["c5_sharded" if world > 1 else "c5"] = c5
        if world == 1:
            line["reference_gpu_build"] = reference_gpu_build()
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.mode)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The one JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print to fd 1 behind Python's back (NCCL's version banner under NCCL_DEBUG=WARN, for one): keep the
    # real stdout for the JSON line alone and send everything else to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

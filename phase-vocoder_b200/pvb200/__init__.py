"""pvb200 -- thin ctypes binding of libpv_b200.so (C ABI: include/pv_b200.h).

Used by the tests and bench.py.  There is no Python or CPU implementation behind it: if the
CUDA library is missing, or no sm_100 device is present, the calls raise.
Tensors are torch CUDA tensors (device entry points) or numpy arrays (host entry points);
torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# PV_B200_LIB: another build of the same library (kernel A/B experiments, tools/ab_variants.sh); still no fallback
LIB_PATH = os.environ.get("PV_B200_LIB") or os.path.join(_DIR, "libpv_b200.so")

MODE_COMPAT, MODE_CORRECTED = 0, 1
WIN_HAMMING, WIN_HANN_SYM, WIN_HANN_PERIODIC = 0, 1, 2
FLAG_NAN_COMPAT = 1
CARRY_IN, CARRY_OUT, REUSE_AGGREGATE = 1, 2, 4
MAX_VOICES = 8

EXPORTS = [
    "pv_last_error", "pv_version", "pv_create", "pv_destroy", "pv_get_params", "pv_window_table",
    "pv_reference_schedule", "pv_analysis", "pv_resynthesis", "pv_test_overlap_add", "pv_analysis_batch",
    "pv_resynthesis_batch", "pv_state_bytes", "pv_process_device", "pv_process_device_ex", "pv_process_host",
    "pv_process_host_pcm16",
    "pv_process_host_pcm24",
    "pv_corrected_aggregate", "pv_corrected_state_from_carry", "pv_launch_count", "pv_timing_enable", "pv_timing_read",
    "pv_rt_open", "pv_rt_close", "pv_rt_reset", "pv_rt_latency_samples", "pv_rt_input", "pv_rt_output", "pv_rt_step",
    "pv_rt_callback", "pv_fft_batch", "pv_corrected_split_aggregate",
    "pv_shard_plan_frames", "pv_shard_carry_elems", "pv_shard_begin", "pv_shard_finish",
]


class PvError(RuntimeError):
    pass


class ShardPlanC(C.Structure):
    """pv_shard_plan (include/pv_b200.h)."""
    _fields_ = [("k0", C.c_int64), ("k1", C.c_int64), ("ks", C.c_int64), ("halo", C.c_int64)]


class Params(C.Structure):
    _fields_ = [("window", C.c_int32), ("hop_in", C.c_int32), ("hop_out", C.c_int32), ("mode", C.c_int32),
                ("window_type", C.c_int32), ("n_voices", C.c_int32), ("pitch", C.c_float * MAX_VOICES),
                ("flags", C.c_int32), ("device", C.c_int32)]


_lib = None


def load():
    """Loads libpv_b200.so; raises PvError if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PvError(f"{LIB_PATH} is missing: build it with `make -C {_DIR}` (or __graft_entry__.build()); "
                      "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.pv_last_error.restype = C.c_char_p
    L.pv_version.restype = C.c_char_p
    L.pv_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.pv_destroy.argtypes = [vp]
    L.pv_destroy.restype = None
    L.pv_get_params.argtypes = [vp, C.POINTER(Params)]
    L.pv_window_table.argtypes = [vp, vp]
    L.pv_reference_schedule.argtypes = [vp, i64, C.POINTER(i64), C.POINTER(i64)]
    L.pv_analysis.argtypes = [vp, vp, vp]
    L.pv_resynthesis.argtypes = [vp, vp, vp, vp]
    L.pv_test_overlap_add.argtypes = [vp, vp, vp, vp]
    L.pv_analysis_batch.argtypes = [vp, vp, i64, i64, vp, vp]
    L.pv_resynthesis_batch.argtypes = [vp, vp, i64, vp, vp, vp]
    L.pv_state_bytes.argtypes = [vp]
    L.pv_state_bytes.restype = C.c_size_t
    L.pv_process_device.argtypes = [vp, vp, i64, i64, i64, i64, i64, vp, i64, i64, vp, i32, vp]
    L.pv_process_host.argtypes = [vp, vp, i64, i64, i64, i64, i64, vp, i64, i64, vp, i32]
    L.pv_process_host_pcm16.argtypes = [vp, vp, i64, i64, i64, i64, i64, vp, i64, i64, vp, i32]
    L.pv_process_host_pcm24.argtypes = [vp, vp, i64, i64, i64, i64, i64, vp, i64, i64, vp, i32]
    L.pv_process_device_ex.argtypes = [vp, vp, i64, i64, i64, i64, i64, i64, vp, i64, i64, vp, i32, vp]
    L.pv_corrected_aggregate.argtypes = [vp, vp, i64, i64, i64, i64, vp, vp, vp, vp, vp]
    L.pv_corrected_state_from_carry.argtypes = [vp, i64, vp, vp, i64, vp, vp, vp]
    L.pv_launch_count.argtypes = [vp]
    L.pv_launch_count.restype = i64
    L.pv_timing_enable.argtypes = [vp, i32]
    L.pv_timing_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    L.pv_rt_open.argtypes = [vp, i64, i32, C.POINTER(vp)]
    L.pv_rt_close.argtypes = [vp]
    L.pv_rt_close.restype = None
    L.pv_rt_reset.argtypes = [vp]
    L.pv_rt_latency_samples.argtypes = [vp]
    L.pv_rt_latency_samples.restype = i64
    L.pv_rt_input.argtypes = [vp]
    L.pv_rt_input.restype = C.POINTER(C.c_float)
    L.pv_rt_output.argtypes = [vp]
    L.pv_rt_output.restype = C.POINTER(C.c_float)
    L.pv_rt_step.argtypes = [vp]
    L.pv_rt_callback.argtypes = [vp, vp, vp, C.c_uint32]
    L.pv_fft_batch.argtypes = [vp, vp, vp, i32, i64, i32, vp]
    L.pv_corrected_split_aggregate.argtypes = [vp, vp, i64, i64, i64, i64, i64, vp, i32, vp, vp]
    L.pv_shard_plan_frames.argtypes = [vp, i64, i32, i32, C.POINTER(ShardPlanC)]
    L.pv_shard_carry_elems.argtypes = [vp]
    L.pv_shard_carry_elems.restype = C.c_size_t
    L.pv_shard_begin.argtypes = [vp, vp, i64, i64, i64, i64, i64, i32, i32, vp, vp]
    L.pv_shard_finish.argtypes = [vp, vp, i64, i64, i64, i64, i64, i64, i32, i32, vp, vp, i64, i64, vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise PvError(f"pv_b200 error {rc}: {load().pv_last_error().decode()}")


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def _cuda_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


class PhaseVocoder:
    """Host-side mirror of the reference's `class PhaseVocoder` (src/phaseVocoder.h:9-139) over the
    C ABI.  `PhaseVocoder(samples, effect, scale, hop)` keeps the reference meaning: window =
    samples, hopSize = samples // hop (:79), outHopSize = int(scale * hopSize) (:104).  Explicit
    hop_in / hop_out override that derivation (the golden 1000hzout.wav needs Ha=1, Hs=128)."""

    def __init__(self, samples, effect="t", scale=1.0, hop=2, *, hop_in=None, hop_out=None, mode=MODE_COMPAT,
                 window_type=WIN_HAMMING, pitch=(1.0,), flags=0, device=-1):
        self.nSamps = int(samples)
        self.hopSize = int(samples) // int(hop) if hop_in is None else int(hop_in)
        self.timeScale = float(scale)
        self.outHopSize = int(float(scale) * self.hopSize) if hop_out is None else int(hop_out)
        self.effect = effect
        p = Params()
        p.window, p.hop_in, p.hop_out = self.nSamps, self.hopSize, self.outHopSize
        p.mode, p.window_type, p.flags, p.device = mode, window_type, flags, device
        p.n_voices = len(pitch) if mode == MODE_CORRECTED else 1
        for i, b in enumerate(pitch[:MAX_VOICES]):
            p.pitch[i] = float(b)
        self.mode, self.n_voices = mode, p.n_voices
        self._h = C.c_void_p()
        _check(load().pv_create(C.byref(p), C.byref(self._h)))
        self.imp = np.empty(self.nSamps, np.float32)        # the window table, as PhaseVocoder::imp
        _check(load().pv_window_table(self._h, self.imp.ctypes.data))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h and _lib is not None:
            _lib.pv_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # ---- reference granularity (device tensors) ----
    def analysis_CUFFT(self, input, output):
        """PhaseVocoder::analysis_CUFFT (src/phaseVocoder.cpp:25-33): input[N] -> output[2N,2]."""
        _check(load().pv_analysis(self._h, _ptr(input), _ptr(output)))

    analysis = analysis_CUFFT

    def resynthesis_CUFFT(self, backFrame, frontFrame, output):
        """PhaseVocoder::resynthesis_CUFFT (src/phaseVocoder.cpp:60-76)."""
        _check(load().pv_resynthesis(self._h, _ptr(backFrame), _ptr(frontFrame), _ptr(output)))

    resynthesis = resynthesis_CUFFT

    def test_overlap_add(self, input, back, output):
        _check(load().pv_test_overlap_add(self._h, _ptr(input), _ptr(back), _ptr(output)))

    # ---- batched ----
    def reference_schedule(self, num_samples):
        na, ns = C.c_int64(), C.c_int64()
        _check(load().pv_reference_schedule(self._h, num_samples, C.byref(na), C.byref(ns)))
        return na.value, ns.value

    def analysis_batch(self, x, n_frames, out=None):
        import torch
        N = self.nSamps
        if out is None:
            out = torch.empty((n_frames, 2 * N, 2), dtype=torch.float32, device=x.device)
        _check(load().pv_analysis_batch(self._h, _ptr(x), x.numel(), n_frames, _ptr(out), _cuda_stream()))
        return out

    def resynthesis_batch(self, spectra, back, out=None):
        import torch
        n_frames = spectra.shape[0]
        if out is None:
            out = torch.empty(n_frames * self.outHopSize, dtype=torch.float32, device=spectra.device)
        _check(load().pv_resynthesis_batch(self._h, _ptr(spectra), n_frames, _ptr(back), _ptr(out), _cuda_stream()))
        return out

    def state_bytes(self):
        return load().pv_state_bytes(self._h)

    def process(self, x, n_frames, n_analysed=None, out=None, state=None, flags=0, skip=0, n_in=None):
        """Fused hot path on device tensors.  x: [streams, n_in] float32 CUDA -> out [streams, V, (n_frames-skip)*Hs].
        `skip` leading (halo) frames are computed but not written; `n_in` overrides the number of valid
        samples per row (rows may overlap: frame-range parts of one long stream)."""
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
        S = x.shape[0]
        n_in = x.shape[1] if n_in is None else n_in
        n_out = (n_frames - skip) * self.outHopSize
        if out is None:
            out = torch.empty((S, self.n_voices, n_out), dtype=torch.float32, device=x.device)
        na = n_frames if n_analysed is None else n_analysed
        _check(load().pv_process_device_ex(self._h, _ptr(x), S, x.stride(0), n_in, na, n_frames, skip, _ptr(out),
                                           out.stride(0), out.stride(1), _ptr(state), flags, _cuda_stream()))
        return out

    # ---- corrected mode: phase carry for frame-range sharding ----
    def aggregate(self, x, n_frames, P_prev=None, n_in=None):
        """Analysis-only pass: (sumD int64 [S, nb], P_first uint32-as-int32 [S, nb], P_last [S, nb])."""
        import torch
        S = x.shape[0]
        n_in = x.shape[1] if n_in is None else n_in
        nb = self.nSamps // 2 + 1
        sumD = torch.zeros((S, nb), dtype=torch.int64, device=x.device)
        P_first = torch.zeros((S, nb), dtype=torch.int32, device=x.device)
        P_last = torch.zeros((S, nb), dtype=torch.int32, device=x.device)
        _check(load().pv_corrected_aggregate(self._h, _ptr(x), S, x.stride(0), n_in, n_frames, _ptr(P_prev), _ptr(sumD),
                                             _ptr(P_first), _ptr(P_last), _cuda_stream()))
        return sumD, P_first, P_last

    def unwrap_decisions(self, x, n_frames):
        """Diagnostic: the unwrapped phase difference D_k[b] (turns * 2^32, int32) the device computes for every frame
        of ONE stream x [n] -- through the public aggregate entry point, applied to the n_frames-1 overlapping two-frame
        windows of the stream (rows of stride hop_in; the aggregate of frames {k-1, k} is D_k).  D_0 = 0.  The
        aggregate shares the processing kernels' forward transform, so these are the decisions process() takes
        (tests/aligned.py uses them for the decision-aligned parity check)."""
        import torch
        N, Ha = self.nSamps, self.hopSize
        need = (n_frames - 1) * Ha + N + Ha
        xp = torch.zeros(need + 8, dtype=torch.float32, device=x.device)
        xp[:min(need, x.numel())] = x.reshape(-1)[:need]
        rows = torch.as_strided(xp, (n_frames - 1, N + Ha), (Ha, 1))
        sumD, _, _ = self.aggregate(rows, 2)
        D = torch.zeros((n_frames, N // 2 + 1), dtype=torch.int32, device=x.device)
        D[1:] = sumD.to(torch.int32)
        return D

    # ---- frame-range sharding across ranks, behind the C ABI (pv_shard_begin / pv_shard_finish) ----
    def shard_plan(self, n_frames, world, rank):
        p = ShardPlanC()
        _check(load().pv_shard_plan_frames(self._h, n_frames, world, rank, C.byref(p)))
        return p

    def shard_carry_elems(self):
        return load().pv_shard_carry_elems(self._h)

    def shard_begin(self, x, first_frame, n_frames, world, rank, n_in=None):
        """Phase 1 of a sharded run: x [S, n] holds the stream(s) from sample first_frame*hop_in on.  Returns this
        rank's carry record [S, elems] int64 (device) for the all-gather."""
        import torch
        S = x.shape[0]
        n_in = x.shape[1] if n_in is None else n_in
        carry = torch.empty((S, self.shard_carry_elems()), dtype=torch.int64, device=x.device)
        _check(load().pv_shard_begin(self._h, _ptr(x), first_frame, S, x.stride(0), n_in, n_frames, world, rank, _ptr(carry),
                                     _cuda_stream()))
        return carry

    def shard_finish(self, x, first_frame, n_frames, world, rank, carry_all, n_analysed=None, n_in=None):
        """Phase 2: carry_all [world, S, elems] int64 (device) as gathered.  Returns out [S, V, (k1-k0)*Hs]."""
        import torch
        S = x.shape[0]
        n_in = x.shape[1] if n_in is None else n_in
        p = self.shard_plan(n_frames, world, rank)
        out = torch.empty((S, self.n_voices, max(0, p.k1 - p.k0) * self.outHopSize), dtype=torch.float32, device=x.device)
        na = n_frames if n_analysed is None else n_analysed
        _check(load().pv_shard_finish(self._h, _ptr(x), first_frame, S, x.stride(0), n_in, na, n_frames, world, rank,
                                      _ptr(carry_all), _ptr(out), out.stride(0) if S else 0, out.stride(1) if S else 0,
                                      _cuda_stream()))
        return out

    def split_aggregate(self, x, n_frames, state=None, skip=0, n_in=None):
        """The analysis pass of process(x, n_frames, state=state, flags=CARRY_IN if state is not None, skip=skip) on
        its own: sumD int64 [S, nb] over all frames of that call.  Follow it with that process call and
        flags | REUSE_AGGREGATE: the per-part sums kept in the handle are reused instead of recomputed."""
        import torch
        S = x.shape[0]
        n_in = x.shape[1] if n_in is None else n_in
        sumD = torch.zeros((S, self.nSamps // 2 + 1), dtype=torch.int64, device=x.device)
        _check(load().pv_corrected_split_aggregate(self._h, _ptr(x), S, x.stride(0), n_in, n_frames, skip, _ptr(state),
                                                   CARRY_IN if state is not None else 0, _ptr(sumD), _cuda_stream()))
        return sumD

    def state_from_carry(self, P_first, sumD, n_before, P_prev):
        """Carried state [S, state_bytes] (uint8) at frame boundary n_before from a phase carry."""
        import torch
        S = sumD.shape[0]
        st = torch.zeros((S, self.state_bytes()), dtype=torch.uint8, device=sumD.device)
        _check(load().pv_corrected_state_from_carry(self._h, S, _ptr(P_first), _ptr(sumD), n_before, _ptr(P_prev),
                                                    _ptr(st), _cuda_stream()))
        return st

    def process_host(self, x, n_frames, n_analysed=None, out=None, state=None, flags=0):
        """Same through host buffers (numpy arrays or pinned CPU torch tensors): H2D + kernel + D2H."""
        S, n_in = x.shape
        n_out = n_frames * self.outHopSize
        if out is None:
            out = np.empty((S, self.n_voices, n_out), np.float32)
        na = n_frames if n_analysed is None else n_analysed
        if isinstance(x, np.ndarray):
            in_stride, os_, ov = x.strides[0] // 4, out.strides[0] // 4, out.strides[1] // 4
        else:
            in_stride, os_, ov = x.stride(0), out.stride(0), out.stride(1)
        _check(load().pv_process_host(self._h, _ptr(x), S, in_stride, n_in, na, n_frames, _ptr(out), os_, ov,
                                      _ptr(state), flags))
        return out

    def process_host_pcm16(self, x, n_frames, n_analysed=None, out=None, state=None, flags=0):
        """16-bit PCM in and out (numpy int16 arrays or pinned CPU int16 tensors); AudioFile's conversions run on
        the device.  x: [streams, n_in] int16 -> out [streams, V, n_frames*Hs] int16."""
        S, n_in = x.shape
        n_out = n_frames * self.outHopSize
        if out is None:
            out = np.empty((S, self.n_voices, n_out), np.int16)
        na = n_frames if n_analysed is None else n_analysed
        if isinstance(x, np.ndarray):
            in_stride, os_, ov = x.strides[0] // 2, out.strides[0] // 2, out.strides[1] // 2
        else:
            in_stride, os_, ov = x.stride(0), out.stride(0), out.stride(1)
        _check(load().pv_process_host_pcm16(self._h, _ptr(x), S, in_stride, n_in, na, n_frames, _ptr(out), os_, ov,
                                            _ptr(state), flags))
        return out

    def process_host_pcm24(self, x, n_frames, n_analysed=None, out=None, state=None, flags=0):
        """Packed 24-bit PCM in and out: x uint8 [streams, n_in, 3] (little-endian bytes of each sample, numpy) ->
        out uint8 [streams, V, n_frames*Hs, 3]; AudioFile's 24-bit conversions run on the device."""
        S, n_in, three = x.shape
        assert three == 3 and x.dtype == np.uint8 and x.strides[1] == 3 and x.strides[2] == 1 and x.strides[0] % 3 == 0
        n_out = n_frames * self.outHopSize
        if out is None:
            out = np.empty((S, self.n_voices, n_out, 3), np.uint8)
        na = n_frames if n_analysed is None else n_analysed
        _check(load().pv_process_host_pcm24(self._h, _ptr(x), S, x.strides[0] // 3, n_in, na, n_frames, _ptr(out),
                                            out.strides[0] // 3, out.strides[1] // 3, _ptr(state), flags))
        return out

    def fft_batch(self, x, inverse=False, out=None):
        """Stand-alone batched complex FFT of a [batch, n] complex64 CUDA tensor (unnormalised both ways)."""
        import torch
        assert x.is_cuda and x.dtype == torch.complex64 and x.dim() == 2 and x.is_contiguous()
        if out is None:
            out = torch.empty_like(x)
        _check(load().pv_fft_batch(self._h, _ptr(x), _ptr(out), x.shape[1], x.shape[0], 1 if inverse else -1,
                                   _cuda_stream()))
        return out

    def launch_count(self):
        return load().pv_launch_count(self._h)

    def timing(self, on=True):
        _check(load().pv_timing_enable(self._h, 1 if on else 0))

    def timing_read(self):
        ms, n = C.c_double(), C.c_int64()
        _check(load().pv_timing_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


class RealtimeServer:
    """Block server for live streams: the reference's RtAudio callback contract (src/main.cpp:45-59) for
    n_streams channels at once.  `input` / `output` are numpy views of the server's page-locked staging."""

    def __init__(self, pv, n_streams, block_frames=1):
        self.pv = pv                      # keeps the handle alive
        self._rt = C.c_void_p()
        _check(load().pv_rt_open(pv._h, n_streams, block_frames, C.byref(self._rt)))
        self.n_streams, self.block_frames = n_streams, block_frames
        self.block_in = block_frames * pv.hopSize
        self.block_out = block_frames * pv.outHopSize
        self.latency = load().pv_rt_latency_samples(self._rt)
        self.input = np.ctypeslib.as_array(load().pv_rt_input(self._rt), shape=(n_streams, self.block_in))
        self.output = np.ctypeslib.as_array(load().pv_rt_output(self._rt), shape=(n_streams, pv.n_voices, self.block_out))

    def step(self):
        _check(load().pv_rt_step(self._rt))
        return self.output

    def callback(self, output_buffer, input_buffer):
        """out, in, nBufferFrames as the reference's callback; planar float32 numpy arrays."""
        _check(load().pv_rt_callback(self._rt, output_buffer.ctypes.data, input_buffer.ctypes.data, input_buffer.shape[-1]))
        return 0

    def reset(self):
        _check(load().pv_rt_reset(self._rt))

    def close(self):
        if self._rt:
            load().pv_rt_close(self._rt)
            self._rt = C.c_void_p()
            self.input = self.output = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

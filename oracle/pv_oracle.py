"""ctypes front-end of the CPU oracle (oracle/pv_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
compat mode is pinned by the reference's golden WAVs; corrected mode is PARITY UNPINNED.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

WIN_HAMMING, WIN_HANN_SYM, WIN_HANN_PERIODIC = 0, 1, 2
FLAG_NAN_COMPAT = 1
MAX_VOICES = 8


def build(fast: bool = False) -> str:
    """Compile the oracle with gcc (idempotent)."""
    target = "libpv_oracle_fast.so" if fast else "libpv_oracle.so"
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True)
    return os.path.join(_HERE, target)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpv_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < max(
            os.path.getmtime(os.path.join(_HERE, f))
            for f in ("pv_oracle.c", "pv_oracle_impl.inc", "pv_oracle.h")
        ):
            build()
        L = C.CDLL(path)
        fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int
        L.pvo_window.argtypes = [ip, ip, fp]
        L.pvo_reference_schedule.argtypes = [C.c_long, ip, ip, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.pvo_analysis_frame_f64.argtypes = [fp, fp, ip, ip, dp]
        L.pvo_analysis_frame_f32.argtypes = [fp, fp, ip, ip, fp]
        L.pvo_resynthesis_frame_f64.argtypes = [dp, dp, fp, ip, ip, dp]
        L.pvo_resynthesis_frame_f32.argtypes = [fp, fp, fp, ip, ip, fp]
        L.pvo_process_compat_f64.argtypes = [fp, C.c_long, ip, ip, ip, fp, C.c_long, C.c_long, C.c_long, ip, dp, dp]
        L.pvo_process_compat_f64.restype = ip
        L.pvo_process_compat_f32.argtypes = [fp, C.c_long, ip, ip, ip, fp, C.c_long, C.c_long, C.c_long, ip, fp, fp]
        L.pvo_process_compat_f32.restype = ip
        L.pvo_corrected_gain.argtypes = [fp, ip, ip]
        L.pvo_corrected_gain.restype = C.c_float
        L.pvo_phase_turns32.argtypes = [C.c_double, C.c_double]
        L.pvo_phase_turns32.restype = C.c_uint32
        L.pvo_fft_f64.argtypes = [dp, dp, ip, ip]
        L.pvo_corrected_tables.argtypes = [ip, ip, ip, C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_uint32)]
        L.pvo_process_corrected.argtypes = [fp, C.c_long, ip, ip, ip, fp, ip, dp, C.c_long,
                                            C.POINTER(_State), ip, dp, C.c_long]
        L.pvo_process_corrected.restype = ip
        L.pvo_process_corrected_traced.argtypes = [fp, C.c_long, ip, ip, ip, fp, ip, dp, C.c_long,
                                                   C.POINTER(_State), ip, dp, C.c_long, C.POINTER(_Trace)]
        L.pvo_process_corrected_traced.restype = ip
        L.pvo_corrected_aggregate.argtypes = [fp, C.c_long, ip, ip, fp, C.c_long, ip, C.POINTER(C.c_uint32), ip,
                                              C.POINTER(C.c_int64), C.POINTER(C.c_uint32)]
        L.pvo_corrected_aggregate.restype = ip
        _LIB = L
    return _LIB


class _State(C.Structure):
    _fields_ = [("have_prev", C.c_int32), ("P_prev", C.POINTER(C.c_uint32)),
                ("psi", C.POINTER(C.c_uint64)), ("tail", C.POINTER(C.c_double))]


class _Trace(C.Structure):
    _fields_ = [("D_out", C.POINTER(C.c_int32)), ("mag_out", C.POINTER(C.c_double)),
                ("unwrap_adjust", C.POINTER(C.c_int8))]


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def window(kind: int, N: int) -> np.ndarray:
    w = np.empty(N, np.float32)
    lib().pvo_window(kind, N, _f(w))
    return w


def reference_schedule(num_samples: int, Ha: int, Hs: int):
    na, ns = C.c_long(), C.c_long()
    lib().pvo_reference_schedule(num_samples, Ha, Hs, C.byref(na), C.byref(ns))
    return na.value, ns.value


def analysis_frame(x: np.ndarray, win: np.ndarray, flags: int = 0, precision: int = 64) -> np.ndarray:
    """-> [2N, 2] {mag, phase} (karnel/kernel.cu:299-348)."""
    N = len(win)
    x = np.ascontiguousarray(x, np.float32)
    assert len(x) == N
    if precision == 64:
        out = np.empty((2 * N, 2), np.float64)
        lib().pvo_analysis_frame_f64(_f(x), _f(win), N, flags, _d(out))
    else:
        out = np.empty((2 * N, 2), np.float32)
        lib().pvo_analysis_frame_f32(_f(x), _f(win), N, flags, _f(out))
    return out


def resynthesis_frame(back: np.ndarray, front: np.ndarray, win: np.ndarray, Hs: int, precision: int = 64) -> np.ndarray:
    """-> out[N] (karnel/kernel.cu:352-432)."""
    N = len(win)
    if precision == 64:
        back = np.ascontiguousarray(back, np.float64)
        front = np.ascontiguousarray(front, np.float64)
        out = np.empty(N, np.float64)
        lib().pvo_resynthesis_frame_f64(_d(back), _d(front), _f(win), N, Hs, _d(out))
    else:
        back = np.ascontiguousarray(back, np.float32)
        front = np.ascontiguousarray(front, np.float32)
        out = np.empty(N, np.float32)
        lib().pvo_resynthesis_frame_f32(_f(back), _f(front), _f(win), N, Hs, _f(out))
    return out


def process_compat(x: np.ndarray, N: int, Ha: int, Hs: int, win: np.ndarray, n_analysed: int, n_frames: int,
                   first_frame: int = 0, back: np.ndarray | None = None, flags: int = 0, precision: int = 64):
    """Whole-stream compat pipeline (src/main.cpp:228-297). Returns (out[n_frames*Hs], back[N])."""
    x = np.ascontiguousarray(x, np.float32)
    dt = np.float64 if precision == 64 else np.float32
    back = np.zeros(N, dt) if back is None else np.array(back, dt)
    out = np.zeros(n_frames * Hs, dt)
    fn = lib().pvo_process_compat_f64 if precision == 64 else lib().pvo_process_compat_f32
    cv = _d if precision == 64 else _f
    rc = fn(_f(x), len(x), N, Ha, Hs, _f(win), n_analysed, first_frame, n_frames, flags, cv(back), cv(out))
    if rc != 0:
        raise ValueError("pvo_process_compat: bad parameters")
    return out, back


class CorrectedState:
    """Stream state of the corrected mode: previous analysis phase, synthesis phase
    accumulators (turns*2^64) and the OLA accumulator per voice."""

    def __init__(self, N: int, V: int):
        nb = N // 2 + 1
        self.N, self.V = N, V
        self.have_prev = 0
        self.P_prev = np.zeros(nb, np.uint32)
        self.psi = np.zeros((V, nb), np.uint64)
        self.tail = np.zeros((V, N), np.float64)

    def copy(self):
        c = CorrectedState(self.N, self.V)
        c.have_prev = self.have_prev
        c.P_prev, c.psi, c.tail = self.P_prev.copy(), self.psi.copy(), self.tail.copy()
        return c


def process_corrected(x: np.ndarray, N: int, Ha: int, Hs: int, win: np.ndarray, betas, n_frames: int,
                      state: CorrectedState | None = None, precision: int = 64):
    """Corrected mode (specification). Returns (out[V, n_frames*Hs], state)."""
    x = np.ascontiguousarray(x, np.float32)
    betas = np.ascontiguousarray(betas, np.float64)
    V = len(betas)
    st = CorrectedState(N, V) if state is None else state
    out = np.zeros((V, n_frames * Hs), np.float64)
    cs = _State(st.have_prev, st.P_prev.ctypes.data_as(C.POINTER(C.c_uint32)),
                st.psi.ctypes.data_as(C.POINTER(C.c_uint64)), _d(st.tail))
    rc = lib().pvo_process_corrected(_f(x), len(x), N, Ha, Hs, _f(win), V, _d(betas), n_frames,
                                     C.byref(cs), precision, _d(out), n_frames * Hs)
    if rc != 0:
        raise ValueError("pvo_process_corrected: bad parameters")
    st.have_prev = cs.have_prev
    return out, st


def process_corrected_traced(x, N, Ha, Hs, win, betas, n_frames, unwrap_adjust=None, precision=64):
    """Corrected mode with the decision trace (pv_oracle.h, pvo_corrected_trace).
    Returns (out[V, n_frames*Hs], D[n_frames, N/2+1] int32, mag[n_frames, N/2+1]); unwrap_adjust (int8, same
    shape as D) moves individual unwrap decisions by whole turns."""
    x = np.ascontiguousarray(x, np.float32)
    betas = np.ascontiguousarray(betas, np.float64)
    V, nb = len(betas), N // 2 + 1
    st = CorrectedState(N, V)
    out = np.zeros((V, n_frames * Hs), np.float64)
    D = np.zeros((n_frames, nb), np.int32)
    mag = np.zeros((n_frames, nb), np.float64)
    adj = None if unwrap_adjust is None else np.ascontiguousarray(unwrap_adjust, np.int8)
    assert adj is None or adj.shape == D.shape
    tr = _Trace(D.ctypes.data_as(C.POINTER(C.c_int32)), _d(mag),
                None if adj is None else adj.ctypes.data_as(C.POINTER(C.c_int8)))
    cs = _State(0, st.P_prev.ctypes.data_as(C.POINTER(C.c_uint32)), st.psi.ctypes.data_as(C.POINTER(C.c_uint64)), _d(st.tail))
    rc = lib().pvo_process_corrected_traced(_f(x), len(x), N, Ha, Hs, _f(win), V, _d(betas), n_frames,
                                            C.byref(cs), precision, _d(out), n_frames * Hs, C.byref(tr))
    if rc != 0:
        raise ValueError("pvo_process_corrected_traced: bad parameters")
    return out, D, mag


def corrected_tables(N: int, Ha: int, Hs: int, beta: float):
    nb = N // 2 + 1
    bq, rq = C.c_uint64(), C.c_uint64()
    a_lo, a_hi = np.empty(nb, np.int32), np.empty(nb, np.int32)
    nomS, nomA = np.empty(nb, np.uint64), np.empty(nb, np.uint32)
    lib().pvo_corrected_tables(N, Ha, Hs, beta, C.byref(bq), C.byref(rq),
                               a_lo.ctypes.data_as(C.POINTER(C.c_int32)), a_hi.ctypes.data_as(C.POINTER(C.c_int32)),
                               nomS.ctypes.data_as(C.POINTER(C.c_uint64)), nomA.ctypes.data_as(C.POINTER(C.c_uint32)))
    return dict(beta_q=bq.value, Rq=rq.value, a_lo=a_lo, a_hi=a_hi, nomS=nomS, nomA=nomA)


def corrected_aggregate(x, N, Ha, win, n_frames, P_prev=None, precision=64):
    """Analysis-only segment aggregate: (sumD int64[N/2+1], P_last uint32[N/2+1])."""
    x = np.ascontiguousarray(x, np.float32)
    nb = N // 2 + 1
    sumD, P_last = np.zeros(nb, np.int64), np.zeros(nb, np.uint32)
    have = 0 if P_prev is None else 1
    pp = np.zeros(nb, np.uint32) if P_prev is None else np.ascontiguousarray(P_prev, np.uint32)
    lib().pvo_corrected_aggregate(_f(x), len(x), N, Ha, _f(win), n_frames, have,
                                  pp.ctypes.data_as(C.POINTER(C.c_uint32)), precision,
                                  sumD.ctypes.data_as(C.POINTER(C.c_int64)),
                                  P_last.ctypes.data_as(C.POINTER(C.c_uint32)))
    return sumD, P_last


def corrected_gain(win: np.ndarray, Hs: int) -> float:
    return float(lib().pvo_corrected_gain(_f(win), len(win), Hs))


def fft(z: np.ndarray, direction: int = -1) -> np.ndarray:
    re = np.ascontiguousarray(z.real, np.float64).copy()
    im = np.ascontiguousarray(z.imag, np.float64).copy()
    lib().pvo_fft_f64(_d(re), _d(im), len(re), direction)
    return re + 1j * im

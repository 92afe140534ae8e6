"""Executes the fused kernel's per-frame body on the CPU (tests/emul: one std::thread per CUDA
thread, std::barrier as __syncthreads) and checks it against the oracle.  This validates the
register-blocked index algebra of pv_fused_core.cuh without a GPU; the GPU parity tests run
the same header compiled by nvcc."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pv_oracle as po
from signals import multitone, snr_db

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul_fused.cpp")
LIB = os.path.join(HERE, "emul", "libemul_fused.so")
CSRC = os.path.join(os.path.dirname(HERE), "phase-vocoder_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("pv_fused_core.cuh", "pv_fused_corrected.cuh", "pv_fft_regs.cuh", "pv_fused_tables.h")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    L = C.CDLL(LIB)
    fp = C.POINTER(C.c_float)
    L.emul_compat.argtypes = [C.c_int, fp, C.c_long, C.c_int, C.c_int, fp, C.c_long, C.c_long, C.c_int, fp]
    return L


@pytest.mark.parametrize("N,Ha,Hs,nf", [(256, 64, 64, 24), (256, 128, 128, 12), (512, 128, 128, 10),
                                        (1024, 102, 512, 8), (2048, 512, 512, 7), (256, 1, 128, 6),
                                        (4096, 1024, 1024, 6), (4096, 1000, 2048, 4)])
def test_fused_body_matches_oracle(emul, N, Ha, Hs, nf):
    n_in = N + (nf - 1) * Ha - 3
    x = multitone(n_in, seed=N + Ha)
    win = po.window(po.WIN_HAMMING, N)
    out = np.zeros(nf * Hs, np.float32)
    fp = C.POINTER(C.c_float)
    rc = emul.emul_compat(int(np.log2(N)), x.ctypes.data_as(fp), n_in, Ha, Hs, win.ctypes.data_as(fp), nf - 1, nf, 0,
                          out.ctypes.data_as(fp))
    assert rc == 0
    want, _ = po.process_compat(x, N, Ha, Hs, win, nf - 1, nf)
    assert snr_db(want, out) > 100, snr_db(want, out)


@pytest.mark.parametrize("N,Ha,Hs,betas,nf", [
    (256, 64, 64, [1.0], 30), (256, 64, 64, [1.0, 2 ** (4 / 12), 2 ** (7 / 12), 2.0], 30),
    (512, 128, 128, [1.5], 16), (1024, 102, 512, [1.0], 12), (2048, 512, 512, [2 ** (7 / 12)], 9),
    (2048, 512, 512, [0.75, 1.0], 7), (4096, 1024, 1024, [2 ** (7 / 12)], 7), (4096, 1024, 512, [0.8, 1.0, 1.5], 5),
])
def test_corrected_body_matches_oracle(emul, N, Ha, Hs, betas, nf):
    emul.emul_corrected.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_long, C.c_int, C.c_int, C.POINTER(C.c_float),
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_float, C.c_long, C.POINTER(C.c_float), C.c_long]
    n_in = N + (nf - 1) * Ha - 3
    x = multitone(n_in, seed=N + Ha)
    win = po.window(po.WIN_HANN_PERIODIC, N)
    V = len(betas)
    tabs = [po.corrected_tables(N, Ha, Hs, b) for b in betas]
    a_lo = np.concatenate([t["a_lo"] for t in tabs]).astype(np.int32)
    a_hi = np.concatenate([t["a_hi"] for t in tabs]).astype(np.int32)
    nomS = np.concatenate([t["nomS"] for t in tabs]).astype(np.uint64)
    Rq = np.array([t["Rq"] for t in tabs], np.uint64)
    bq = np.array([t["beta_q"] for t in tabs], np.uint64)
    nomA = tabs[0]["nomA"]
    out = np.zeros((V, nf * Hs), np.float32)
    fp = C.POINTER(C.c_float)
    rc = emul.emul_corrected(int(np.log2(N)), x.ctypes.data_as(fp), n_in, Ha, Hs, win.ctypes.data_as(fp), V,
                             nomA.ctypes.data, a_lo.ctypes.data, a_hi.ctypes.data, nomS.ctypes.data, Rq.ctypes.data,
                             bq.ctypes.data, po.corrected_gain(win, Hs), nf, out.ctypes.data_as(fp), nf * Hs)
    assert rc == 0
    want, _ = po.process_corrected(x, N, Ha, Hs, win, betas, nf)
    for v in range(V):
        assert snr_db(want[v], out[v]) > 100, (v, snr_db(want[v], out[v]))
    # the stored-analysis split (analysis pass keeps {|X|, D} of every frame, processing pass synthesises from the stored rows:
    # frame_corrected MODE 2 then MODE 1, three barriers per frame) gives the same samples bit for bit
    out2 = np.zeros_like(out)
    emul.emul_stored_analysis(1)
    try:
        rc = emul.emul_corrected(int(np.log2(N)), x.ctypes.data_as(fp), n_in, Ha, Hs, win.ctypes.data_as(fp), V,
                                 nomA.ctypes.data, a_lo.ctypes.data, a_hi.ctypes.data, nomS.ctypes.data, Rq.ctypes.data,
                                 bq.ctypes.data, po.corrected_gain(win, Hs), nf, out2.ctypes.data_as(fp), nf * Hs)
    finally:
        emul.emul_stored_analysis(0)
    assert rc == 0 and np.array_equal(out, out2)

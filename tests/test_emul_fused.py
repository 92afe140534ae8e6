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
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("pv_fused_core.cuh", "pv_fft_regs.cuh", "pv_fused_tables.h")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    L = C.CDLL(LIB)
    fp = C.POINTER(C.c_float)
    L.emul_compat.argtypes = [C.c_int, fp, C.c_long, C.c_int, C.c_int, fp, C.c_long, C.c_long, C.c_int, fp]
    return L


@pytest.mark.parametrize("N,Ha,Hs,nf", [(256, 64, 64, 24), (256, 128, 128, 12), (512, 128, 128, 10),
                                        (1024, 102, 512, 8), (2048, 512, 512, 7), (256, 1, 128, 6)])
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

"""The C++ mirror of the reference's `class PhaseVocoder` (phase-vocoder_b200/host/phaseVocoder.h):
compiles with plain g++ against the C ABI; exits like checkCUDAError_ (src/io.cpp:115-124) when no
device is present (CPU) and processes audio on a GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "shim_smoke.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "shim_smoke")
LIBDIR = os.path.join(ROOT, "phase-vocoder_b200")


def _build():
    if not os.path.exists(os.path.join(LIBDIR, "libpv_b200.so")):
        subprocess.run(["make", "-C", LIBDIR], check=True)
    subprocess.run(["g++", "-std=c++17", "-o", EXE, SRC, "-L" + LIBDIR, "-lpv_b200", "-Wl,-rpath," + LIBDIR], check=True)


def test_shim_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 1
    assert "Cuda error: PhaseVocoder" in r.stderr


@pytest.mark.gpu
def test_shim_processes_audio_on_gpu():
    _build()
    r = subprocess.run([EXE, "run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "shim ok" in r.stdout

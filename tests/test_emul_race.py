"""Race check of the fused kernel bodies without a GPU tool (SURVEY 5 asks for racecheck; compute-sanitizer is closed
on this GPU pool, see profiles/r02_sanitizer.md).  tests/emul runs pv_fused_core.cuh / pv_fused_corrected.cuh with one
std::thread per CUDA thread and std::barrier as the group barrier; under ThreadSanitizer every conflicting pair of
shared-buffer accesses that no barrier orders is a reported data race: the in-place exchange with its elided barriers,
the input ring refilled one frame ahead, the overlap-add ring with the deferred emit, the mag / D / psi arrays.
Self test: with ONE barrier dropped ThreadSanitizer must complain."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def tsan_exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("tsan")
    obj, exe = str(d / "pvo.o"), str(d / "tsan_emul")
    subprocess.run(["gcc", "-O1", "-std=c11", "-c", "-o", obj, os.path.join(ROOT, "oracle", "pv_oracle.c")], check=True)
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-pthread", "-o", exe,
                        os.path.join(HERE, "emul", "tsan_main.cpp"), obj, "-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer build not available here: " + r.stderr[-300:])
    return exe


def _run(exe, *args):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=600)
    return r.returncode, (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")


def test_fused_bodies_have_no_unordered_shared_accesses(tsan_exe):
    rc, warnings = _run(tsan_exe)                      # windows 256..2048, compat + corrected (2 voices)
    assert rc == 0 and warnings == 0


@pytest.mark.parametrize("log2n,barrier", [(8, 3), (8, 9), (11, 5), (11, 12)])
def test_a_dropped_barrier_is_reported(tsan_exe, log2n, barrier):
    rc, warnings = _run(tsan_exe, log2n, barrier)
    assert rc != 0 and warnings > 0

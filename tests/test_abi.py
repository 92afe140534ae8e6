"""CPU-side checks of the drop-in boundary: the C ABI library loads and exports every symbol
include/pv_b200.h declares (no compute calls: there is no GPU here and no CPU fallback)."""
import os
import re
import subprocess

import pytest

import pvb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pv_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(pvb200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = pvb200.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pv_b200.h but not exported"
    assert sorted(pvb200.EXPORTS) == names


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pvb200.PvError):
        pvb200.PhaseVocoder(256)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under phase-vocoder_b200/ or include/ may include,
    import, load or link it (comments that cite it as the specification are fine)."""
    bad = re.compile(r'#\s*include\s*[<"][^>"]*oracle|import\s+pv_oracle|from\s+pv_oracle|import\s+wav_oracle|'
                     r'libpv_oracle|CDLL\([^)]*oracle|-lpv_oracle')
    for base in ("phase-vocoder_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".so", ".o", ".log", ".pyc")):
                    continue
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(txt), os.path.join(dp, f)
    # and the built library has no undefined pvo_* symbols
    out = subprocess.run(["nm", "-D", pvb200.LIB_PATH], capture_output=True, text=True).stdout
    assert "pvo_" not in out


def test_null_arguments_are_rejected_before_any_device_work():
    """Every entry point checks its handle / pointers first and reports through pv_last_error (no CUDA call is made
    for these, so this runs without a GPU)."""
    import ctypes as C
    lib = pvb200.load()
    vp = C.c_void_p
    buf = (C.c_float * 8)()
    rt = vp()
    calls = [
        lambda: lib.pv_process_device(None, buf, 1, 8, 8, 1, 1, buf, 8, 8, None, 0, None),
        lambda: lib.pv_process_device_ex(None, buf, 1, 8, 8, 1, 1, 0, buf, 8, 8, None, 0, None),
        lambda: lib.pv_process_host(None, buf, 1, 8, 8, 1, 1, buf, 8, 8, None, 0),
        lambda: lib.pv_process_host_pcm16(None, buf, 1, 8, 8, 1, 1, buf, 8, 8, None, 0),
        lambda: lib.pv_process_host_pcm24(None, buf, 1, 8, 8, 1, 1, buf, 8, 8, None, 0),
        lambda: lib.pv_corrected_aggregate(None, buf, 1, 8, 8, 1, None, buf, None, None, None),
        lambda: lib.pv_corrected_split_aggregate(None, buf, 1, 8, 8, 1, 0, None, 0, buf, None),
        lambda: lib.pv_corrected_state_from_carry(None, 1, buf, buf, 0, buf, None, None),
        lambda: lib.pv_fft_batch(None, buf, buf, 4, 1, -1, None),
        lambda: lib.pv_shard_plan_frames(None, 10, 2, 0, C.byref(pvb200.ShardPlanC())),
        lambda: lib.pv_shard_begin(None, buf, 0, 1, 8, 8, 4, 2, 0, buf, None),
        lambda: lib.pv_shard_finish(None, buf, 0, 1, 8, 8, 4, 4, 2, 0, buf, buf, 8, 8, None),
        lambda: lib.pv_rt_open(None, 1, 1, C.byref(rt)),
        lambda: lib.pv_rt_step(None),
        lambda: lib.pv_rt_reset(None),
        lambda: lib.pv_rt_callback(None, buf, buf, 8),
        lambda: lib.pv_timing_enable(None, 1),
        lambda: lib.pv_create(None, C.byref(rt)),
    ]
    for i, call in enumerate(calls):
        assert call() != 0, f"call {i} accepted a null argument"
        assert lib.pv_last_error()
    assert lib.pv_rt_latency_samples(None) == 0 and not lib.pv_rt_input(None) and not lib.pv_rt_output(None)
    lib.pv_rt_close(None)
    lib.pv_destroy(None)
    assert lib.pv_launch_count(None) == 0 and lib.pv_state_bytes(None) == 0 and lib.pv_shard_carry_elems(None) == 0


def test_public_header_is_plain_c():
    """include/pv_b200.h is the FFI surface: it must compile as C99 and as C++11 on its own."""
    hdr = os.path.join(ROOT, "include", "pv_b200.h")
    for args in (["gcc", "-x", "c", "-std=c99"], ["g++", "-x", "c++", "-std=c++11"]):
        r = subprocess.run(args + ["-fsyntax-only", "-Wall", "-Wextra", "-Werror", hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr

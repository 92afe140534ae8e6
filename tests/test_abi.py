"""CPU-side checks of the drop-in boundary: the C ABI library loads and exports every symbol
include/pv_b200.h declares (no compute calls: there is no GPU here and no CPU fallback)."""
import os
import re

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
    """The oracle is test infrastructure: nothing under phase-vocoder_b200/ or include/ may name it."""
    for base in ("phase-vocoder_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".so", ".o", ".log", ".pyc")):
                    continue
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pv_oracle" not in txt or "oracle/pv_oracle.h" in txt and f.endswith(".cu"), f

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "phase-vocoder_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_compat.npz"))


@pytest.fixture(scope="session")
def golden_wav():
    """Raw-integer slices of testtones/440sine.wav (C1) and testtones/MAT_ZO_24_bit.wav (C2), decoded with the
    reference's AudioFile rules (src/AudioFile.h:1038-1042 s/32768; :508-518 s/8388608)."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_wav.npz"))
    return {"c1": (g["c1_440sine"].astype(np.float32) / np.float32(32768.0)),
            "c2": (g["c2_matzo"].astype(np.float32) / np.float32(8388608.0)),
            "c1_raw": g["c1_440sine"], "c2_raw": g["c2_matzo"], "c2_offset": int(g["c2_offset"]),
            "c1_num_samples": int(g["c1_num_samples"]), "c2_num_samples": int(g["c2_num_samples"]),
            "sha256": [str(h) for h in g["sha256"]]}


@pytest.fixture(scope="session")
def reference_dir():
    if not os.path.isdir(REFERENCE):
        pytest.skip("reference checkout not present (GPU box)")
    return REFERENCE

"""Pins the oracle -- and through it the CUDA path -- to the reference's OWN code running on the GPU.

oracle/_ref/pv_ref_harness is the unmodified reference pipeline (karnel/kernel.cu, hpfft.cu, common.cu,
src/phaseVocoder.cpp, src/io.cpp compiled from /root/reference by oracle/ref_harness/Makefile) driven
like src/main.cpp:204-297.  Unlike the 16-bit golden WAVs this compares float buffers, including the
intermediate {mag, phase} analysis output that no committed fixture pins."""
import json
import os
import subprocess

import numpy as np
import pytest

import pv_oracle as po
from signals import multitone, snr_db

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "pv_ref_harness")
# windows > 512: the same sources with ONLY the launch geometry of karnel/kernel.cu:301,314,337,354,380,393,406,419
# rewritten to blocks of <= 256 threads (oracle/ref_harness/Makefile: a generated, git-ignored copy)
EXE_PATCHED = os.path.join(ROOT, "oracle", "_ref", "pv_ref_harness_patched")


@pytest.fixture(scope="module")
def harness():
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/pv_ref_harness not built (needs /root/reference at build time)")
    return EXE


@pytest.mark.parametrize("N,hopdiv", [(256, 2), (256, 4), (512, 4), (128, 2),
                                      (256, -4), (1024, -4), (2048, -4), (2048, -2), (4096, -4)])
def test_reference_build_matches_oracle_and_cuda_path(harness, tmp_path, N, hopdiv):
    """hopdiv < 0: the launch-patched build (the only one that can run windows > 512; window 256 runs on both
    builds to show that the patch changes nothing).  (2048, 4) is BASELINE.json's headline shape / C2."""
    import pvb200
    if hopdiv < 0:
        hopdiv = -hopdiv
        harness = EXE_PATCHED
        if not os.path.exists(harness):
            pytest.skip("oracle/_ref/pv_ref_harness_patched not built (needs /root/reference at build time)")
    H = N // hopdiv
    nf = 120 if N <= 512 else 40
    n = nf * H + 17                                    # ragged end: last frames read past the input
    x = multitone(n, seed=N + hopdiv)
    fin, fout, fspec = tmp_path / "in.f32", tmp_path / "out.f32", tmp_path / "spec.f32"
    x.tofile(fin)
    r = subprocess.run([harness, str(fin), str(fout), str(N), str(hopdiv), str(fspec), "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    ref = np.fromfile(fout, np.float32)
    spec = np.fromfile(fspec, np.float32).reshape(8, 2 * N, 2)
    win = po.window(po.WIN_HAMMING, N)
    nA, nS = po.reference_schedule(n, H, H)
    assert (info["frames_analysed"], info["frames_synth"]) == (nA, nS)
    # 1. oracle (fp64) vs the reference's float output
    want, _ = po.process_compat(x, N, H, H, win, nA, nS)
    assert snr_db(want, ref) > 100, snr_db(want, ref)
    # 2. the reference's intermediate {mag, phase} buffers vs the oracle's step D
    for k in range(8):
        o = po.analysis_frame(x[k * H:k * H + N], win)
        mmax = o[:, 0].max()
        assert np.abs(spec[k, :, 0] - o[:, 0]).max() <= 2e-6 * mmax * np.log2(2 * N)
        strong = o[:, 0] >= 1e-3 * mmax
        dph = np.abs(spec[k, :, 1] - o[:, 1])
        dph = np.minimum(dph, np.pi - dph)
        assert dph[strong].max() <= 1e-3
    # 3. our CUDA path vs the reference's float output, fused and per-frame entry points
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H)
    xd = torch.from_numpy(x).cuda()
    got = pv.process(xd[None, :], nS, n_analysed=nA).cpu().numpy()[0, 0]
    assert snr_db(ref, got) > 100, snr_db(ref, got)
    gspec = pv.analysis_batch(xd, 8).cpu().numpy()
    mmax = spec[:, :, 0].max()
    assert np.abs(gspec[:, :, 0] - spec[:, :, 0]).max() <= 4e-6 * mmax * np.log2(2 * N)

"""GPU parity tests of the compat-mode path, through the C ABI, against the oracle (fp64).

Stated fp32 tolerances (BASELINE north_star: per-bin magnitude/phase error and output SNR >= 100 dB):
  analysis  |d mag|   <= 2e-6 * max|mag| * log2(2N)      per bin
            |d phase| <= 1e-3 rad (mod pi, atanf range)   for bins with mag >= 1e-3 * max|mag|
  output    SNR >= 100 dB against the fp64 oracle; golden WAV slices +-1 LSB of 16 bit.
"""
import numpy as np
import pytest

import pv_oracle as po
import wav_oracle as wo
from signals import c3_multitone, multitone, snr_db

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pvb200  # noqa: E402

WT = {po.WIN_HAMMING: pvb200.WIN_HAMMING, po.WIN_HANN_SYM: pvb200.WIN_HANN_SYM,
      po.WIN_HANN_PERIODIC: pvb200.WIN_HANN_PERIODIC}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make(N, Ha, Hs, wt=po.WIN_HAMMING, **kw):
    return pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs, window_type=WT[wt], **kw)


def test_window_tables_match_oracle():
    for wt in WT:
        for N in (64, 256, 2048, 4096):
            pv = make(N, N // 4, N // 4, wt)
            assert np.array_equal(pv.imp, po.window(wt, N))
            assert pv.reference_schedule(441000) == po.reference_schedule(441000, N // 4, N // 4)


@pytest.mark.parametrize("N", [64, 256, 1024, 2048, 4096])
def test_analysis_frame_parity(N):
    pv = make(N, N // 4, N // 4)
    x = multitone(N, seed=N)
    want = po.analysis_frame(x, pv.imp)
    out = torch.full((2 * N, 2), 7.0, device="cuda")
    pv.analysis_CUFFT(dev(x), out)
    got = out.cpu().numpy().astype(np.float64)
    mmax = want[:, 0].max()
    assert np.abs(got[:, 0] - want[:, 0]).max() <= 2e-6 * mmax * np.log2(2 * N)
    strong = want[:, 0] >= 1e-3 * mmax
    dph = np.abs(got[:, 1] - want[:, 1])
    dph = np.minimum(dph, np.pi - dph)
    assert dph[strong].max() <= 1e-3


def test_analysis_batch_equals_frames_and_zero_fill():
    N, Ha = 256, 64
    x = multitone(N + 20 * Ha, seed=2)
    pv = make(N, Ha, Ha)
    nf = 30                                   # the last frames read past the end -> zeros
    got = pv.analysis_batch(dev(x), nf).cpu().numpy()
    xp = np.concatenate([x, np.zeros(nf * Ha + N, np.float32)])
    for k in (0, 7, 19, 22, 29):
        want = po.analysis_frame(xp[k * Ha:k * Ha + N], pv.imp)
        mmax = max(want[:, 0].max(), 1e-30)
        assert np.abs(got[k, :, 0] - want[:, 0]).max() <= 4e-5 * mmax


@pytest.mark.parametrize("N,Hs", [(256, 128), (256, 64), (2048, 512), (1024, 102)])
def test_resynthesis_frame_parity(N, Hs):
    pv = make(N, Hs, Hs)
    x = multitone(N, seed=N + Hs)
    spec = po.analysis_frame(x, pv.imp)
    back = np.random.default_rng(1).normal(size=N)
    want = po.resynthesis_frame(back, spec, pv.imp, Hs)
    out = torch.empty(N, device="cuda")
    pv.resynthesis_CUFFT(dev(back.astype(np.float32)), dev(spec.astype(np.float32)), out)
    assert snr_db(want, out.cpu().numpy()) > 100


def test_frame_loop_like_main_cpp(golden):
    """The reference's own two host loops (src/main.cpp:228-297) driven through the per-frame ABI
    reproduce the golden testout.wav head to +-1 LSB."""
    N, H = 256, 128
    x = golden["testout_head_in"].astype(np.float32) / np.float32(32768)
    want = golden["testout_head_out"].astype(np.int32)
    pv = make(N, H, H)
    xd = dev(x)
    nf = 32
    spectra = [torch.zeros((2 * N, 2), device="cuda") for _ in range(nf)]
    for k in range(nf):
        pv.analysis_CUFFT(xd[k * H:], spectra[k])
    back = torch.zeros(N, device="cuda")
    outs = []
    for k in range(nf):
        final = torch.empty(N, device="cuda")
        pv.resynthesis_CUFFT(back, spectra[k], final)
        back = final
        outs.append(final[:H].cpu().numpy())
    got = wo.float_to_s16(np.concatenate(outs)).astype(np.int32)
    assert np.abs(got - want[:nf * H]).max() <= 1


def test_test_overlap_add():
    N, H = 256, 128
    pv = make(N, H, H)
    rng = np.random.default_rng(0)
    x, back = rng.normal(size=N).astype(np.float32), rng.normal(size=N).astype(np.float32)
    out = torch.empty(N, device="cuda")
    pv.test_overlap_add(dev(x), dev(back), out)
    want = x * pv.imp * pv.imp
    want[:N - H] += back[H:]
    assert np.allclose(out.cpu().numpy(), want, rtol=1e-6, atol=1e-7)


def test_process_golden_testout(golden):
    N, H = 256, 128
    x = golden["testout_head_in"].astype(np.float32) / np.float32(32768)
    want = golden["testout_head_out"].astype(np.int32)
    pv = make(N, H, H)
    nf = len(want) // H
    out = pv.process(dev(x)[None, :], nf).cpu().numpy()[0, 0]
    assert np.abs(wo.float_to_s16(out).astype(np.int32) - want).max() <= 1


def test_process_golden_tail_zero_fill(golden):
    N, H = 256, 128
    x = golden["testout_tail_in"].astype(np.float32) / np.float32(32768)
    k0 = int(golden["testout_tail_first_frame"])
    nA, nS = po.reference_schedule(int(golden["num_samples"]), H, H)
    pv = make(N, H, H)
    out = pv.process(dev(x)[None, :], nS - k0, n_analysed=nA - k0).cpu().numpy()[0, 0]
    want = golden["testout_tail_out"].astype(np.int32)
    got = wo.float_to_s16(out).astype(np.int32)[H:]
    assert np.abs(got - want[:len(got)]).max() <= 1


def test_process_golden_sine1000(golden):
    N, H = 256, 128
    x = golden["sine1000_head_in"].astype(np.float32) / np.float32(32768)
    want = golden["sine1000_head_out"].astype(np.int32)
    pv = make(N, 1, H, po.WIN_HANN_SYM)
    nf = len(want) // H
    out = pv.process(dev(x)[None, :], nf).cpu().numpy()[0, 0]
    assert np.abs(wo.float_to_s16(out).astype(np.int32) - want).max() <= 1


CASES = [
    # N, Ha, Hs, window, n_frames
    (256, 64, 64, po.WIN_HAMMING, 300),          # C1 shape
    (2048, 512, 512, po.WIN_HAMMING, 120),       # C2 / headline shape
    (1024, 102, 512, po.WIN_HAMMING, 90),        # C3 (hop divisors 10 / 2)
    (1024, 10, 2, po.WIN_HAMMING, 700),          # C3 literal reading
    (4096, 1024, 1024, po.WIN_HANN_PERIODIC, 40),  # C5 shape
    (512, 256, 256, po.WIN_HANN_SYM, 50),
    (64, 16, 16, po.WIN_HAMMING, 64),
    (128, 64, 128, po.WIN_HAMMING, 33),          # no overlap at all (Hs == N)
]


@pytest.mark.parametrize("N,Ha,Hs,wt,nf", CASES)
def test_process_parity_vs_oracle(N, Ha, Hs, wt, nf):
    S = 3
    n_in = N + (nf - 1) * Ha - 5                  # ragged: the last frame reads 5 samples past the end
    x = np.stack([c3_multitone(n_in) if s == 0 else multitone(n_in, seed=100 + s) for s in range(S)])
    pv = make(N, Ha, Hs, wt)
    got = pv.process(dev(x), nf).cpu().numpy()
    for s in range(S):
        want, _ = po.process_compat(x[s], N, Ha, Hs, pv.imp, nf, nf)
        assert snr_db(want, got[s, 0]) > 100, (s, snr_db(want, got[s, 0]))


def test_process_segment_split_invariance_and_state():
    """One long stream is split into frame-range segments with recomputed halo; the result must equal
    the oracle's single pass, and carry-out/carry-in must chain two calls seamlessly."""
    N, H = 256, 64
    nf = 6000
    x = multitone(N + nf * H, seed=77)
    pv = make(N, H, H)
    xd = dev(x)[None, :]
    full = pv.process(xd, nf).cpu().numpy()[0, 0]
    want, back = po.process_compat(x, N, H, H, pv.imp, nf, nf)
    assert snr_db(want, full) > 100
    st = torch.zeros(pv.state_bytes() // 4, device="cuda")
    a = pv.process(xd, 2500, state=st, flags=pvb200.CARRY_OUT).cpu().numpy()[0, 0]
    b = pv.process(xd[:, 2500 * H:], nf - 2500, state=st, flags=pvb200.CARRY_IN | pvb200.CARRY_OUT).cpu().numpy()[0, 0]
    assert snr_db(want, np.concatenate([a, b])) > 100
    assert snr_db(back, st.cpu().numpy()) > 100


def test_process_host_equals_device():
    N, H, nf, S = 1024, 256, 64, 5
    x = np.stack([multitone(N + nf * H, seed=s) for s in range(S)])
    pv = make(N, H, H)
    d = pv.process(dev(x), nf).cpu().numpy()
    h = pv.process_host(x, nf)
    assert np.array_equal(d, h)


def test_unanalysed_and_silent_frames():
    N, H = 256, 128
    pv = make(N, H, H)
    x = np.zeros((1, N + 40 * H), np.float32)
    x[0, 2000:3000] = multitone(1000, seed=4)
    got = pv.process(dev(x), 40, n_analysed=30).cpu().numpy()[0, 0]
    want, _ = po.process_compat(x[0], N, H, H, pv.imp, 30, 40)
    assert np.isfinite(got).all()
    assert snr_db(want, got) > 100
    pvn = make(N, H, H, flags=pvb200.FLAG_NAN_COMPAT)
    gn = pvn.process(dev(x), 40, n_analysed=30).cpu().numpy()[0, 0]
    wn, _ = po.process_compat(x[0], N, H, H, pv.imp, 30, 40, flags=po.FLAG_NAN_COMPAT)
    assert np.array_equal(np.isnan(gn), np.isnan(wn)) and np.isnan(gn).any()


def test_bad_parameters_are_rejected():
    for kw in (dict(samples=100), dict(samples=8192), dict(samples=256, hop_out=512), dict(samples=256, hop_in=0)):
        with pytest.raises(pvb200.PvError):
            pvb200.PhaseVocoder(**kw)


@pytest.mark.parametrize("N,Ha,Hs,nf", [(2048, 512, 512, 64), (256, 64, 64, 200), (1024, 102, 512, 40)])
def test_generic_kernel_still_matches(monkeypatch, N, Ha, Hs, nf):
    """The shape-generic stream kernel (used for windows / hops the tuned kernels do not cover)."""
    monkeypatch.setenv("PV_FORCE_GENERIC", "1")
    x = multitone(N + nf * Ha, seed=5)
    pv = make(N, Ha, Hs)
    got = pv.process(dev(x)[None, :], nf).cpu().numpy()[0, 0]
    want, _ = po.process_compat(x, N, Ha, Hs, pv.imp, nf, nf)
    assert snr_db(want, got) > 100


def test_many_streams_and_odd_strides():
    """More streams than resident groups, unaligned row pitch (scalar load path) and odd frame counts."""
    N, H, nf, S = 256, 64, 37, 700
    n_in = N + (nf - 1) * H + 1                   # odd pitch: no 8-byte alignment for most rows
    rng = np.random.default_rng(0)
    x = (rng.normal(size=(S, n_in)) * 0.1).astype(np.float32)
    pv = make(N, H, H)
    got = pv.process(dev(x), nf).cpu().numpy()
    for s in (0, 1, 333, 699):
        want, _ = po.process_compat(x[s], N, H, H, pv.imp, nf, nf)
        assert snr_db(want, got[s, 0]) > 100


def test_process_host_pcm16_applies_audiofile_rules(golden):
    """16-bit PCM in/out with the conversions on the device == float path + AudioFile's rules on the host, and it
    reproduces the golden WAV head directly from the raw int16 samples (+-1 LSB)."""
    N, H = 256, 128
    xi = golden["testout_head_in"].astype(np.int16)
    want = golden["testout_head_out"].astype(np.int32)
    nf = len(want) // H
    pv = make(N, H, H)
    x2 = np.stack([xi, np.roll(xi, 7)])
    got = pv.process_host_pcm16(x2, nf)
    assert got.dtype == np.int16 and got.shape == (2, 1, nf * H)
    assert np.abs(got[0, 0].astype(np.int32) - want).max() <= 1
    ref = pv.process_host(x2.astype(np.float32) / np.float32(32768), nf)
    assert np.array_equal(got, wo.float_to_s16(ref))
    # odd row length (device rows are padded) and many streams (chunked pipeline)
    rng = np.random.default_rng(5)
    x3 = rng.integers(-20000, 20000, size=(37, N + 19 * H + 3)).astype(np.int16)
    g3 = pv.process_host_pcm16(x3, 20)
    r3 = pv.process_host(x3.astype(np.float32) / np.float32(32768), 20)
    assert np.array_equal(g3, wo.float_to_s16(r3))


@pytest.mark.parametrize("N,Ha,Hs,nf,na", [(1024, 256, 256, 101, 101), (256, 64, 128, 90, 70), (512, 100, 128, 77, 77)])
def test_host_pipeline_chunks_over_frames_bit_exact(monkeypatch, N, Ha, Hs, nf, na):
    """The host path pipelines chunks of frames, carrying the overlap-add tail on the device from chunk to chunk:
    same bits as one device call, for float and 16-bit PCM buffers, with the caller's carry state as well."""
    S = 6
    x = np.stack([multitone(N + nf * Ha + 5, seed=40 + s) for s in range(S)])
    pv = make(N, Ha, Hs)
    # the host path pads device rows to 4 samples (aligned loads): give the device call the same row pitch
    xp = np.pad(x, ((0, 0), (0, -x.shape[1] % 4)))
    d = pv.process(dev(xp), nf, n_analysed=na, n_in=x.shape[1]).cpu().numpy()
    monkeypatch.setenv("PV_HOST_CHUNKS", "6")
    assert np.array_equal(d, pv.process_host(x, nf, n_analysed=na))
    xi = (x * 20000).astype(np.int16)
    monkeypatch.setenv("PV_HOST_CHUNKS", "1")
    one = pv.process_host_pcm16(xi, nf, n_analysed=na)
    monkeypatch.setenv("PV_HOST_CHUNKS", "5")
    assert np.array_equal(one, pv.process_host_pcm16(xi, nf, n_analysed=na))
    # two host calls joined by the caller's state == one call
    k = 40
    st = np.zeros((S, pv.state_bytes()), np.uint8)
    a = pv.process_host(x, k, n_analysed=min(na, k), state=st, flags=pvb200.CARRY_OUT)
    b = pv.process_host(np.ascontiguousarray(x[:, k * Ha:]), nf - k, n_analysed=max(0, na - k), state=st,
                        flags=pvb200.CARRY_IN | pvb200.CARRY_OUT)
    assert np.array_equal(np.concatenate([a, b], axis=2), d)

"""GPU parity tests of the CORRECTED mode (phase-difference unwrap -> true frequency -> pitch/time
scaling -> fixed-point phase accumulation -> WOLA), through the C ABI, against the fp64 oracle.

The reference does not implement this stage (PARITY UNPINNED): the oracle is the
specification and is itself checked by analytic properties in tests/test_oracle_props.py.
Tolerance: output SNR >= 100 dB against the fp64 oracle in EVERY case.  The integer phase path is exact, so the
only differences are the fp32 FFTs, atan2 and sincos -- with one inherent exception (DESIGN.md
"conditioning of the phase unwrap"): the unwrap princarg(P_k - P_{k-1} - nomA) is discontinuous at +-1/2 turn, so
for a bin whose phase difference lies within its fp32 phase error of that boundary the two implementations may
pick neighbouring aliases, and for a fractional R = beta*Hs/Ha that bin's accumulator differs by R turns from then
on.  For integer R the direct comparison is held to 100 dB.  For fractional R the comparison is DECISION-ALIGNED
(tests/aligned.py): (1) every per-frame, per-bin phase difference of the device equals the oracle's modulo one turn
within the bin's fp32 uncertainty, (2) alias disagreements are rare among the bins that carry energy, (3) with
exactly those decisions moved in the oracle the output agrees to >= 100 dB.  tools/spec_conditioning.py shows why the
spec keeps the per-bin unwrap: peak-picking / phase locking has MORE discontinuous decisions, not fewer."""
import numpy as np
import pytest

import pv_oracle as po
from aligned import aligned_parity
from signals import multitone, snr_db

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pvb200  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make(N, Ha, Hs, betas, wt=pvb200.WIN_HANN_PERIODIC):
    return pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs, mode=pvb200.MODE_CORRECTED, window_type=wt, pitch=tuple(betas))


def f32(b):
    """pv_params.pitch is a float: hand the oracle the same float-rounded ratio (a 4e-8 relative
    difference in beta is a 1e-5 turn/frame drift at the top bins -- visible at 100 dB)."""
    return float(np.float32(b))


SEMI7 = f32(2 ** (7 / 12))
SEMI4 = f32(2 ** (4 / 12))
CASES = [
    # N, Ha, Hs, pitch ratios, frames, input noise floor
    (256, 64, 64, [1.5], 200, 0.0),                                    # C1: window 256 hop 64, pitch x1.5
    (2048, 512, 512, [SEMI7], 60, 0.0),                                # C2 / headline: +7 semitones
    (1024, 102, 512, [1.0], 60, 0.0),                                  # C3: time stretch, hop divisors 10 / 2
    (256, 64, 64, [1.0, SEMI4, SEMI7, 2.0], 150, 0.0),         # C4: 4-voice harmoniser
    (512, 128, 128, [0.75], 80, 0.0),
    (1024, 256, 256, [1.0, 1.5], 50, 0.0),
    (2048, 512, 256, [1.0], 40, 0.0),                                  # time compression
    (4096, 1024, 1024, [1.0, 2.0], 24, 1e-3),                          # C5 shape (generic kernel)
    (4096, 1024, 1024, [SEMI7], 24, 0.0),
    (128, 32, 32, [1.0, 1.5], 100, 0.0),                               # small window (generic kernel)
    (2048, 512, 512, [1.0, 2.0], 60, 1e-3),                            # integer R: noise floor is harmless
    (256, 64, 128, [1.0], 150, 1e-3),                                  # x2 stretch, R = 2
]


@pytest.mark.parametrize("N,Ha,Hs,betas,nf,noise", CASES)
def test_corrected_parity_vs_oracle(N, Ha, Hs, betas, nf, noise):
    S = 3
    n_in = N + (nf - 1) * Ha - 7
    x = np.stack([multitone(n_in, seed=50 + s, noise=noise) for s in range(S)])
    pv = make(N, Ha, Hs, betas)
    win = po.window(po.WIN_HANN_PERIODIC, N)
    assert np.array_equal(pv.imp, win)
    xd = dev(x)
    got = pv.process(xd, nf).cpu().numpy()
    assert got.shape == (S, len(betas), nf * Hs)
    integer_R = all(abs(b * Hs / Ha - round(b * Hs / Ha)) < 1e-9 for b in betas)
    for s in range(S):
        if integer_R:                                          # no alias can matter: direct comparison
            want, _ = po.process_corrected(x[s], N, Ha, Hs, win, betas, nf)
            for v in range(len(betas)):
                assert snr_db(want[v], got[s, v]) > 100, (s, v, snr_db(want[v], got[s, v]))
            continue
        D = pv.unwrap_decisions(xd[s], nf).cpu().numpy()
        r = aligned_parity(x[s], N, Ha, Hs, win, betas, nf, D, got[s])
        assert r["phase_ratio"] < 1.0, r                       # per-bin phase parity modulo one turn
        assert r["frac"] < 1e-3, r                             # flips are rare among the bins that carry energy
        assert min(r["aligned"]) > 100, (s, r)                 # and nothing else differs


@pytest.mark.parametrize("noise", [1e-3])
def test_corrected_fractional_ratio_disagrees_only_by_unwrap_flips(noise):
    """Fractional R: the GPU and the fp64 oracle may unwrap a weak bin to opposite sides of +-1/2 turn.
    The accumulators then differ by an exact multiple of R turns -- and by nothing else."""
    N, Ha, Hs, beta, nf = 2048, 512, 512, SEMI7, 60
    x = multitone(N + nf * Ha, seed=50, noise=noise)
    pv = make(N, Ha, Hs, [beta])
    win = po.window(po.WIN_HANN_PERIODIC, N)
    st = torch.zeros(pv.state_bytes(), dtype=torch.uint8, device="cuda")
    got = pv.process(dev(x)[None, :], nf, state=st, flags=pvb200.CARRY_OUT).cpu().numpy()[0, 0]
    want, ost = po.process_corrected(x, N, Ha, Hs, win, [beta], nf)
    assert snr_db(want[0], got) > 55
    nb = N // 2 + 1
    raw = st.cpu().numpy()
    off = 8 + ((nb * 4 + 7) // 8) * 8
    psi = raw[off:off + 8 * nb].view(np.uint64)
    t = po.corrected_tables(N, Ha, Hs, beta)
    ok = t["a_lo"] <= t["a_hi"]
    d = ((psi[ok] - ost.psi[0][ok]).astype(np.int64) / 2.0 ** 64)            # turns, in [-0.5, 0.5)
    R = t["Rq"] / 2.0 ** 32
    small = np.abs(d) < 1e-3
    assert small.mean() > 0.99
    for dv in d[~small]:                                                      # each outlier = m * R turns
        m = np.arange(-4, 5)
        m = m[m != 0]
        resid = np.abs((dv - m * R + 0.5) % 1.0 - 0.5)
        assert resid.min() < 1e-3, dv


def test_corrected_identity_reconstructs_input():
    N, H, nf = 2048, 512, 40
    x = multitone(N + nf * H, seed=1)
    pv = make(N, H, H, [1.0])
    got = pv.process(dev(x)[None, :], nf).cpu().numpy()[0, 0]
    assert snr_db(x[N:nf * H], got[N:nf * H]) > 100


def test_corrected_pitch_moves_a_sine_on_gpu():
    N, H, fs, f0, beta = 2048, 512, 44100.0, 1000.0, 1.5
    n = N + 80 * H
    x = (0.25 * np.sin(2 * np.pi * f0 * np.arange(n) / fs) + 1e-3 * np.random.default_rng(0).normal(size=n)).astype(np.float32)
    pv = make(N, H, H, [beta])
    got = pv.process(dev(x)[None, :], 80).cpu().numpy()[0, 0]
    seg = got[8 * H:72 * H]
    spec = np.abs(np.fft.rfft(seg * np.hanning(len(seg))))
    fpk = np.argmax(spec) * fs / len(seg)
    assert abs(fpk - beta * f0) < 0.004 * beta * f0


def test_corrected_state_carry_is_bit_exact():
    """Two chained calls with carry-out / carry-in equal one call bit for bit (the phase path is
    integer, the OLA order is fixed)."""
    N, Ha, Hs, nf = 512, 128, 128, 90
    betas = [1.0, f32(1.26)]
    x = multitone(N + nf * Ha, seed=9)
    pv = make(N, Ha, Hs, betas)
    xd = dev(x)[None, :]
    full = pv.process(xd, nf).cpu().numpy()
    st = torch.zeros(pv.state_bytes(), dtype=torch.uint8, device="cuda")
    a = pv.process(xd, 37, state=st, flags=pvb200.CARRY_OUT).cpu().numpy()
    b = pv.process(xd[:, 37 * Ha:], nf - 37, state=st, flags=pvb200.CARRY_IN | pvb200.CARRY_OUT).cpu().numpy()
    assert np.array_equal(np.concatenate([a, b], axis=2), full)
    # and the carried state matches the oracle's: integer parts exactly up to fp32 phase noise
    _, ost = po.process_corrected(x, N, Ha, Hs, po.window(po.WIN_HANN_PERIODIC, N), betas, nf)
    raw = st.cpu().numpy()
    nb = N // 2 + 1
    assert int(raw[:4].view(np.uint32)[0]) == 1
    P = raw[8:8 + 4 * nb].view(np.uint32)
    dP = (P.astype(np.int64) - ost.P_prev.astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    strong = np.ones(nb, bool)
    assert np.abs(dP[strong]).max() < 2 ** 32 * 1e-3        # < 1e-3 turns on every bin (noise floor present)


def test_four_voices_carry_and_launch_modes_are_bit_exact(monkeypatch):
    """Three or four voices run as launches of two voices each (calls of >= 16 frames) or as one launch (short calls, or
    PV_VOICES_ONE_LAUNCH): same bits either way, and chained calls through the carried state equal one call."""
    N, H, nf = 512, 128, 70
    betas = [1.0, f32(1.26), SEMI7, 2.0]
    x = dev(np.stack([multitone(N + nf * H, seed=40 + s, noise=1e-3) for s in range(3)]))
    pv = make(N, H, H, betas)
    full = pv.process(x, nf).cpu().numpy()
    st = torch.zeros((3, pv.state_bytes()), dtype=torch.uint8, device="cuda")
    parts, k = [], 0
    for n in (30, 5, 35):                               # 30 and 35 frames: voice pairs; 5 frames: one launch
        parts.append(pv.process(x[:, k * H:], n, state=st, flags=(pvb200.CARRY_IN if k else 0) | pvb200.CARRY_OUT).cpu().numpy())
        k += n
    assert np.array_equal(np.concatenate(parts, axis=2), full)
    monkeypatch.setenv("PV_VOICES_ONE_LAUNCH", "1")     # read once per process by the launcher: may already be latched
    win = po.window(po.WIN_HANN_PERIODIC, N)
    want, _ = po.process_corrected(x[1].cpu().numpy(), N, H, H, win, [1.0, 2.0], nf)
    assert snr_db(want[0], full[1, 0]) > 100 and snr_db(want[1], full[1, 3]) > 100


def test_corrected_many_streams_host_path():
    N, H, nf, S = 256, 64, 40, 300
    rng = np.random.default_rng(3)
    x = (rng.normal(size=(S, N + nf * H)) * 0.1).astype(np.float32)
    pv = make(N, H, H, [1.0, 1.5])
    d = pv.process(dev(x), nf).cpu().numpy()
    h = pv.process_host(x, nf)
    assert np.array_equal(d, h)
    win = po.window(po.WIN_HANN_PERIODIC, N)
    for s in (0, 151, 299):
        want, _ = po.process_corrected(x[s], N, H, H, win, [1.0, 1.5], nf)
        assert snr_db(want[0], d[s, 0]) > 100 and snr_db(want[1], d[s, 1]) > 100      # white noise: no bin near the floor


def test_generic_corrected_kernel_matches_fused(monkeypatch):
    """The shape-generic corrected kernel (windows outside 256..2048) against the tuned one, incl. state carry."""
    N, Ha, Hs, nf = 1024, 256, 256, 50
    betas = [1.0, f32(1.5)]
    x = multitone(N + nf * Ha, seed=12, noise=1e-3)      # a noise floor: the two kernels use different atan2 / sincos
    xd = dev(x)[None, :]
    fused = make(N, Ha, Hs, betas).process(xd, nf).cpu().numpy()
    monkeypatch.setenv("PV_FORCE_GENERIC", "1")
    pv = make(N, Ha, Hs, betas)
    gen = pv.process(xd, nf).cpu().numpy()
    assert snr_db(fused[0, 0], gen[0, 0]) > 100 and snr_db(fused[0, 1], gen[0, 1]) > 100
    st = torch.zeros(pv.state_bytes(), dtype=torch.uint8, device="cuda")
    a = pv.process(xd, 21, state=st, flags=pvb200.CARRY_OUT).cpu().numpy()
    b = pv.process(xd[:, 21 * Ha:], nf - 21, state=st, flags=pvb200.CARRY_IN | pvb200.CARRY_OUT).cpu().numpy()
    assert np.array_equal(np.concatenate([a, b], axis=2), gen)


@pytest.mark.parametrize("N,Ha,Hs,betas,nf,S", [(2048, 512, 512, [SEMI7], 3000, 1), (256, 64, 64, [1.0, SEMI4, 2.0], 6890, 2),
                                                (1024, 102, 512, [1.0], 4300, 1), (4096, 1024, 1024, [SEMI7], 2000, 2)])
def test_corrected_intra_gpu_split_is_bit_exact(monkeypatch, N, Ha, Hs, betas, nf, S):
    """Few long streams are cut into frame-range parts on ONE GPU (aggregate -> per-part state -> process);
    the result and the carried-out state are bit-identical to the sequential single-segment run."""
    x = torch.from_numpy(np.stack([multitone(N + nf * Ha, seed=60 + s, noise=1e-3) for s in range(S)])).cuda()
    pv = make(N, Ha, Hs, betas)
    st_a = torch.zeros((S, pv.state_bytes()), dtype=torch.uint8, device="cuda")
    split = pv.process(x, nf, state=st_a, flags=pvb200.CARRY_OUT).cpu().numpy()
    assert pv.launch_count() == 3                      # aggregate + state build + process
    monkeypatch.setenv("PV_NO_SPLIT", "1")
    pv2 = make(N, Ha, Hs, betas)
    st_b = torch.zeros((S, pv2.state_bytes()), dtype=torch.uint8, device="cuda")
    seq = pv2.process(x, nf, state=st_b, flags=pvb200.CARRY_OUT).cpu().numpy()
    assert pv2.launch_count() == 1
    assert np.array_equal(split, seq)
    assert torch.equal(st_a, st_b)


@pytest.mark.parametrize("N,Ha,Hs,betas,nf,S", [(512, 128, 128, [0.8, SEMI7], 4000, 1), (1024, 256, 256, [1.0], 3000, 2), (2048, 512, 384, [0.75], 2500, 2),
                                                (4096, 1024, 1024, [1.0, 0.9, SEMI7], 1500, 1)])
def test_stored_analysis_split_equals_recomputing_split(monkeypatch, N, Ha, Hs, betas, nf, S):
    """A frame-range split keeps {|X|, D} of every frame from its analysis pass and synthesises from them (PvAggArgs::md /
    PvProcessArgs::md: no second forward transform).  PV_NO_MD_STORE=1 makes the processing pass repeat the forward transform,
    PV_NO_SPLIT=1 runs the streams sequentially: all three bit-identical, output and carried state -- with pitch ratios < 1
    (multi-source bins), several voices and the voice-group launches."""
    x = torch.from_numpy(np.stack([multitone(N + nf * Ha, seed=90 + s, noise=1e-3) for s in range(S)])).cuda()

    def run():
        pv = make(N, Ha, Hs, betas)
        st = torch.zeros((S, pv.state_bytes()), dtype=torch.uint8, device="cuda")
        y = pv.process(x, nf, state=st, flags=pvb200.CARRY_OUT).cpu().numpy()
        return y, st.cpu().numpy(), pv.launch_count()

    stored = run()
    monkeypatch.setenv("PV_NO_MD_STORE", "1")
    recomputed = run()
    monkeypatch.setenv("PV_NO_SPLIT", "1")
    sequential = run()
    assert stored[2] >= 3 and recomputed[2] >= 3            # both were split
    assert np.array_equal(stored[0], recomputed[0]) and np.array_equal(stored[1], recomputed[1])
    assert np.array_equal(stored[0], sequential[0]) and np.array_equal(stored[1], sequential[1])


def test_corrected_split_with_carry_in_and_skip_is_bit_exact(monkeypatch):
    """The on-GPU split also starts from a carried-in state and honours skip_frames (what a frame-range rank
    of a multi-GPU run does); and the public aggregate is split the same way.  All bit-identical to sequential."""
    N, Ha, Hs, nf, k_cut, skip = 1024, 256, 256, 2600, 900, 3
    betas = [SEMI7, 1.0]
    x = torch.from_numpy(multitone(N + nf * Ha, seed=31, noise=1e-3)).cuda()[None, :]

    def run(pv):
        st = torch.zeros((1, pv.state_bytes()), dtype=torch.uint8, device="cuda")
        a = pv.process(x, k_cut, state=st, flags=pvb200.CARRY_OUT)
        b = pv.process(x[:, k_cut * Ha:], nf - k_cut, state=st, flags=pvb200.CARRY_IN | pvb200.CARRY_OUT, skip=skip)
        sumD, Pf, Pl = pv.aggregate(x, nf)
        return a.cpu().numpy(), b.cpu().numpy(), st.cpu().numpy(), sumD.cpu().numpy(), Pf.cpu().numpy(), Pl.cpu().numpy()

    got = run(make(N, Ha, Hs, betas))
    monkeypatch.setenv("PV_NO_SPLIT", "1")
    want = run(make(N, Ha, Hs, betas))
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert got[1].shape[2] == (nf - k_cut - skip) * Hs


@pytest.mark.parametrize("S,nf", [(5, 120), (1, 900)])
def test_host_pipeline_chunks_over_frames_bit_exact(monkeypatch, S, nf):
    """Chunks of frames with the phase accumulators and the OLA tail carried on the device: same bits as one
    device call (few long streams take the frame-range split inside every chunk)."""
    N, Ha, Hs = 512, 128, 128
    betas = [1.0, f32(1.26)]
    rng = np.random.default_rng(9)
    x = (rng.normal(size=(S, N + nf * Ha)) * 0.1).astype(np.float32)
    pv = make(N, Ha, Hs, betas)
    d = pv.process(dev(x), nf).cpu().numpy()
    # an all-zero state carried in is a fresh start (also through the frame-range split of few long streams)
    z = torch.zeros((S, pv.state_bytes() // 4), device="cuda")
    assert np.array_equal(d, pv.process(dev(x), nf, state=z, flags=pvb200.CARRY_IN | pvb200.CARRY_OUT).cpu().numpy())
    for chunks in ("4", "7"):
        monkeypatch.setenv("PV_HOST_CHUNKS", chunks)
        assert np.array_equal(d, pv.process_host(x, nf))
    xi = (x * 20000).astype(np.int16)
    monkeypatch.setenv("PV_HOST_CHUNKS", "1")
    one = pv.process_host_pcm16(xi, nf)
    monkeypatch.setenv("PV_HOST_CHUNKS", "6")
    assert np.array_equal(one, pv.process_host_pcm16(xi, nf))
    k = 50
    st = np.zeros((S, pv.state_bytes()), np.uint8)
    a = pv.process_host(x, k, state=st, flags=pvb200.CARRY_OUT)
    b = pv.process_host(np.ascontiguousarray(x[:, k * Ha:]), nf - k, state=st, flags=pvb200.CARRY_IN | pvb200.CARRY_OUT)
    assert np.array_equal(np.concatenate([a, b], axis=2), d)


@pytest.mark.parametrize("N,Ha,betas,force", [(128, 32, [1.5], False), (4096, 1024, [SEMI7], False),
                                              (4096, 1024, [1.0, 1.26, 1.5], False), (1024, 256, [1.5], True),
                                              (1024, 256, [1.5], False)])
def test_carry_in_only_leaves_the_state_untouched(monkeypatch, N, Ha, betas, force):
    """include/pv_b200.h: the state is written only under PV_PROCESS_CARRY_OUT.  ADVICE r01: the generic kernels
    (windows 64/128/4096, PV_FORCE_GENERIC) used to update a carried-in state in place; now every kernel family
    lets a caller re-run from a saved state and get the same samples."""
    if force:
        monkeypatch.setenv("PV_FORCE_GENERIC", "1")
    monkeypatch.setenv("PV_NO_SPLIT", "1")
    nf, k = 30, 11
    betas = [f32(b) for b in betas]
    x = dev(np.stack([multitone(N + nf * Ha, seed=77 + s, noise=1e-3) for s in range(2)]))
    pv = make(N, Ha, Ha, betas)
    st = torch.zeros((2, pv.state_bytes()), dtype=torch.uint8, device="cuda")
    pv.process(x, k, state=st, flags=pvb200.CARRY_OUT)
    saved = st.clone()
    a = pv.process(x[:, k * Ha:], nf - k, state=st, flags=pvb200.CARRY_IN).cpu().numpy()
    assert torch.equal(st, saved)
    b = pv.process(x[:, k * Ha:], nf - k, state=st, flags=pvb200.CARRY_IN).cpu().numpy()
    assert np.array_equal(a, b)
    full = pv.process(x, nf).cpu().numpy()
    assert np.array_equal(a, full[:, :, k * Ha:])


def test_pcm16_host_path_takes_more_than_65535_streams():
    """rows ride in gridDim.y of the conversion kernels; both directions now loop over slabs."""
    N, H, nf, S = 64, 16, 6, 70000
    rng = np.random.default_rng(5)
    xi = rng.integers(-20000, 20000, size=(S, N + nf * H), dtype=np.int16)
    pv = make(N, H, H, [1.0])
    got = pv.process_host_pcm16(xi, nf)
    assert got.shape == (S, 1, nf * H)
    ref = pv.process_host_pcm16(np.ascontiguousarray(xi[65530:65540]), nf)
    assert np.array_equal(got[65530:65540], ref)

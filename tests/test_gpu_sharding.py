"""Frame-range sharding of one long stream on the GPU engine: four virtual ranks (threads with an
in-process all_gather) reproduce the single-launch result BIT FOR BIT in both modes -- the only data
crossing rank boundaries is the per-bin int64 phase carry (corrected) or nothing (compat)."""
import threading

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pvb200  # noqa: E402
from pvb200 import sharding  # noqa: E402
from sharding_engines import ThreadComm  # noqa: E402
from signals import multitone  # noqa: E402


def _run_ranks(world, fn):
    comms = ThreadComm.make(world)
    res, err = [None] * world, []

    def body(r):
        try:
            torch.cuda.set_device(0)
            res[r] = fn(comms[r])
        except Exception as e:       # pragma: no cover
            err.append(e)
            comms[r].sh["bar"].abort()

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not err, err
    return res


@pytest.mark.parametrize("N,Ha,Hs,betas,nf,world", [(2048, 512, 512, [1.4983071], 203, 4), (256, 64, 64, [1.0, 1.5, 2.0], 1001, 4),
                                                    (4096, 1024, 1024, [1.4983071], 81, 4),          # C5 shape
                                                    (1024, 256, 512, [1.0], 150, 3)])
def test_corrected_frame_sharding_is_bit_exact(N, Ha, Hs, betas, nf, world):
    x = torch.from_numpy(multitone(N + nf * Ha, seed=8, noise=1e-3)).cuda()
    pv = pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs, mode=pvb200.MODE_CORRECTED,
                             window_type=pvb200.WIN_HANN_PERIODIC, pitch=tuple(betas))
    full = pv.process(x[None, :], nf).cpu().numpy()
    launches0 = pv.launch_count()

    def fn(comm):
        # one handle per rank, as on real ranks (a handle is not thread-safe: it caches its segment plan)
        pvr = pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs, mode=pvb200.MODE_CORRECTED,
                                  window_type=pvb200.WIN_HANN_PERIODIC, pitch=tuple(betas))
        out, p = sharding.process_corrected_sharded(pvr, lambda k: x[None, k * Ha:], nf, comm, Ha, Hs, N)
        torch.cuda.synchronize()
        assert pvr.launch_count() >= 2
        return out.cpu().numpy()

    parts = _run_ranks(world, fn)
    got = np.concatenate(parts, axis=2)
    assert np.array_equal(got, full)
    assert launches0 in (1, 3)        # 3 when the single stream is split on the GPU (aggregate + states + process)


@pytest.mark.parametrize("N,Ha,Hs,nf,world", [(2048, 512, 512, 203, 4), (256, 64, 64, 999, 8), (4096, 1024, 1024, 50, 2)])
def test_compat_frame_sharding_is_bit_exact(N, Ha, Hs, nf, world):
    x = torch.from_numpy(multitone(N + nf * Ha, seed=9)).cuda()
    pv = pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs)
    full = pv.process(x[None, :], nf, n_analysed=nf - 2).cpu().numpy()

    def fn(comm):
        pvr = pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs)
        p = sharding.plan(nf, comm.world, comm.rank, N, Hs)
        out, _ = sharding.process_compat_sharded(pvr, x[None, p.ks * Ha:], nf, nf - 2, comm, Ha, Hs, N)
        torch.cuda.synchronize()
        return out.cpu().numpy()

    got = np.concatenate(_run_ranks(world, fn), axis=2)
    assert np.array_equal(got, full)


def test_aggregate_matches_oracle():
    import pv_oracle as po
    N, Ha, nf = 1024, 256, 40
    x = multitone(N + nf * Ha, seed=3, noise=1e-3)
    pv = pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Ha, mode=pvb200.MODE_CORRECTED, window_type=pvb200.WIN_HANN_PERIODIC)
    sumD, P_first, P_last = pv.aggregate(torch.from_numpy(x).cuda()[None, :], nf)
    want, wlast = po.corrected_aggregate(x, N, Ha, po.window(po.WIN_HANN_PERIODIC, N), nf)
    d = (P_last[0].cpu().numpy().view(np.uint32).astype(np.int64) - wlast.astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.abs(d).max() < 2 ** 32 * 1e-3
    # sums agree up to fp32 phase noise and whole-turn unwrap flips
    ds = (sumD[0].cpu().numpy() - want) / 2.0 ** 32
    frac = np.abs(ds - np.round(ds))
    assert frac.max() < 1e-3 and (np.round(ds) != 0).mean() < 0.02


def test_split_aggregate_matches_aggregate_and_reuse_changes_nothing():
    """pv_corrected_split_aggregate = the analysis pass of the processing call: its sums equal the plain aggregate's
    (integer, exact) and the processing call that reuses them produces the same bits as one that recomputes them."""
    N, H, nf = 1024, 256, 3000
    x = torch.from_numpy(np.stack([multitone(N + nf * H, seed=90 + s, noise=1e-3) for s in range(2)])).cuda()
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=pvb200.MODE_CORRECTED, window_type=pvb200.WIN_HANN_PERIODIC,
                             pitch=(float(np.float32(2 ** (7 / 12))),))
    # fresh start: D over frames 1..nf-1
    total = pv.split_aggregate(x, nf)
    n0 = pv.launch_count()
    a = pv.process(x, nf, flags=pvb200.REUSE_AGGREGATE)
    assert pv.launch_count() - n0 == 2                                  # state build + process: no second analysis pass
    plain, P_first, _ = pv.aggregate(x, nf)
    assert torch.equal(total, plain)
    n0 = pv.launch_count()
    b = pv.process(x, nf)
    assert pv.launch_count() - n0 == 3 and torch.equal(a, b)           # b: aggregate + state build + process
    # carried in at frame k: D over frames k..nf-1, the state supplies the phase of frame k-1
    k = 1000
    sd, P0, _ = pv.aggregate(x, k)
    P_km1 = pv.aggregate(x[:, (k - 1) * H:], 1)[1]
    st = pv.state_from_carry(P0, sd, k, P_km1)
    total_k = pv.split_aggregate(x[:, k * H:], nf - k, state=st)
    tail, _, _ = pv.aggregate(x[:, (k - 1) * H:], nf - k + 1)
    assert torch.equal(total_k, tail)
    c = pv.process(x[:, k * H:], nf - k, state=st, flags=pvb200.CARRY_IN | pvb200.REUSE_AGGREGATE)
    d = pv.process(x[:, k * H:], nf - k, state=st, flags=pvb200.CARRY_IN)
    halo = (N - 1) // H                    # the carried state has empty OLA accumulators: the first frames lack the tail
    assert torch.equal(c, d) and torch.equal(c[:, :, halo * H:], b[:, :, (k + halo) * H:])
    # the flag alone (nothing to reuse) is harmless
    assert torch.equal(pv.process(x, nf, flags=pvb200.REUSE_AGGREGATE), b)


@pytest.mark.parametrize("mode,N,Ha,Hs,betas,nf,world,S", [
    ("corrected", 2048, 512, 512, [1.4983071], 203, 4, 1), ("corrected", 4096, 1024, 1024, [1.4983071], 3000, 4, 2),   # C5 shape, stereo
    ("corrected", 256, 64, 64, [1.0, 1.5, 2.0], 1001, 8, 1), ("corrected", 1024, 256, 512, [1.0], 150, 3, 2),
    ("corrected", 512, 128, 128, [1.26], 7, 4, 1),                                                # ranges shorter than the halo, empty ranks
    ("compat", 2048, 512, 512, [1.0], 203, 4, 2), ("compat", 4096, 1024, 1024, [1.0], 2000, 2, 2)])
def test_sharding_through_the_c_abi_is_bit_exact(mode, N, Ha, Hs, betas, nf, world, S):
    """pv_shard_begin -> ONE all-gather of the carry records -> pv_shard_finish (include/pv_b200.h): the library does the
    analysis, the bookkeeping and the state rebuild; the caller only moves pv_shard_carry_elems() int64 per stream.
    Every rank sees only its own part of the input (range + halo + one frame), and the concatenated ranges equal the
    single-call result bit for bit."""
    corrected = mode == "corrected"
    x = torch.from_numpy(np.stack([multitone(N + nf * Ha, seed=8 + s, noise=1e-3) for s in range(S)])).cuda()
    mk = lambda: pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs, mode=pvb200.MODE_CORRECTED if corrected else pvb200.MODE_COMPAT,
                                     window_type=pvb200.WIN_HANN_PERIODIC if corrected else pvb200.WIN_HAMMING, pitch=tuple(betas))
    n_an = nf if corrected else nf - 2
    full = mk().process(x, nf, n_analysed=n_an).cpu().numpy()

    def fn(comm):
        pvr = mk()
        p = pvr.shard_plan(nf, comm.world, comm.rank)
        first = max(0, p.ks - 1)                                   # the rank's own view: nothing before its halo
        last = min(x.shape[1], (max(p.k1, 1) - 1) * Ha + N)        # ... and nothing after its last frame
        xr = x[:, first * Ha:last].contiguous() if p.k1 > p.k0 else x[:, :N].contiguous()
        assert pvr.shard_carry_elems() == (N // 2 + 1) + (N // 2 + 2) // 2
        out, _ = sharding.process_sharded_capi(pvr, xr, first if p.k1 > p.k0 else 0, nf, comm, n_analysed=n_an)
        torch.cuda.synchronize()
        return out.cpu().numpy()

    got = np.concatenate(_run_ranks(world, fn), axis=2)
    assert np.array_equal(got, full)

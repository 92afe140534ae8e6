"""Frame-range sharding of ONE long stream over ranks (GPUs) -- SURVEY 8e, second row.

Independent streams need nothing from this module: they are simply dealt to the ranks.  A single
long stream is cut into contiguous frame ranges.  What crosses rank boundaries:

  compat     nothing.  Frames are independent; the (N-1)//Hs frames in front of a range are
             recomputed from the input halo and skipped on output (pv_process_device_ex).
  corrected  the per-bin phase carry: every rank reduces its own range to sumD[bin] (int64 sum of
             the unwrapped phase differences) with an analysis-only pass, the sums are exchanged with
             ONE all-gather of (N/2+1) int64 per rank, and each rank rebuilds the accumulator state at
             the start of its halo.  Integer addition is associative, so the sharded result is
             bit-identical to the single-pass one.

The module is engine-agnostic: `engine` provides process / aggregate / state_from_carry (the GPU
engine is pvb200.PhaseVocoder; the CPU tests plug an oracle-backed engine in), `comm` provides
all_gather (torch.distributed over NCCL on GPUs, gloo in the CPU tests, or LocalComm for one rank).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class ShardPlan:
    k0: int        # first frame this rank emits
    k1: int        # one past the last
    ks: int        # first frame it computes (k0 - halo, clipped at 0)
    halo: int


def plan(n_frames: int, world: int, rank: int, N: int, Hs: int) -> ShardPlan:
    per = (n_frames + world - 1) // world
    k0 = min(n_frames, rank * per)
    k1 = min(n_frames, k0 + per)
    halo = (N - 1) // Hs
    return ShardPlan(k0, k1, max(0, k0 - halo), halo)


def shard_streams(n_streams: int, world: int, rank: int):
    """Contiguous block of streams for this rank (no communication needed)."""
    per = (n_streams + world - 1) // world
    lo = min(n_streams, rank * per)
    return lo, min(n_streams, lo + per)


class LocalComm:
    """world_size 1."""
    world, rank = 1, 0

    def all_gather(self, t):
        return [t]


class TorchComm:
    """torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()

    def all_gather(self, t):
        import torch
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t.contiguous())
        return out


def process_sharded_capi(pv, x, first_frame, n_frames, comm, n_analysed=None):
    """The same through the C ABI (pv_shard_begin / pv_shard_finish, include/pv_b200.h): the library does the
    analysis, the carry bookkeeping, the state rebuild and the processing; this function only issues the ONE
    collective in between -- what a C caller does with ncclAllGather.  x [S, n]: the stream(s) from sample
    first_frame*hop_in on (a rank needs its range, the overlap-add halo and one more frame; rank 0 needs frame 0).
    Returns (out [S, V, (k1-k0)*Hs], plan).  Both modes; compat contributes an all-zero record."""
    import torch
    carry = pv.shard_begin(x, first_frame, n_frames, comm.world, comm.rank)
    gathered = torch.stack(comm.all_gather(carry)) if comm.world > 1 else carry[None]
    out = pv.shard_finish(x, first_frame, n_frames, comm.world, comm.rank, gathered.contiguous(), n_analysed=n_analysed)
    return out, pv.shard_plan(n_frames, comm.world, comm.rank)


def process_compat_sharded(engine, x, n_frames, n_analysed, comm, Ha, Hs, N):
    """x: the rank's view of the stream(s) as a [S, n] tensor starting at sample ks*Ha.  Returns this
    rank's output block [S, 1, (k1-k0)*Hs] and its plan."""
    p = plan(n_frames, comm.world, comm.rank, N, Hs)
    nf = p.k1 - p.ks
    if p.k1 <= p.k0:
        return None, p
    out = engine.process(x, nf, n_analysed=max(0, min(nf, n_analysed - p.ks)), skip=p.k0 - p.ks)
    return out, p


def process_corrected_sharded(engine, x_from, n_frames, comm, Ha, Hs, N):
    """Corrected mode.  `x_from(first_frame)` returns the stream(s) from sample first_frame*Ha on as a
    [S, n] tensor on the engine's device (S streams cut at the same frames, e.g. the channels of a file).
    Returns (out [S, V, (k1-k0)*Hs], plan).

    Engines with `split_aggregate` (the GPU engine) learn their contribution from the very analysis pass that
    their processing call would run anyway and reuse it afterwards: one pass over the range instead of two."""
    if hasattr(engine, "split_aggregate"):
        return _process_corrected_sharded_reuse(engine, x_from, n_frames, comm, Ha, Hs, N)
    p = plan(n_frames, comm.world, comm.rank, N, Hs)
    empty = p.k1 <= p.k0
    # 1. local aggregate over [k0, k1): D_k needs P_{k0-1}, so start one frame early (rank 0: frame 0)
    a0 = max(0, p.k0 - 1)
    na = 0 if empty else p.k1 - a0
    sumD, P_first, _ = engine.aggregate(x_from(a0), max(na, 1))      # an empty rank still joins the gather
    if empty:
        sumD = sumD * 0
    # 2. the only exchange: per-bin sums (and rank 0's P_0)
    sums = comm.all_gather(sumD)
    P0 = comm.all_gather(P_first)[0]
    if empty:
        return None, p
    if p.ks == 0:
        return engine.process(x_from(0), p.k1, skip=p.k0), p
    # 3. prefix up to ks: everything before k0, minus the D of the halo frames [ks, k0)
    prefix = sum(sums[:comm.rank]) if comm.rank > 0 else sumD * 0
    h_sum, P_ksm1, _ = engine.aggregate(x_from(p.ks - 1), p.k0 - p.ks + 1)
    prefix = prefix - h_sum
    state = engine.state_from_carry(P0, prefix, p.ks, P_ksm1)
    # 4. halo frames fill the OLA accumulators, then the owned range is written
    CARRY_IN = 1
    out = engine.process(x_from(p.ks), p.k1 - p.ks, state=state, flags=CARRY_IN, skip=p.k0 - p.ks)
    return out, p


def _process_corrected_sharded_reuse(engine, x_from, n_frames, comm, Ha, Hs, N):
    CARRY_IN, REUSE = 1, 4
    p = plan(n_frames, comm.world, comm.rank, N, Hs)
    empty = p.k1 <= p.k0
    # phase of frame 0 (rank 0's is the one everybody needs) and a zero of the right shape
    zero, P_0, _ = engine.aggregate(x_from(0), 1)
    zero = zero * 0
    owned, h_sum, P_ksm1 = zero, None, None
    if not empty and p.ks == 0:
        # the range reaches frame 0: fresh start, D over [1, k1); what lies before k0 is not this rank's to report
        total = engine.split_aggregate(x_from(0), p.k1, skip=p.k0)
        owned = total - engine.aggregate(x_from(0), p.k0)[0] if p.k0 > 1 else total
    elif not empty:
        # D over the halo [ks, k0) and the phase of frame ks-1, from a handful of frames
        h_sum, P_ksm1, _ = engine.aggregate(x_from(p.ks - 1), p.k0 - p.ks + 1)
        st0 = engine.state_from_carry(P_0 * 0, zero, 1, P_ksm1)          # previous phase only; accumulators unknown yet
        total = engine.split_aggregate(x_from(p.ks), p.k1 - p.ks, state=st0, skip=p.k0 - p.ks)   # D over [ks, k1)
        owned = total - h_sum
    # the only exchange: per-bin sums (and rank 0's P_0)
    sums = comm.all_gather(owned)
    P0 = comm.all_gather(P_0)[0]
    if empty:
        return None, p
    if p.ks == 0:
        return engine.process(x_from(0), p.k1, skip=p.k0, flags=REUSE), p
    prefix = (sum(sums[:comm.rank]) if comm.rank > 0 else zero) - h_sum      # up to ks
    state = engine.state_from_carry(P0, prefix, p.ks, P_ksm1)
    out = engine.process(x_from(p.ks), p.k1 - p.ks, state=state, flags=CARRY_IN | REUSE, skip=p.k0 - p.ks)
    return out, p

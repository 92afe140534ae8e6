"""Engines and communicators for the frame-range sharding tests."""
import threading

import numpy as np
import torch

import pv_oracle as po


class ThreadComm:
    """In-process all_gather between `world` threads (virtual ranks on one device)."""

    def __init__(self, world, rank, shared):
        self.world, self.rank, self.sh = world, rank, shared

    @staticmethod
    def make(world):
        shared = {"bar": threading.Barrier(world), "slots": {}}
        return [ThreadComm(world, r, shared) for r in range(world)]

    def all_gather(self, t):
        key = self.sh.setdefault("n%d" % self.rank, 0)
        self.sh["n%d" % self.rank] = key + 1
        self.sh["slots"][(key, self.rank)] = t
        self.sh["bar"].wait()
        out = [self.sh["slots"][(key, r)] for r in range(self.world)]
        self.sh["bar"].wait()
        return out


class OracleEngine:
    """CPU engine with the interface sharding.py expects, backed by the oracle (tests only)."""

    def __init__(self, N, Ha, Hs, betas, mode="corrected", win_type=None):
        self.N, self.Ha, self.Hs, self.betas, self.mode = N, Ha, Hs, [float(np.float32(b)) for b in betas], mode
        self.win = po.window(po.WIN_HANN_PERIODIC if mode == "corrected" else po.WIN_HAMMING, N) if win_type is None \
            else po.window(win_type, N)
        self.nb = N // 2 + 1
        self.tabs = [po.corrected_tables(N, Ha, Hs, b) for b in self.betas]

    def aggregate(self, x, n_frames, P_prev=None):
        xs = x[0].numpy()
        sumD, P_last = po.corrected_aggregate(xs, self.N, self.Ha, self.win, n_frames)
        _, P_first = po.corrected_aggregate(xs, self.N, self.Ha, self.win, 1)
        as_t = lambda a, dt: torch.from_numpy(a.astype(dt))[None, :]
        return as_t(sumD, np.int64), as_t(P_first.view(np.int32), np.int32), as_t(P_last.view(np.int32), np.int32)

    def state_from_carry(self, P_first, sumD, n_before, P_prev):
        st = po.CorrectedState(self.N, len(self.betas))
        st.have_prev = 1
        st.P_prev = P_prev[0].numpy().view(np.uint32).copy()
        P0 = P_first[0].numpy().view(np.uint32)
        sd = sumD[0].numpy()
        for v, t in enumerate(self.tabs):
            ok = t["a_lo"] <= t["a_hi"]
            a = t["a_hi"][ok]
            st.psi[v][ok] = (P0[a].astype(np.uint64) << np.uint64(32)) + np.uint64(n_before - 1) * t["nomS"][ok] \
                + (sd[a] * np.int64(t["Rq"])).astype(np.uint64)
        return st

    def process(self, x, n_frames, n_analysed=None, state=None, flags=0, skip=0):
        xs = x[0].numpy()
        if self.mode == "corrected":
            out, _ = po.process_corrected(xs, self.N, self.Ha, self.Hs, self.win, self.betas, n_frames,
                                          state=state.copy() if state is not None else None)
        else:
            na = n_frames if n_analysed is None else n_analysed
            o, _ = po.process_compat(xs, self.N, self.Ha, self.Hs, self.win, na, n_frames)
            out = o[None, :]
        return torch.from_numpy(out[:, skip * self.Hs:].astype(np.float64))[None]

"""Decision-aligned parity check of the corrected mode (test helper; see DESIGN.md "conditioning of the phase unwrap").

The phase-difference unwrap D = (int32)(P_k - P_{k-1} - nomA) is discontinuous at +-1/2 turn.  An implementation whose
fp32 phase of a bin differs from the oracle's by delta turns gets the SAME D (up to delta) unless the oracle's D lies
within delta of the boundary; then it gets the neighbouring alias, D -+ 1 turn, and for a fractional R = beta*Hs/Ha that
bin's accumulator differs by R turns from then on.  No choice of arithmetic removes this: the map is discontinuous
(tools/spec_conditioning.py measures it, and shows that peak-picking / phase locking has MORE such decisions).

So parity is checked in three steps, each with a hard bound:
  1. per-bin phase parity: every D of the implementation equals the oracle's D modulo one turn, within the bin's fp32
     phase uncertainty (tolerance scales with frame-max / bin-magnitude);
  2. alias disagreements ("flips") are rare, and by (1) each one is a boundary case of exactly that size;
  3. with those decisions moved in the oracle (pvo_corrected_trace.unwrap_adjust), the OUTPUT matches to the full
     tolerance (SNR >= 100 dB), which shows that nothing but the aliases differed.
"""
import numpy as np

import pv_oracle as po
from signals import snr_db

TURN = 2.0 ** 32


def compare_decisions(D_impl, D_or, mag_or, eps=2e-7, floor=2e-6):
    """Returns (adjust int8 [frames, bins], max ratio |phase disagreement| / tolerance, number of flips, fraction of
    the bins within 100 dB of their frame's peak that flipped -- bins at the fp32 noise floor flip freely and carry
    no energy)."""
    diff = D_impl.astype(np.int64) - D_or.astype(np.int64)
    adj = np.rint(diff / TURN).astype(np.int64)
    resid = np.abs(diff - adj * (1 << 32)) / TURN                    # turns, after removing whole-turn aliases
    M = mag_or.max(axis=1, keepdims=True)
    inv = M / np.maximum(mag_or, 1e-300)                             # frame max / bin magnitude
    inv_prev = np.vstack([inv[:1], inv[:-1]])
    tol = floor + eps * (inv + inv_prev)                             # D involves the phases of frames k-1 and k
    ratio = resid / tol
    ratio[0] = 0.0                                                   # the first frame has no phase difference
    sig = (inv <= 1e5) & (inv_prev <= 1e5)
    sig[0] = False
    frac_sig = float(np.count_nonzero(adj[sig])) / max(1, int(sig.sum()))
    return adj.astype(np.int8), float(ratio.max()), int(np.count_nonzero(adj[1:])), frac_sig


def aligned_parity(x, N, Ha, Hs, win, betas, n_frames, D_impl, out_impl):
    """out_impl: [V, n_frames*Hs].  Returns dict(direct=[dB per voice], aligned=[dB per voice], flips, phase_ratio)."""
    want, D_or, mag = po.process_corrected_traced(x, N, Ha, Hs, win, betas, n_frames)
    adj, ratio, flips, frac_sig = compare_decisions(D_impl, D_or, mag)
    direct = [snr_db(want[v], out_impl[v]) for v in range(len(betas))]
    if flips:
        want2, D2, _ = po.process_corrected_traced(x, N, Ha, Hs, win, betas, n_frames, unwrap_adjust=adj)
        assert np.array_equal(D2, D_or)                              # the adjustment does not change the analysis
        aligned = [snr_db(want2[v], out_impl[v]) for v in range(len(betas))]
    else:
        aligned = direct
    return dict(direct=direct, aligned=aligned, flips=flips, phase_ratio=ratio, frac=frac_sig)

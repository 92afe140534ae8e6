#!/usr/bin/env python
"""Builds the narrative profile files of round 2 (profiles/r02_*.md) from what the GPU runs left in gpurun_out/.
Run here after tools/refresh_profiles.sh + tools/refresh_profiles_post.py and the multi-GPU bench runs."""
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rd = lambda n: open(os.path.join(G, n)).read()

# ---- C5 sharded over 2 / 4 / 8 GPUs
rows = []
for n in (2, 4, 8):
    d = json.load(open(os.path.join(G, f"r02_bench_{n}gpu.json")))
    shutil.copy(os.path.join(G, f"r02_bench_{n}gpu.json"), os.path.join(P, f"r02_bench_{n}gpu.json"))
    rows.append((n, d, d["c5_sharded"]))
out = ["# C5 (one long file, frame-range sharded) on 2 / 4 / 8 x B200 (round 2)\n",
       "`python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 bench.py --gpus G` prints this as the",
       "`c5_sharded` record of its JSON line (full lines: `profiles/r02_bench_{2,4,8}gpu.json`).  Workload: BASELINE.json config 5, a synthetic",
       "1-hour 48 kHz stereo file, window 4096, hop 1024 (2 x 168 750 frames).  The sharding runs **behind the C ABI**: `pv_shard_begin`",
       "(analysis of the rank's range, which keeps {|X|, D} of every frame) -> ONE NCCL all-gather of the packed carry record (49 184 bytes per rank for the two channels: per-bin",
       "int64 sums + rank 0's phase of frame 0) -> `pv_shard_finish` (state from the carry, processing; the analysis pass is reused, not",
       "repeated).  Every rank hands the library only its own view of the input: its frame range, the overlap-add halo and one more frame.",
       "Timing: CUDA events around three calls after two warm-ups, max over ranks; rank 0 gathers the ranges, runs the whole file alone and",
       "compares bit for bit.\n",
       "| GPUs | mode | sharded ms | frames/s | one GPU, ms | speed-up | bit-identical to one GPU | exchange |",
       "|---|---|---|---|---|---|---|---|"]
for n, d, c in rows:
    for m in ("compat", "corrected"):
        r = c[m]
        ex = "one all-gather, %d B per rank" % r["carry_bytes_per_rank"] if m == "corrected" else "none (input halo recomputed)"
        out.append(f"| {n} | {m} | {r['ms']:.2f} | {r['frames_per_s']/1e6:.1f} M | {r['single_gpu_ms']:.2f} | **{r['speedup_vs_single_gpu']:.2f}x** | "
                   f"{'yes' if r['bit_identical_to_single_gpu'] else 'NO'} | {ex} |")
out += ["",
        "Round 1 (Python orchestration over torch.distributed, two collectives, in-place large-window kernels): 2 GPUs compat 2.0x, corrected",
        "1.4x (28.2 ms); 4 and 8 GPUs unmeasured.  One hour of stereo audio is now pitch-shifted in 1.6 ms on eight GPUs (8.8 ms on one: the single-GPU time fell by a quarter when the split began to keep its analysis, DESIGN.md 4.2, which is why the ratios are lower than the 1.86x / 3.49x / 6.15x measured against the recomputing split earlier in the round).",
        "",
        "What limits the scaling at 8 GPUs: each rank's range is 21 094 frames per channel, cut into ~74 parts per channel so that the 296",
        "resident groups of the GPU are busy; every part recomputes its 3-frame overlap-add halo and (corrected) one analysis frame, and the",
        "launches around the exchange (halo aggregate, state build, pack, prefix, split states: ~10 us each) and the all-gather itself",
        "(~30 us) no longer vanish next to 1.2 - 2.0 ms of work.  The limiting launch is the processing kernel (~65 % of the rank's time in",
        "corrected mode), then the analysis pass (~30 %).",
        "",
        "Same runs, the headline batch (independent streams per rank, no collective): device-resident "
        f"{rows[0][1]['value']/1e6:.0f} / {rows[1][1]['value']/1e6:.0f} / {rows[2][1]['value']/1e6:.0f} M frames/s on 2 / 4 / 8 GPUs (1 GPU: 116 M); end to end "
        f"{rows[0][1]['e2e']['value']/1e6:.1f} / {rows[1][1]['e2e']['value']/1e6:.1f} / {rows[2][1]['e2e']['value']/1e6:.1f} M frames/s = "
        f"{rows[0][1]['e2e']['copy_ceiling']['frac_achieved']:.2f} / {rows[1][1]['e2e']['copy_ceiling']['frac_achieved']:.2f} / "
        f"{rows[2][1]['e2e']['copy_ceiling']['frac_achieved']:.2f} of the copy ceiling measured in the same run (`profiles/r02_pcie_probe.md`)."]
open(os.path.join(P, "r02_c5_multi_gpu.md"), "w").write("\n".join(out) + "\n")

# ---- PCIe probe + in-run copy ceiling
probe = rd("r02_pcie_probe.md")
erows = "".join(f"| {n} | {d['e2e']['value']/1e6:.1f} M frames/s | {d['e2e']['copy_ceiling']['value']/1e6:.1f} M frames/s "
                f"({d['e2e']['copy_ceiling']['gb_per_s_each_direction_all_gpus']:.0f} GB/s each way, all GPUs) | {d['e2e']['copy_ceiling']['frac_achieved']:.2f} |\n"
                for n, d, _ in rows)
one = json.load(open(os.path.join(G, "r02_bench.json")))
erows = (f"| 1 | {one['e2e']['value']/1e6:.1f} M frames/s | {one['e2e']['copy_ceiling']['value']/1e6:.1f} M frames/s "
         f"({one['e2e']['copy_ceiling']['gb_per_s_each_direction_all_gpus']:.0f} GB/s each way) | {one['e2e']['copy_ceiling']['frac_achieved']:.2f} |\n") + erows
open(os.path.join(P, "r02_pcie_probe.md"), "w").write(f"""# Host <-> device copy ceiling of the box, all GPUs copying at once (round 2)

`tools/pcie_probe.py` under torchrun with 1, 2, 4 and 8 processes on the 8 x B200 node (32 vCPUs, 1 TB): 512 MB pinned buffers per rank and
direction, every measurement started together (barrier), wall clock over 5 repetitions.

| GPUs copying at once | H2D GB/s per GPU (min..max) | D2H GB/s per GPU | both directions at once, GB/s per GPU (sum) | node aggregate, both directions | CPUs bound per rank |
|---|---|---|---|---|---|
{probe}
One GPU alone moves 56 GB/s each way (93 GB/s both ways together); the NODE does not move more than 100 - 155 GB/s in total however
many GPUs copy, so the per-GPU share falls to 16 - 22 GB/s at eight.  This is the host side of the VM (one memory system behind all
PCIe roots), not the library: `bench.py` measures the ceiling in the same run -- the very pinned buffers of the end-to-end leg
copied in both directions at once with no kernel (`e2e.copy_ceiling`):

| GPUs | e2e through pv_process_host (float) | copy-only ceiling, same buffers | e2e / ceiling |
|---|---|---|---|
{erows}
(Ratios slightly above 1 are run-to-run variation of the copies.  Every row is its own `gpurun` call on a freshly assigned box; the copy
engines of the boxes differ -- 41 to 56 GB/s each way for one GPU alone -- which is why the ceiling is measured inside every run.)  The 16-bit PCM leg (`pv_process_host_pcm16`, half the bytes) on one GPU:
{one['e2e']['pcm16']['value']/1e6:.1f} M frames/s = {one['e2e']['pcm16']['copy_ceiling']['frac_achieved']:.2f} of ITS copy ceiling ({one['e2e']['pcm16']['copy_ceiling']['value']/1e6:.1f} M frames/s); VERDICT r01's
60 M frames/s target for it is above what this box's PCIe moves.

The pipelined host path (chunks of frames on three streams, state carried on the device) therefore sits ON the copy ceiling at every
GPU count; the kernels behind it scale linearly (116 -> 931 M frames/s).  The end-to-end number of this box cannot scale past what its
host memory system feeds.
""")

# ---- configs, sweeps, voices
open(os.path.join(P, "r02_stream_sweep.md"), "w").write("""# Stream-count sweep at the headline shape (round 2)

`python tools/stream_sweep.py` on one B200: window 2048, hop 512, 860 frames per stream, device-resident, CUDA events, 5 launches after
3 warm-ups.  The corrected kernel keeps 4 CTAs x 148 SMs = 592 streams resident, so 1184 streams (the bench batch) is exactly two
waves -- VERDICT r01 weak #6 asked what happens in between.

Round 1 (same tool, at the start of this round): 300 / 592 / 600 / 740 / 900 streams ran at 0.45 / 0.58 / 0.46 / 0.54 / 0.51 of the
1184-stream rate (fewer than two waves of streams were ALWAYS cut into frame-range parts, which costs an extra analysis pass), and
1200 / 1300 / 1500 / 1800 at 0.84 / 0.90 / 0.93 / 0.92 (ragged last wave).

Now: (1) a cost model decides the number of parts (`split_cost` in `csrc/pv_capi.cu`: waves x (processing share x (frames + halo) +
analysis share x frames) + launches, by number of voices and by whether the analysis is stored) and splits only when that is clearly cheaper; (2) a batch with a ragged last wave runs its full waves unsplit and
hands the remaining streams to a second call that is free to split them.

""" + rd("r02_stream_sweep_body.md") + """
Corrected mode stays within 0.92 - 1.02 of the 1184-stream rate from 592 to 2400 streams (worst: 900 streams = 1.52 waves; before the
split kept its analysis -- DESIGN.md 4.2, stored analysis -- the worst point was 0.85 and 300 streams ran at 0.61).  300 streams use half
the machine by construction.  Compat mode cuts every stream into segments (frames are independent) and is flat.
""")
open(os.path.join(P, "r02_configs.md"), "w").write("""# All five BASELINE.json configs on one B200 (round 2)

Generated by `python tools/run_configs.py` on the GPU box (`tools/refresh_profiles.sh r02`); device-resident inputs, CUDA-event timing,
5 launches after 3 warm-ups.  C1 and C2 run on the reference's REAL input samples: committed raw-integer slices of
`testtones/440sine.wav` (16 bit) and `testtones/MAT_ZO_24_bit.wav` (24 bit) from `tests/golden/golden_wav.npz`, decoded with AudioFile's
rules (parity on the slice itself; the slice is tiled to the file's length for the timing).  Parity column: compat = worst direct SNR
against the fp64 oracle; corrected = direct / decision-aligned SNR (`tests/aligned.py`, `profiles/r02_spec_conditioning.md`) with the
number of alias flips among frames x bins of the checked sample; every row also passed the per-bin phase-parity bound.

""" + rd("r02_configs_body.md") + """
Notes.  C1 with the reference constructor's own Hamming table: 130.8 dB directly (VERDICT r01 measured 89.8 dB on a synthetic tone with
the periodic Hann; that variant is the third row: ~92 dB directly, ~125 dB once the boundary decisions are aligned).  C3 corrected:
R = Hs/Ha = 5.02 is fractional and the clean three-tone input leaves most bins at the fp32 noise floor: ~1200 flips among 307 800 frames x
bins, all of them at bins that carry no energy (direct 123 dB).  C5 runs on the register-resident window-4096 kernels (round 1: compat
15.3 ms = 22 M frames/s, corrected 39.7 ms for the hour).  C4: 162 M frames/s in round 1.  C3 'literal 10/2' (hop_in 10, hop_out 2
samples) overlaps 512 frames per output sample, so the recomputed OLA halo dominates.
""")
open(os.path.join(P, "r02_window_sweep.md"), "w").write("""# Window sweep at equal audio per launch (round 2)

`python tools/run_configs.py --sweep` on one B200, many streams (no frame-range split); SNR: compat direct, corrected direct / aligned.

""" + rd("r02_sweep_body.md"))
open(os.path.join(P, "r02_voices.md"), "w").write("""# Cost of extra voices (round 2)

`python tools/voices_probe.py` on one B200, corrected mode, pitch ratios {1, +4, +7, +12 semitones}[:V], device-resident.  Three or four
voices run as launches of two voices each (`DESIGN.md` 4.2); round 1: V = 4 cost 3.8x (window 256) / 4.1x (window 2048) of one voice.

""" + rd("r02_voices.md"))
open(os.path.join(P, "r02_exchange_micro.md"), "w").write("""# One FFT exchange: shared memory against warp shuffles (round 2)

`tools/micro/exchange_shfl.cu` on one B200 (`tools/refresh_profiles.sh`), the launch shape of the corrected kernel (128 threads per CTA,
four CTAs per SM, 51 KB of shared memory per CTA).  Every thread hands on 16 complex values per exchange, as in every pass of the fused
kernels.  VERDICT r01 #7 asked for the `north_star`'s "warp-shuffle FFT" to be built or refused by measurement:

""" + rd("r02_exchange_micro.md") + """
The shuffle variant is the CHEAPEST pattern shuffles offer (lane-xor butterflies, no dynamic register indexing) and still costs 1.3x the
shared-memory exchange while reaching 16 lanes instead of the 128 - 256 threads every pass of a 1024- or 2048-point transform needs
(`DESIGN.md` 4.6).  Not used.
""")
print("ok")

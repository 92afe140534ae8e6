#!/usr/bin/env python
"""C5: one long stereo file, window 4096 / hop 1024, frame-range sharded over the ranks of one
box with torch.distributed / NCCL.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/shard_long_file.py

Every rank generates the whole input (so that rank 0 can check the result) but hands the library only ITS view: its
frame range, the overlap-add halo and one more frame.  The sharding itself runs behind the C ABI (pv_shard_begin ->
ONE all-gather of pv_shard_carry_elems() int64 per channel -> pv_shard_finish; include/pv_b200.h).  Rank 0 then checks
the concatenated result bit for bit against a single-GPU run of the same file and prints a JSON line with the timing
(max over ranks, CUDA events)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import pvb200
from pvb200 import sharding
from signals import multitone


def main():
    mode = os.environ.get("PV_MODE", "corrected")
    seconds = float(os.environ.get("PV_SECONDS", "600"))          # 10 min by default (1 h = 3600)
    N, H, fs = 4096, 1024, 48000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nf = int(seconds * fs) // H
    n = N + nf * H
    base = np.stack([multitone(min(n, 1 << 20), fs=fs, seed=c, noise=1e-3) for c in range(2)])
    x = torch.from_numpy(np.tile(base, (1, (n + base.shape[1] - 1) // base.shape[1]))[:, :n].copy()).cuda()
    beta = float(np.float32(2 ** (7 / 12)))
    corrected = mode == "corrected"
    mk = lambda: pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, device=local,
                                     mode=pvb200.MODE_CORRECTED if corrected else pvb200.MODE_COMPAT,
                                     window_type=pvb200.WIN_HANN_PERIODIC if corrected else pvb200.WIN_HAMMING, pitch=(beta,))
    pv = mk()
    comm = sharding.TorchComm()

    p0 = pv.shard_plan(nf, world, rank)
    first = max(0, p0.ks - 1)
    xr = x[:, first * H:min(n, (max(p0.k1, 1) - 1) * H + N)].contiguous() if p0.k1 > p0.k0 else x[:, :N].contiguous()
    first = first if p0.k1 > p0.k0 else 0

    def run_sharded():
        # both channels in one call: they are independent streams cut at the same frames
        return sharding.process_sharded_capi(pv, xr, first, nf, comm)

    for _ in range(2):
        out, p = run_sharded()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        out, p = run_sharded()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # gather the ranges on rank 0 and compare with a single-GPU run
    per = (nf + world - 1) // world
    pad = torch.zeros((2, 1, per * H), device="cuda")
    pad[:, :, :out.shape[2]] = out
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, parts, dst=0)
    if rank == 0:
        got = torch.cat(parts, 2)[:, :, :nf * H]
        one = mk()
        ref = one.process(x, nf)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            one.process(x, nf, out=ref)
        t1.record()
        torch.cuda.synchronize()
        single_ms = t0.elapsed_time(t1) / reps
        exact = bool(torch.equal(got, ref))
        print(json.dumps({"config": "C5 long file", "mode": mode, "window": N, "hop": H, "channels": 2, "frames_per_channel": nf,
                          "audio_seconds": nf * H / fs, "n_gpus": world, "ms": float(ms.item()),
                          "frames_per_s": 2 * nf / (float(ms.item()) * 1e-3),
                          "exchange": f"one all_gather of {pv.shard_carry_elems()} int64 per channel and rank (C ABI: pv_shard_begin/finish)" if corrected else "none (input halo recomputed)",
                          "single_gpu_ms": single_ms, "speedup_vs_single_gpu": single_ms / float(ms.item()),
                          "bit_identical_to_single_gpu": exact}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""world_size-2 gloo test (CPU) of the frame-range sharding logic: two ranks exchange the per-bin phase
carry with ONE all_gather and reproduce the single-pass result bit for bit.  The compute engine is
the oracle here (the GPU version of the same test is tests/test_gpu_sharding.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "phase-vocoder_b200"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def _worker(rank, world, port, mode, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pv_oracle as po
    from pvb200 import sharding
    from sharding_engines import OracleEngine
    from signals import multitone
    N, Ha, Hs, nf = 512, 128, 128, 57
    betas = [1.0, 1.26] if mode == "corrected" else [1.0]
    x = multitone(N + nf * Ha, seed=4, noise=1e-3)
    eng = OracleEngine(N, Ha, Hs, betas, mode)
    comm = sharding.TorchComm()
    xt = torch.from_numpy(x)
    if mode == "corrected":
        out, p = sharding.process_corrected_sharded(eng, lambda k: xt[None, k * Ha:], nf, comm, Ha, Hs, N)
    else:
        pl = sharding.plan(nf, world, rank, N, Hs)
        out, p = sharding.process_compat_sharded(eng, xt[None, pl.ks * Ha:], nf, nf - 1, comm, Ha, Hs, N)
    np.save(os.path.join(tmp, f"out_{mode}_{rank}.npy"), out[0].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["corrected", "compat"])
def test_two_ranks_reproduce_single_pass(tmp_path, mode):
    import pv_oracle as po
    from signals import multitone
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, mode, str(tmp_path)), nprocs=world, join=True)
    N, Ha, Hs, nf = 512, 128, 128, 57
    x = multitone(N + nf * Ha, seed=4, noise=1e-3)
    got = np.concatenate([np.load(tmp_path / f"out_{mode}_{r}.npy") for r in range(world)], axis=1)
    if mode == "corrected":
        betas = [float(np.float32(b)) for b in (1.0, 1.26)]
        want, _ = po.process_corrected(x, N, Ha, Hs, po.window(po.WIN_HANN_PERIODIC, N), betas, nf)
    else:
        w, _ = po.process_compat(x, N, Ha, Hs, po.window(po.WIN_HAMMING, N), nf - 1, nf)
        want = w[None, :]
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_plan_covers_every_frame_once():
    from pvb200 import sharding
    for nf in (1, 7, 57, 1000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                p = sharding.plan(nf, world, r, 2048, 512)
                assert p.ks == max(0, p.k0 - 3) and p.halo == 3
                seen += list(range(p.k0, p.k1))
            assert seen == list(range(nf))
    assert sharding.shard_streams(10, 4, 3) == (9, 10) and sharding.shard_streams(10, 4, 0) == (0, 3)

#!/usr/bin/env python
"""Pinned host<->device copy bandwidth, one process per GPU, ALL ranks copying at the same time: the ceiling of
bench.py's end-to-end number on this box (the 8-GPU ceiling was asserted, not measured).

    python tools/pcie_probe.py                                            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py

Per rank: H2D alone, D2H alone, both directions at once (two streams), 512 MB each way, pinned buffers first-touched on
the CPUs next to the GPU (same NUMA binding as bench.py).  Ranks start every measurement together (barrier); the table
gives per-rank min / max and the node aggregate."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import bench
    cpus = bench.bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 512 << 20
    h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h1.zero_(); h2.zero_()
    d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def both():
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    res = torch.tensor([n / timed(lambda: d1.copy_(h1, non_blocking=True)) / 1e9,
                        n / timed(lambda: h2.copy_(d2, non_blocking=True)) / 1e9,
                        2 * n / timed(both) / 1e9], device="cuda", dtype=torch.float64)
    allr = [torch.empty_like(res) for _ in range(world)] if dist else [res]
    if dist:
        dist.all_gather(allr, res)
    if rank == 0:
        r = torch.stack(allr).cpu()
        print(f"| GPUs copying at once | H2D GB/s per GPU (min..max) | D2H GB/s per GPU | both directions, GB/s per GPU (sum of the two) | "
              f"node aggregate, both directions | CPUs bound per rank |\n|---|---|---|---|---|---|")
        print(f"| {world} | {r[:,0].min():.1f}..{r[:,0].max():.1f} | {r[:,1].min():.1f}..{r[:,1].max():.1f} | "
              f"{r[:,2].min():.1f}..{r[:,2].max():.1f} | {r[:,2].sum():.0f} GB/s | {len(cpus) if cpus else 'all'} |", flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

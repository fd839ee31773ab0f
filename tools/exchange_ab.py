#!/usr/bin/env python3
"""A/B of the root exchange of the subtree-sharded build under torchrun: peer-memory mailboxes (k_exchange_top) against
ncclAllGather + finish launch, SAME process, same leaves, alternating.  Weak case (2^24 leaves per rank, K builds back to back,
as bench.py times them) and strong cases (one 2^24-leaf tree / one MMR of 2^24 leaves over all ranks).  CUDA events on every
rank's stream, max over ranks; one JSON line per measurement on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plonky2_merkle_trees_b200 import _lib, sharded  # noqa: E402


def main():
    world, rank, local_rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    engines = {}
    for mode in ("nccl", "p2p"):
        os.environ["PMT_EXCHANGE"] = mode
        e = sharded.CudaEngine(_lib.Context(local_rank))
        e.comm_init()
        engines[mode] = e
    assert engines["p2p"].peer_memory and not engines["nccl"].peer_memory, "the ranks could not map each other's mailboxes"

    def timed(eng, fn, steps, warm=2):
        stream = torch.cuda.ExternalStream(eng.ctx.stream, device=dev)
        for _ in range(warm):
            fn()
        eng.sync(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        eng.sync(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        return [round(float(x.item()), 4) for x in every]

    cases = []
    lgw = 24
    d_weak = bench.splitmix_torch(rank * (1 << lgw) * 4, (1 << lgw) * 4, dev).view(1 << lgw, 4)
    cases.append(("weak_2p24_per_rank", lambda e: sharded.build_sharded_tree(d_weak, world << lgw, 0, e), 10))
    per = (1 << 24) // world
    d_strong = bench.splitmix_torch(rank * per * 4, per * 4, dev).view(per, 4)
    cases.append(("strong_tree_2p24", lambda e: sharded.build_sharded_tree(d_strong, 1 << 24, 0, e), 20))
    d_mmr = bench.splitmix_torch(7 + rank * per, per, dev)
    cases.append(("strong_mmr_2p24", lambda e: sharded.build_sharded_mmr(d_mmr, 1 << 24, e), 20))
    for name, fn, steps in cases:
        for rep in range(3):
            for mode in ("p2p", "nccl"):
                e = engines[mode]
                ms = timed(e, lambda: fn(e), steps)
                if rank == 0:
                    print(json.dumps({"exchange_ab": name, "world": world, "exchange": mode, "rep": rep, "steps": steps, "ms_per_build": max(ms),
                                      "ms_per_rank": ms}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""BASELINE.json configs that are sharded across the GPUs of one box, at FULL size, under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        tools/bench_configs_multi.py

  C4  FRI-style commitment: 2^20 rows x 135 felt columns, cap_height 4, sharded across the ranks
  C5  2^28 leaves x 4 felts, cap_height 0, subtree-sharded with the NVLink root all_gather
  C3  MMR over 2^24 and 2^24 - 1 single-felt leaves, sharded in balanced rounds (sharded.build_sharded_mmr)

Timing: CUDA events on each rank's ctx stream around the whole collective build (local subtree + all_gather + top), max
over ranks, median of 5 after 2 warm-ups.  Parity at full size: the cap / root / bagged peaks of the sharded build must
equal those of the SAME tree built in one piece on rank 0's GPU (which tests/test_gpu_parity.py ties to the oracle).
One JSON line per config on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plonky2_merkle_trees_b200 import _lib, merkle_tree, mmr, sharded  # noqa: E402
from plonky2_merkle_trees_b200.device import to_host  # noqa: E402


def timed_collective(ctx, dev, fn, reps=5, warm=2):
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    out = None
    for _ in range(warm):
        out = fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); ctx.sync(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = fn()
        e1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return out, sorted(ts)[len(ts) // 2], min(ts)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    ctx = _lib.Context(local_rank)
    eng = sharded.CudaEngine(ctx)
    eng.comm_init()      # the roots are exchanged inside libpmt (peer-memory mailboxes, or ncclAllGather with PMT_EXCHANGE=nccl)
    only = sys.argv[1:] or ["C4", "C5", "C3"]

    def emit(**kw):
        if rank == 0:
            kw["roots_exchange"] = "peer-memory mailboxes (k_exchange_top)" if eng.peer_memory else "ncclAllGather"
            print(json.dumps(kw), flush=True)

    for name, lg, w, h in [("C4", 20, 135, 4), ("C5", 28, 4, 0)]:
        if name not in only:
            continue
        n = 1 << lg
        per = n // world
        d_local = bench.splitmix_torch(rank * per * w, per * w, dev).view(per, w)
        tree, med, best = timed_collective(ctx, dev, lambda: sharded.build_sharded_tree(d_local, n, h, eng))
        cap = to_host(tree.cap)
        del tree
        ok = None
        if rank == 0:     # the same tree in one piece on one GPU (C5: 8 GiB of leaves + 16 GiB of digests)
            torch.cuda.empty_cache()
            d_all = bench.splitmix_torch(0, n * w, dev).view(n, w)
            ref = merkle_tree.MerkleTree.new_dev(d_all, h, ctx)
            ok = bool(np.array_equal(cap, ref.cap))
            del ref, d_all
            torch.cuda.empty_cache()
        perms = (n - (1 << h)) + (n * ((w + 7) // 8) if w > 4 else 0)
        emit(config=name, n_gpus=world, log2_leaves=lg, width=w, cap_height=h, ms_median=med, ms_best=best,
             leaves_per_s=n / (med * 1e-3), Gperm_per_s=perms / (med * 1e-3) / 1e9, cap_equals_single_gpu_build=ok)
        del d_local
        torch.cuda.empty_cache()

    if "C3" in only:
        for n in (1 << 24, (1 << 24) - 1):
            leaves_ranges = sharded.mmr_shard_ranges(n, world, rank)
            d_mine = torch.cat([bench.splitmix_torch(a, c, dev) for a, c in leaves_ranges]) if leaves_ranges else torch.zeros(0, dtype=torch.int64, device=dev)
            sm, med, best = timed_collective(ctx, dev, lambda: sharded.build_sharded_mmr(d_mine, n, eng), reps=3, warm=1)
            bag, peaks = sm.bagging_the_peaks(), sm.get_peaks()
            # every rank proves and verifies (on its GPU) 128 leaves it owns against the replicated peaks / bag
            rnd = np.random.default_rng(rank)
            own = [i for i in rnd.integers(0, n, size=4096).tolist() if sm.owner(i) == rank][:128]
            ok_p = all(sm.get_proof_normal_index(i).verify(int(bench.splitmix_numpy(i, 1)[0]), bag, ctx) for i in own)
            okt = torch.tensor([1 if ok_p else 0], device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            del sm
            ok = None
            if rank == 0:
                torch.cuda.empty_cache()
                ref = mmr.MMR.new(ctx)
                ref.extend_dev(bench.splitmix_torch(0, n, dev))
                ok = bool(np.array_equal(peaks, ref.get_peaks()) and np.array_equal(bag, ref.bagging_the_peaks()))
                del ref
                torch.cuda.empty_cache()
            emit(config="C3-sharded", n_gpus=world, n_leaves=n, rounds=len(sharded.mmr_shard_plan(n, world)[0]), ms_median=med, ms_best=best,
                 leaves_per_s=n / (med * 1e-3), peaks_and_bag_equal_single_gpu_build=ok, proofs_verified_on_every_rank=bool(okt.item()))
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

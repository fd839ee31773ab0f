#!/usr/bin/env python3
"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py

Every rank builds its subtree shard (plonky2_merkle_trees_b200.sharded), the chunks are gathered on rank 0 and compared
BIT FOR BIT with the same tree built in one piece on rank 0's GPU (which tests/test_gpu_parity.py ties to the oracle).
Covers cap_height < log2 G (top levels after the all_gather), cap_height >= log2 G (cap slices) and wide leaves.
Prints one JSON line per case; exit code 1 on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plonky2_merkle_trees_b200 import _lib, merkle_tree, sharded  # noqa: E402
from plonky2_merkle_trees_b200.device import to_host  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import datetime
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)       # a hung collective shows where, and costs 150 s, not the NCCL timeout
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    ctx = _lib.Context(local_rank)
    eng = sharded.CudaEngine(ctx)
    ok_all = True
    cases = [(14, 4, 0), (14, 1, 0), (12, 135, 4), (13, 9, 1), (16, 4, 5), (10, 4, 3)]
    # every case three times: the roots exchanged by torch.distributed's all_gather (plumbing in Python), then inside libpmt on
    # the ctx's own stream (pmt_comm_init + pmt_merkle_tree_build_sharded_dev: one call, no host sync) -- through peer-memory
    # mailboxes (k_exchange_top) where the ranks can map each other, and with PMT_EXCHANGE=nccl through ncclAllGather
    eng_nccl = sharded.CudaEngine(_lib.Context(local_rank))
    os.environ["PMT_EXCHANGE"] = "nccl"
    eng_nccl.comm_init()
    os.environ.pop("PMT_EXCHANGE")
    eng.comm_init()
    assert not eng_nccl.peer_memory
    engines = {"torch": sharded.CudaEngine(ctx), "lib": eng, "lib-nccl": eng_nccl}     # "torch": same ctx, no communicator
    label = {"torch": "torch.distributed all_gather", "lib-nccl": "ncclAllGather inside libpmt",
             "lib": "peer-memory mailboxes inside libpmt" if eng.peer_memory else "ncclAllGather inside libpmt (no peer mapping)"}
    for lg, w, h, how in [c + (m,) for m in engines for c in cases]:
        eng_used = engines[how]
        n = 1 << lg
        per = n // world
        d_local = bench.splitmix_torch(rank * per * w, per * w, dev).view(per, w)
        tree = sharded.build_sharded_tree(d_local, n, h, eng_used)
        eng_used.sync()
        chunk = tree.local_digests.contiguous()
        chunks = [torch.empty_like(chunk) for _ in range(world)] if rank == 0 else None
        dist.gather(chunk, chunks, dst=0)
        if rank == 0:
            full_leaves = bench.splitmix_numpy(0, n * w).reshape(n, w)
            ref = merkle_tree.MerkleTree.new(full_leaves, h, ctx)
            got = tree.assemble_global([to_host(c) for c in chunks])
            cap = to_host(tree.cap)
            ok = bool(np.array_equal(got, ref.digests) and np.array_equal(cap, ref.cap))
            ok_all &= ok
            print(json.dumps({"check": "sharded_vs_single", "world": world, "log2_n": lg, "width": w, "cap_height": h,
                              "roots_exchange": label[how],
                              "digests": int(got.shape[0]), "ok": ok}), flush=True)
    # ---- sharded MMR (balanced rounds + tail) against one MMR built on rank 0's GPU ------------------------------------
    from plonky2_merkle_trees_b200 import mmr
    # the three forms again: torch.distributed's all_gather between library calls, then pmt_mmr_build_sharded_dev (one call)
    for n, how in [(n, m) for m in engines for n in [1 << 14, (1 << 14) - 1, 100100, 77, world]]:
        eng_used = engines[how]
        leaves = bench.splitmix_numpy(7, n)
        rngs = sharded.mmr_shard_ranges(n, world, rank)
        mine = np.concatenate([leaves[a:a + c] for a, c in rngs]) if rngs else np.zeros(0, np.uint64)
        d_mine = torch.from_numpy(mine.view(np.int64)).to(dev)
        sm = sharded.build_sharded_mmr(d_mine, n, eng_used)
        loc = torch.from_numpy(np.ascontiguousarray(sm.local.elements).view(np.int64)).to(dev)
        chunks = [torch.empty_like(loc) for _ in range(world)]
        if loc.numel():
            dist.all_gather(chunks, loc)
        tail_len = torch.tensor([0 if sm.tail is None else sm.tail.elements.shape[0]], device=dev)
        dist.broadcast(tail_len, world - 1)
        tail_el = torch.zeros((int(tail_len.item()), 4), dtype=torch.int64, device=dev)
        if sm.tail is not None:
            tail_el.copy_(torch.from_numpy(sm.tail.elements.view(np.int64)))
        if tail_el.numel():
            dist.broadcast(tail_el, world - 1)
        # every rank proves + verifies (on its GPU) a few leaves it owns
        own = [i for i in range(0, n, max(1, n // 64)) if sm.owner(i) == rank][:16]
        bag = sm.bagging_the_peaks()
        ok_proofs = all(sm.get_proof_normal_index(i).verify(int(leaves[i]), bag, ctx) for i in own)
        okt = torch.tensor([1 if ok_proofs else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if rank == 0:
            ref = mmr.MMR.new(ctx)
            ref.extend(leaves)
            got = sm.assemble_global([to_host(c) for c in chunks], to_host(tail_el))
            ok = bool(np.array_equal(got, ref.elements) and np.array_equal(sm.get_peaks(), ref.get_peaks())
                      and np.array_equal(bag, ref.bagging_the_peaks()) and int(okt.item()) == 1)
            ok_all &= ok
            print(json.dumps({"check": "sharded_mmr_vs_single", "world": world, "n_leaves": n, "rounds": len(sm.rounds),
                              "tail": sm.plan[1],
                              "roots_exchange": label[how], "elements": int(got.shape[0]), "ok": ok}), flush=True)
    flag = torch.tensor([1 if ok_all else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())

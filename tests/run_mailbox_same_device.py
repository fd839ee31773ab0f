#!/usr/bin/env python3
"""Helper of tests/test_gpu_parity.py::test_peer_memory_exchange_on_one_device (run as a subprocess, because the knobs below are
read when CUDA starts): the mailbox exchange of pmt_merkle_tree_build_multi_dev (k_exchange_top: P2P stores + flags + the top
levels in one launch per context) with all contexts on cuda:0, against the oracle and against the copy path."""
import json
import os
import sys

os.environ["PMT_EXCHANGE_SAME_DEVICE"] = "1"          # contexts that share a device normally keep the copy path
os.environ.setdefault("PMT_EXCHANGE_TIMEOUT_MS", "5000")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per stream: a spinning kernel must not block its peer

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import splitmix_felts  # noqa: E402
import oracle  # noqa: E402  (tests/conftest.py puts oracle/ on sys.path: the module is oracle/oracle.py)
from plonky2_merkle_trees_b200 import _lib, merkle_tree, sharded  # noqa: E402
from plonky2_merkle_trees_b200.device import to_device, to_host  # noqa: E402


def main():
    pool = [_lib.Context(0) for _ in range(8)]
    ok_all = True
    cases = [(10, 4, 0, 2), (10, 4, 0, 8), (10, 4, 1, 4), (10, 4, 3, 8), (9, 135, 4, 8), (8, 4, 5, 8), (15, 7, 0, 4), (12, 4, 0, 8), (16, 4, 2, 2)]
    for lg, w, h, G in cases:
        n = 1 << lg
        rows = splitmix_felts(31 * lg + w + h + G, n * w).reshape(n, w)
        odg, ocap = oracle.merkle_tree_new(rows, h, threads=4, fast=True)
        ctxs = pool[:G]
        per = n // G
        d_leaves = [to_device(rows[r * per:(r + 1) * per], "cuda:0") for r in range(G)]
        g = G.bit_length() - 1
        for mode in ("mailbox", "copy"):
            os.environ["PMT_EXCHANGE"] = "nccl" if mode == "copy" else "p2p"
            ctxs[0].profile(True)
            reps = 11 if mode == "mailbox" else 1            # more exchanges than the ring has slots
            for _ in range(reps):
                d_dig, d_roots, d_top, d_cap = merkle_tree.MerkleTree.new_multi_dev(d_leaves, n, h, ctxs)
            ctxs[0].sync()
            for c in ctxs:
                c.sync()
            prof = ctxs[0].profile_read()
            ctxs[0].profile(False)
            used = "k_exchange_top" in prof
            chunks = [to_host(t) for t in d_dig]
            if h >= g:
                t = sharded.ShardedMerkleTree(n, w, h, G, 0, None, None, None, to_host(d_cap))
            else:
                t = sharded.ShardedMerkleTree(n, w, h, G, 0, None, to_host(d_roots), to_host(d_top)[:G - (1 << h)], to_host(d_cap))
            ok = bool(np.array_equal(to_host(d_cap), ocap) and np.array_equal(t.assemble_global(chunks), odg) and used == (mode == "mailbox"))
            ok_all &= ok
            print(json.dumps({"check": "multi_dev_same_device", "log2_n": lg, "width": w, "cap_height": h, "contexts": G, "exchange": mode,
                              "k_exchange_top_launched": used, "ok": ok}), flush=True)
    os.environ["PMT_EXCHANGE"] = "p2p"
    for c in pool:
        c.close()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())

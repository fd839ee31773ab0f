#!/usr/bin/env python3
"""Tuning aid: end-to-end time of the host-buffer pmt_merkle_tree_build (pinned leaves up, every digest down) for several
pipeline depths, next to the raw PCIe copy times of the same buffers.  One JSON line per measurement."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plonky2_merkle_trees_b200 import _lib  # noqa: E402
from plonky2_merkle_trees_b200._lib import u64p  # noqa: E402


def main():
    lg, w = 24, 4
    n = 1 << lg
    ctx = _lib.default_context(0)
    dev = torch.device("cuda", 0)
    h_leaves = torch.empty((n, w), dtype=torch.int64).pin_memory()
    h_leaves.copy_(bench.splitmix_torch(0, n * w, dev).view(n, w))
    h_dig = torch.empty((2 * (n - 1), 4), dtype=torch.int64).pin_memory()
    h_cap = torch.empty((1, 4), dtype=torch.int64).pin_memory()
    d_a = torch.empty_like(h_leaves, device=dev)
    d_b = torch.empty_like(h_dig, device=dev)
    for name, fn in [("h2d_512MiB", lambda: d_a.copy_(h_leaves, non_blocking=True)), ("d2h_1GiB", lambda: h_dig.copy_(d_b, non_blocking=True))]:
        ts = []
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(json.dumps({"copy": name, "ms": 1e3 * min(ts)}), flush=True)
    del d_a, d_b
    torch.cuda.empty_cache()
    # reference digests: the device-resident build (tied to the oracle by tests/test_gpu_parity.py)
    from plonky2_merkle_trees_b200 import merkle_tree
    t = merkle_tree.MerkleTree.new_dev(h_leaves.to(dev), 0, ctx)
    want = t.digests.copy()
    want_cap = t.cap.copy()
    del t
    torch.cuda.empty_cache()
    ref = None
    for chunks in (3, 4, 5, 6):
        os.environ["PMT_PIPELINE_LOG2_CHUNKS"] = str(chunks)
        ts = []
        for _ in range(5):
            ctx.sync(); t0 = time.perf_counter()
            ctx.call("pmt_merkle_tree_build", C.cast(h_leaves.data_ptr(), u64p), n, w, 0, C.cast(h_dig.data_ptr(), u64p), C.cast(h_cap.data_ptr(), u64p))
            ts.append(time.perf_counter() - t0)
        cs = int(np.bitwise_xor.reduce(h_dig.numpy().view(np.uint64).reshape(-1)[::997]))
        ref = cs if ref is None else ref
        print(json.dumps({"e2e_log2_chunks": chunks, "ms_best": 1e3 * min(ts[1:]), "ms_median": 1e3 * sorted(ts[1:])[2], "M_leaves_per_s": n / min(ts[1:]) / 1e6,
                          "same_digests": cs == ref,
                          "equals_device_build": bool(np.array_equal(h_dig.numpy().view(np.uint64), want) and
                                                      np.array_equal(h_cap.numpy().view(np.uint64), want_cap))}), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Workload for the ncu captures of the leaf-read kernels (profiles/k_leaves_*): builds BASELINE config C4 (2^20 rows x 135
felts, cap 4: k_leaves runs the 17-block sponge per row) and one 2^24 x 4 tree (k_leaves_level1: leaf read fused into
level 1) from device-resident inputs.  Run under `ncu --set full -k regex:<kernel> -c 1`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plonky2_merkle_trees_b200 import _lib, merkle_tree  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "wide"
    ctx = _lib.default_context(0)
    dev = torch.device("cuda", 0)
    if which == "wide":
        n, w, h = 1 << 20, 135, 4
    else:
        n, w, h = 1 << 24, 4, 0
    leaves = bench.splitmix_torch(bench.SEED, n * w, dev).view(n, w)
    for _ in range(2):
        t = merkle_tree.MerkleTree.new_dev(leaves, h, ctx)
        ctx.sync()
    print(which, "cap[0] =", t.cap[0])


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""sha256 (first 16 hex digits) of the sources that define the dominant kernel's code (k_level and the thread-per-state
permutation under it).  profiles/ncu_k_level_*.json records it next to the ncu figures; bench.py prints those figures only
while the hash still matches the tree it runs from -- a changed kernel makes them null instead of stale."""
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["goldilocks.cuh", "poseidon.cuh", "poseidon_freq.cuh", "poseidon_freq_constants.cuh", "poseidon_constants.cuh", "merkle_kernels.cuh"]


def kernel_sources_hash(root=ROOT):
    h = hashlib.sha256()
    for f in FILES:
        with open(os.path.join(root, "plonky2_merkle_trees_b200", "csrc", f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(kernel_sources_hash())

#!/usr/bin/env python3
"""sha256 (first 16 hex digits) of the sources that define the dominant kernel's code: the arithmetic headers under the
thread-per-state permutation, and -- from merkle_kernels.cuh -- only the parts k_level<Plonky2> is made of (Digest, digest
loads / stores, two_to_one, the LevelMajor / Plonky2 layouts, the block-size constants, k_level itself), so that work on the
cooperative kernels in the same file does not invalidate the capture.  profiles/ncu_k_level_*.json records it next to the ncu
figures; bench.py prints those figures only while the hash still matches the tree it runs from -- a changed kernel makes them
null instead of stale."""
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["goldilocks.cuh", "poseidon.cuh", "poseidon_freq.cuh", "poseidon_freq_constants.cuh", "poseidon_constants.cuh"]
# (start marker, end marker) pairs inside merkle_kernels.cuh; the end marker is not part of the region
REGIONS = [("struct Digest {", "struct Mmr {"), ("#ifndef PMT_BLOCK", "// level 0: digest(0, k0 + i)"),
           ("// one level: digest(l, k) = two_to_one(children)", "// ----")]


def kernel_sources_hash(root=ROOT):
    h = hashlib.sha256()
    csrc = os.path.join(root, "plonky2_merkle_trees_b200", "csrc")
    for f in FILES:
        with open(os.path.join(csrc, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    with open(os.path.join(csrc, "merkle_kernels.cuh")) as fh:
        text = fh.read()
    for start, end in REGIONS:
        a = text.index(start)
        b = text.index(end, a + len(start))
        h.update(text[a:b].encode())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(kernel_sources_hash())

"""Host mirror of /root/reference/src/simple_merkle_tree/simple_merkle_tree.rs over libpmt (same names, same argument
meaning, same error behaviour: the reference's panics become exceptions).

    MerkleTree.build(leaves)            :28-51   -> count_levels, tree (list of level arrays), root
    tree.get_merkle_proof(i)            :55-74
    tree.get_in_between_hashes(i)       :76-87
    verify_merkle_proof(leaf, i, root, hashes)  :91-109

The tree is built on the GPU and stays resident there (`d_levels`) so batches of proofs are device gathers; the host
copy (`tree`, the reference's Vec<Vec<HashOut>>) is downloaded once, on first access.
"""
import numpy as np
import torch

from . import _lib
from ._lib import PmtError, as_u64
from .device import dev_u64, dptr, to_device, to_host


class MerkleTree:
    def __init__(self, ctx, n, d_levels, d_root):
        self.ctx = ctx
        self.n = n
        self.count_levels = n.bit_length() - 1
        self.d_levels = d_levels      # (2n - 2, 4) level-major on the device
        self.d_root = d_root
        self._levels = None
        self._root = None

    @classmethod
    def build(cls, leaves, ctx=None):
        ctx = ctx or _lib.default_context()
        leaves = as_u64(leaves).reshape(-1)
        n = leaves.size
        if n == 0 or n & (n - 1):
            raise PmtError(_lib.PMT_E_NOT_POW2, "log2_strict: %d leaves is not a power of two (simple_merkle_tree.rs:30)" % n)
        if n < 2:
            raise PmtError(_lib.PMT_E_INVALID_ARG, "MerkleTree::build needs at least 2 leaves (simple_merkle_tree.rs:38)")
        dev = "cuda:%d" % ctx.device
        d_leaves = to_device(leaves, dev)
        d_levels = dev_u64((2 * n - 2, 4), dev)
        d_root = dev_u64((4,), dev)
        ctx.call("pmt_simple_tree_build_dev", dptr(d_leaves), n, dptr(d_levels), dptr(d_root))
        ctx.sync()
        return cls(ctx, n, d_levels, d_root)

    # ---- the reference's public fields -------------------------------------------------------------------------
    @property
    def tree(self):
        """Vec<Vec<HashOut>>: level 0 (n leaf digests) ... level count_levels-1 (2 digests)."""
        if self._levels is None:
            flat = to_host(self.d_levels)
            out, off, m = [], 0, self.n
            while m >= 2:
                out.append(flat[off:off + m])
                off += m
                m //= 2
            self._levels = out
        return self._levels

    @property
    def root(self):
        if self._root is None:
            self._root = to_host(self.d_root)
        return self._root

    # ---- proofs ----------------------------------------------------------------------------------------------------
    def get_merkle_proofs(self, leaf_indices):
        """Batch form of get_merkle_proof: (len(idx), count_levels, 4)."""
        idx = as_u64(leaf_indices).reshape(-1)
        if idx.size and int(idx.max()) >= self.n:
            raise IndexError("assert!(leaf_index < self.tree[0].len()) (simple_merkle_tree.rs:56)")
        dev = self.d_levels.device
        d_idx = to_device(idx, dev)
        d_out = dev_u64((idx.size, self.count_levels, 4), dev)
        self.ctx.call("pmt_simple_tree_prove_dev", dptr(self.d_levels), self.n, dptr(d_idx), idx.size, dptr(d_out))
        self.ctx.sync()
        return to_host(d_out)

    def get_merkle_proof(self, leaf_index):
        return self.get_merkle_proofs([leaf_index])[0]

    def get_in_between_hashes(self, leaf_index):
        if leaf_index >= self.n:
            raise IndexError("assert!(leaf_index < self.tree[0].len()) (simple_merkle_tree.rs:77)")
        index = leaf_index // 2
        hashes = []
        for i in range(1, self.count_levels):
            hashes.append(self.tree[i][index])
            index //= 2
        hashes.append(self.root)
        return np.stack(hashes)


def verify_merkle_proofs(leaves, leaf_indices, root, proofs, ctx=None):
    """Batch verify_merkle_proof against one root: bool array."""
    ctx = ctx or _lib.default_context()
    leaves = as_u64(leaves).reshape(-1)
    idx = as_u64(leaf_indices).reshape(-1)
    proofs = as_u64(proofs)
    proofs = proofs.reshape(idx.size, -1, 4)
    dev = "cuda:%d" % ctx.device
    d_ok = torch.empty(idx.size, dtype=torch.uint8, device=dev)
    d_l, d_i, d_r, d_p = to_device(leaves, dev), to_device(idx, dev), to_device(as_u64(root).reshape(4), dev), to_device(proofs, dev)
    ctx.call("pmt_simple_tree_verify_dev", dptr(d_l), dptr(d_i), idx.size, dptr(d_r), dptr(d_p), proofs.shape[1], dptr(d_ok))
    ctx.sync()
    return d_ok.cpu().numpy().astype(bool)


def verify_merkle_proof(leaf, leaf_index, root, hashes, ctx=None):
    hashes = as_u64(hashes).reshape(1, -1, 4)
    return bool(verify_merkle_proofs([leaf], [leaf_index], root, hashes, ctx)[0])

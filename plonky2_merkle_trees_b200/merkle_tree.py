"""Host mirror of plonky2's `MerkleTree::new(leaves, cap_height)` / `prove` / `verify_merkle_proof_to_cap`
([UPSTREAM] plonky2 v0.1.3 @ 3b21b87d hash/merkle_tree.rs, hash/merkle_proofs.rs -- what the reference's verifier tests
run implicitly inside `circuit_data.prove`, e.g. /root/reference/src/mmr/mmr_plonky2_verifier.rs:148).

Fields keep upstream's names: `leaves`, `digests` (interleaved layout), `cap`.
"""
import numpy as np
import torch

from . import _lib
from ._lib import PmtError, as_u64
from .device import dev_u64, dptr, to_device, to_host


class MerkleTree:
    def __init__(self, ctx, n, width, cap_height, d_leaves, d_digests, d_cap):
        self.ctx, self.n, self.width, self.cap_height = ctx, n, width, cap_height
        self.d_leaves, self.d_digests, self.d_cap = d_leaves, d_digests, d_cap
        self._digests = self._cap = None

    @classmethod
    def new(cls, leaves, cap_height, ctx=None):
        ctx = ctx or _lib.default_context()
        leaves = as_u64(leaves)
        if leaves.ndim != 2:
            raise ValueError("leaves must be (n, width)")
        n, w = leaves.shape
        return cls.new_dev(to_device(leaves, "cuda:%d" % ctx.device), cap_height, ctx)

    @classmethod
    def new_dev(cls, d_leaves, cap_height, ctx=None):
        """d_leaves: (n, width) int64 CUDA tensor holding u64 felts."""
        ctx = ctx or _lib.default_context()
        n, w = d_leaves.shape
        lg = n.bit_length() - 1
        if n == 0 or n & (n - 1):
            raise PmtError(_lib.PMT_E_NOT_POW2, "log2_strict: %d leaves is not a power of two" % n)
        if cap_height > lg:
            raise PmtError(_lib.PMT_E_RANGE, "cap_height=%d should be at most log2(leaves.len())=%d" % (cap_height, lg))
        ncap = 1 << cap_height
        d_digests = dev_u64((2 * (n - ncap), 4), d_leaves.device)
        d_cap = dev_u64((ncap, 4), d_leaves.device)
        ctx.call("pmt_merkle_tree_build_dev", dptr(d_leaves), n, w, cap_height, dptr(d_digests), dptr(d_cap))
        ctx.sync()
        return cls(ctx, n, w, cap_height, d_leaves, d_digests, d_cap)

    @classmethod
    def from_columns_dev(cls, d_columns, cap_height, bit_reverse=True, keep_leaves=False, ctx=None):
        """The commitment the prover builds from its LDE values ([UPSTREAM] plonky2 fri module, PolynomialBatch::from_values):
        d_columns is (width, n) -- one row per polynomial, column-major for the tree -- and the leaves are
        reverse_index_bits(transpose(d_columns)).  The transpose + bit reversal are fused into the leaf kernel; the row-major
        `leaves` are only materialised when keep_leaves is set."""
        ctx = ctx or _lib.default_context()
        w, n = d_columns.shape
        lg = n.bit_length() - 1
        if n == 0 or n & (n - 1):
            raise PmtError(_lib.PMT_E_NOT_POW2, "log2_strict: %d leaves is not a power of two" % n)
        if cap_height > lg:
            raise PmtError(_lib.PMT_E_RANGE, "cap_height=%d should be at most log2(leaves.len())=%d" % (cap_height, lg))
        ncap = 1 << cap_height
        d_digests = dev_u64((2 * (n - ncap), 4), d_columns.device)
        d_cap = dev_u64((ncap, 4), d_columns.device)
        d_leaves = dev_u64((n, w), d_columns.device) if keep_leaves else None
        ctx.call("pmt_merkle_tree_build_from_columns_dev", dptr(d_columns.contiguous()), n, w, 1 if bit_reverse else 0, cap_height,
                 dptr(d_leaves) if keep_leaves else None, dptr(d_digests), dptr(d_cap))
        ctx.sync()
        return cls(ctx, n, w, cap_height, d_leaves, d_digests, d_cap)

    @staticmethod
    def new_multi(leaves, cap_height, ctxs):
        """MerkleTree::new over several GPUs from ONE process (pmt_merkle_tree_build_multi): `ctxs` is a power-of-two list
        of distinct Contexts, normally one per device; context r builds the subtree over its n / len(ctxs) leaves on its
        own device.  Host buffers in, host buffers out: returns upstream's (digests, cap) as numpy arrays, identical to
        MerkleTree.new(leaves, cap_height).digests / .cap."""
        import ctypes as C
        leaves = as_u64(leaves)
        if leaves.ndim != 2:
            raise ValueError("leaves must be (n, width)")
        n, w = leaves.shape
        if not ctxs:
            raise ValueError("new_multi needs at least one Context")
        ncap = 1 << cap_height
        digests = np.zeros((max(2 * (n - ncap), 0), 4), np.uint64)
        cap = np.zeros((ncap, 4), np.uint64)
        handles = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        rc = ctxs[0].lib.pmt_merkle_tree_build_multi(handles, len(ctxs), _lib.ptr(leaves), n, w, cap_height, _lib.ptr(digests),
                                                     _lib.ptr(cap))
        ctxs[0].check(rc)
        return digests, cap

    @staticmethod
    def new_multi_dev(d_leaves_per_ctx, n, cap_height, ctxs):
        """MerkleTree::new over several GPUs from ONE process, device resident (pmt_merkle_tree_build_multi_dev):
        d_leaves_per_ctx[r] = the (n / G, width) leaf rows of context r ON ITS DEVICE.  Returns (digest slices per context,
        d_roots, d_top, d_cap) -- the last three on ctxs[0]'s device; only enqueued: sync ctxs[0] before reading."""
        import ctypes as C
        G = len(ctxs)
        w = d_leaves_per_ctx[0].shape[1]
        g = G.bit_length() - 1
        ncap = 1 << cap_height
        per = n // G
        local_cap = max(1, ncap // G) if cap_height >= g else 1
        d_dig = [dev_u64((2 * (per - local_cap), 4), t.device) for t in d_leaves_per_ctx]
        dev0 = d_leaves_per_ctx[0].device
        d_roots, d_top, d_cap = dev_u64((G, 4), dev0), dev_u64((max(G - ncap, 1), 4), dev0), dev_u64((ncap, 4), dev0)
        handles = (C.c_void_p * G)(*[c.h for c in ctxs])
        leaves = (C.c_void_p * G)(*[t.data_ptr() for t in d_leaves_per_ctx])
        digs = (C.c_void_p * G)(*[t.data_ptr() for t in d_dig])
        from .device import order_after_torch
        for c in ctxs:
            order_after_torch(c)
        rc = ctxs[0].lib.pmt_merkle_tree_build_multi_dev(handles, G, leaves, n, w, cap_height, digs, dptr(d_roots), dptr(d_top), dptr(d_cap))
        ctxs[0].check(rc)
        return d_dig, d_roots, d_top, d_cap

    @property
    def leaves(self):
        return to_host(self.d_leaves)

    @property
    def digests(self):
        if self._digests is None:
            self._digests = to_host(self.d_digests)
        return self._digests

    @property
    def cap(self):
        if self._cap is None:
            self._cap = to_host(self.d_cap)
        return self._cap

    def prove_batch(self, leaf_indices):
        idx = as_u64(leaf_indices).reshape(-1)
        if idx.size and int(idx.max()) >= self.n:
            raise IndexError("leaf_index out of range")
        depth = (self.n.bit_length() - 1) - self.cap_height
        d_idx = to_device(idx, self.d_digests.device)
        d_out = dev_u64((idx.size, depth, 4), self.d_digests.device)
        self.ctx.call("pmt_merkle_prove_dev", dptr(self.d_digests), self.n, self.cap_height, dptr(d_idx), idx.size, dptr(d_out))
        self.ctx.sync()
        return to_host(d_out)

    def prove(self, leaf_index):
        """MerkleProof.siblings for one leaf."""
        return self.prove_batch([leaf_index])[0]


def verify_merkle_proofs_to_cap(leaf_rows, leaf_indices, cap, cap_height, proofs, ctx=None):
    ctx = ctx or _lib.default_context()
    rows = as_u64(leaf_rows)
    idx = as_u64(leaf_indices).reshape(-1)
    rows = rows.reshape(idx.size, -1)
    proofs = as_u64(proofs).reshape(idx.size, -1, 4)
    dev = "cuda:%d" % ctx.device
    d_ok = torch.empty(idx.size, dtype=torch.uint8, device=dev)
    d_r, d_i, d_c, d_p = to_device(rows, dev), to_device(idx, dev), to_device(as_u64(cap).reshape(-1, 4), dev), to_device(proofs, dev)
    ctx.call("pmt_merkle_verify_dev", dptr(d_r), rows.shape[1], dptr(d_i), idx.size, dptr(d_c), cap_height, dptr(d_p),
             proofs.shape[1], dptr(d_ok))
    ctx.sync()
    return d_ok.cpu().numpy().astype(bool)


def verify_merkle_proof_to_cap(leaf_data, leaf_index, cap, cap_height, siblings, ctx=None):
    return bool(verify_merkle_proofs_to_cap(as_u64(leaf_data).reshape(1, -1), [leaf_index], cap, cap_height,
                                            as_u64(siblings).reshape(1, -1, 4), ctx)[0])

"""Host mirror of /root/reference/src/mmr/merkle_mountain_ranges.rs over libpmt.

    MMR.new() / add_leaf(leaf)                    :84-120   (add_leaf == extend([leaf]); `extend` is the batch form)
    mmr.get_peaks() / bagging_the_peaks()          :179-200, :122-127
    mmr.get_proof(mmr_index) / get_proof_normal_index(i)   :203-223
    MMR_proof{mmr_size, merkle_proof, peaks}.verify(leaf, root)   :15-23, :232-252
    get_heights_bitmap_for_mmr_size / get_mmr_index        :39-81, :257-270  (pure index math, host side)

`elements` (the reference's only state, :8-12) lives on the GPU in post-order; `mmr.elements` downloads it.
"""
import numpy as np
import torch

from . import _lib
from ._lib import PmtError, as_u64
from .device import dev_u64, dptr, to_device, to_host


def get_heights_bitmap_for_mmr_size(mmr_size):
    """merkle_mountain_ranges.rs:39-81 -> (peaks bitmap, remainder)."""
    if mmr_size == 0:
        return (0, 0)
    subtree_size = (1 << mmr_size.bit_length()) - 1
    updated, peaks = mmr_size, 0
    while subtree_size > 0:
        peaks <<= 1
        if updated >= subtree_size:
            peaks |= 1
            updated -= subtree_size
        subtree_size >>= 1
    return (peaks, updated)


def get_mmr_index(leaf_normal_index):
    """merkle_mountain_ranges.rs:257-270 (== 2 i - popcount(i)); i32 arithmetic in the reference => i < 2^30."""
    if leaf_normal_index >= 1 << 30:
        raise OverflowError("get_mmr_index: i32 overflow in the reference (merkle_mountain_ranges.rs:264)")
    index, height, res = leaf_normal_index, 1, 0
    while index > 0:
        if index & 1:
            res += (1 << height) - 1
        height += 1
        index >>= 1
    return res


def _normal_index(mmr_index):
    """inverse of get_mmr_index for leaf positions; raises for positions that are not leaves."""
    lo, hi = 0, mmr_index + 1
    while lo < hi:
        mid = (lo + hi) // 2
        if 2 * mid - bin(mid).count("1") < mmr_index:
            lo = mid + 1
        else:
            hi = mid
    if 2 * lo - bin(lo).count("1") != mmr_index:
        raise ValueError("mmr_index %d is not a leaf position" % mmr_index)
    return lo


class MMR_proof:
    def __init__(self, mmr_size, merkle_proof, peaks):
        self.mmr_size = mmr_size          # elements.len() when the proof was made
        self.merkle_proof = merkle_proof  # list of (digest (4,) u64, sibling_on_left: bool)
        self.peaks = peaks                # (k, 4) u64

    def verify(self, leaf, root, ctx=None):
        """merkle_mountain_ranges.rs:232-252.  Raises AssertionError where the reference's assert! (:245) panics."""
        ctx = ctx or _lib.default_context()
        sib = np.zeros((1, 32, 4), np.uint64)
        left = np.zeros((1, 32), np.uint8)
        for j, (d, on_left) in enumerate(self.merkle_proof):
            sib[0, j] = d
            left[0, j] = 1 if on_left else 0
        status = verify_batch([leaf], sib, left, [len(self.merkle_proof)], self.peaks, root, ctx)[0]
        if status < 0:
            raise AssertionError("assert!(self.peaks.contains(&next_hash)) (merkle_mountain_ranges.rs:245)")
        return bool(status)


def verify_batch(leaves, siblings, on_left, path_len, peaks, root, ctx=None):
    """status per proof: 1 true, 0 false, -1 = the reference would panic."""
    ctx = ctx or _lib.default_context()
    dev = "cuda:%d" % ctx.device
    leaves = as_u64(leaves).reshape(-1)
    n = leaves.size
    d_sib = to_device(as_u64(siblings).reshape(n, 32, 4), dev)
    d_left = torch.from_numpy(np.ascontiguousarray(np.asarray(on_left, dtype=np.uint8)).reshape(n, 32)).to(dev)
    d_len = torch.from_numpy(np.ascontiguousarray(np.asarray(path_len, dtype=np.uint32)).view(np.int32)).to(dev)
    peaks = as_u64(peaks).reshape(-1, 4)
    d_peaks, d_root, d_leaves = to_device(peaks, dev), to_device(as_u64(root).reshape(4), dev), to_device(leaves, dev)
    d_status = torch.empty(n, dtype=torch.int8, device=dev)
    ctx.call("pmt_mmr_verify_dev", dptr(d_leaves), n, dptr(d_sib), dptr(d_left), dptr(d_len), dptr(d_peaks), peaks.shape[0],
             dptr(d_root), dptr(d_status))
    ctx.sync()
    return d_status.cpu().numpy()


def multi_plan(n_before, m, n_ctx):
    """the cut pmt_mmr_extend_multi makes (pure index math in libpmt, no device): None if one context does the whole
    append, else (log2_block, first_aligned, last_aligned)."""
    import ctypes as C
    lib = _lib.load()
    b, a, z = C.c_uint32(0), C.c_size_t(0), C.c_size_t(0)
    if not lib.pmt_mmr_multi_plan(n_before, m, n_ctx, C.byref(b), C.byref(a), C.byref(z)):
        return None
    return b.value, a.value, z.value


def extend_multi(elements, n_before, new_leaves, ctxs):
    """Batch append over several GPUs from ONE process (pmt_mmr_extend_multi): `elements` is the HOST post-order array of the
    MMR with n_before leaves (or None), `ctxs` a list of distinct Contexts, normally one per device.  Returns the host
    array of the MMR with the new leaves appended -- identical to MMR.extend on one device."""
    import ctypes as C
    leaves = as_u64(new_leaves).reshape(-1)
    m = leaves.size
    s0 = 2 * n_before - bin(n_before).count("1")
    s1 = 2 * (n_before + m) - bin(n_before + m).count("1")
    out = np.zeros((s1, 4), np.uint64)
    if n_before:
        out[:s0] = as_u64(elements).reshape(-1, 4)[:s0]
    if m == 0:
        return out
    handles = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    rc = ctxs[0].lib.pmt_mmr_extend_multi(handles, len(ctxs), _lib.ptr(out), n_before, _lib.ptr(leaves), m)
    ctxs[0].check(rc)
    return out


class MMR:
    GROW = 1 << 12  # minimum capacity, in leaves

    def __init__(self, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.dev = "cuda:%d" % self.ctx.device
        self.n_leaves = 0
        self._cap_leaves = 0
        self.d_elements = None
        self._pending = False     # an append was only enqueued (extend_dev(sync=False)): sync before torch reads d_elements

    @classmethod
    def new(cls, ctx=None):
        return cls(ctx)

    @classmethod
    def adopt(cls, ctx, d_elements, n_leaves, pending=True):
        """an MMR over device elements some other call produced (the sharded build): (>= mmr_size(n_leaves), 4) rows"""
        m = cls(ctx)
        m.d_elements, m.n_leaves, m._pending = d_elements, n_leaves, pending
        m._cap_leaves = d_elements.shape[0] // 2
        return m

    def __len__(self):
        return 2 * self.n_leaves - bin(self.n_leaves).count("1")

    def _settle(self):
        if self._pending:
            self.ctx.sync()
            self._pending = False

    @property
    def elements(self):
        """Vec<HashOut>: (len, 4) u64, post-order."""
        if self.n_leaves == 0:
            return np.zeros((0, 4), np.uint64)
        self._settle()
        return to_host(self.d_elements[:len(self)])

    def _reserve(self, n_total):
        if n_total <= self._cap_leaves:
            return
        self._settle()
        cap = max(self.GROW, 1 << (n_total - 1).bit_length())
        new = dev_u64((2 * cap, 4), self.dev)
        if self.n_leaves:
            new[:len(self)].copy_(self.d_elements[:len(self)])
        self.d_elements, self._cap_leaves = new, cap

    def extend(self, leaves):
        """Append a batch of single-felt leaves: identical result to calling add_leaf once per leaf."""
        leaves = as_u64(leaves).reshape(-1)
        if leaves.size == 0:
            return
        return self.extend_dev(to_device(leaves, self.dev))

    def extend_dev(self, d_leaves, sync=True):
        """sync=False only enqueues the append on the ctx stream (the sharded build chains further device work behind it)"""
        m = d_leaves.numel()
        if self.n_leaves + m > 1 << 30:
            raise PmtError(_lib.PMT_E_RANGE, "MMR: more than 2^30 leaves (get_mmr_index is i32, merkle_mountain_ranges.rs:264)")
        self._reserve(self.n_leaves + m)
        self.ctx.call("pmt_mmr_extend_dev", dptr(self.d_elements), self.n_leaves, dptr(d_leaves), m)
        if sync:
            self.ctx.sync()
        self._pending = not sync
        self.n_leaves += m

    def get_peaks_dev(self, out=None):
        """get_peaks into a device tensor (popcount(n_leaves), 4), enqueued on the ctx stream, no host sync"""
        k = bin(self.n_leaves).count("1")
        if out is None:
            out = dev_u64((max(k, 1), 4), self.dev)
        if k:
            import ctypes as C
            kk = C.c_uint32(0)
            self.ctx.call("pmt_mmr_peaks_dev", dptr(self.d_elements), self.n_leaves, dptr(out), C.byref(kk))
        return out[:k]

    def add_leaf(self, leaf):
        self.extend([leaf])

    def get_peaks(self):
        if self.n_leaves == 0:
            return np.zeros((0, 4), np.uint64)
        d_peaks = dev_u64((64, 4), self.dev)
        import ctypes as C
        k = C.c_uint32(0)
        self.ctx.call("pmt_mmr_peaks_dev", dptr(self.d_elements), self.n_leaves, dptr(d_peaks), C.byref(k))
        self.ctx.sync()
        return to_host(d_peaks[:k.value])

    def bagging_the_peaks(self):
        if self.n_leaves == 0:
            raise PmtError(_lib.PMT_E_INVALID_ARG, "bagging_the_peaks on an empty MMR")
        d_root = dev_u64((4,), self.dev)
        self.ctx.call("pmt_mmr_bag_dev", dptr(self.d_elements), self.n_leaves, dptr(d_root))
        self.ctx.sync()
        return to_host(d_root)

    def prove_batch(self, normal_indices):
        """-> siblings (q, 32, 4) u64, on_left (q, 32) u8, path_len (q,) u32 for NORMAL leaf indices."""
        idx = as_u64(normal_indices).reshape(-1)
        if idx.size and int(idx.max()) >= self.n_leaves:
            raise IndexError("leaf index out of range")
        q = idx.size
        d_idx = to_device(idx, self.dev)
        d_sib = torch.zeros((q, 32, 4), dtype=torch.int64, device=self.dev)
        d_left = torch.zeros((q, 32), dtype=torch.uint8, device=self.dev)
        d_len = torch.zeros(q, dtype=torch.int32, device=self.dev)
        self.ctx.call("pmt_mmr_prove_dev", dptr(self.d_elements), self.n_leaves, dptr(d_idx), q, dptr(d_sib), dptr(d_left), dptr(d_len))
        self.ctx.sync()
        return to_host(d_sib), d_left.cpu().numpy(), d_len.cpu().numpy().view(np.uint32)

    def get_proof_normal_index(self, normal_index):
        sib, left, ln = self.prove_batch([normal_index])
        path = [(sib[0, j].copy(), bool(left[0, j])) for j in range(int(ln[0]))]
        return MMR_proof(len(self), path, self.get_peaks())

    def get_proof(self, mmr_index):
        return self.get_proof_normal_index(_normal_index(mmr_index))

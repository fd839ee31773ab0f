"""Subtree-sharded multi-GPU build of plonky2's MerkleTree (one process per GPU, torch.distributed for the plumbing).

Rank r of G = 2^g owns leaves [r n/G, (r+1) n/G) = one height-(log2 n - g) subtree and builds it completely on its own
GPU; nothing but subtree roots ever crosses NVLink:

  cap_height h >= g : every rank yields 2^(h-g) cap entries; one all_gather of 32 * 2^(h-g) bytes per rank.  The rank's
                      digests ARE the slice [r D/G, (r+1) D/G) of upstream's `digests` -- no reshuffle.
  cap_height h <  g : every rank yields one root; all_gather of 32 bytes per rank, then the g - h top levels
                      (<= G - 1 two_to_one) are computed redundantly on every rank (pmt_top_levels_dev).  In upstream's
                      layout each rank's digests are still one contiguous chunk; the 2G - 2^(h+1) top digests sit
                      between the chunks at the closed-form positions of `global_digest_index`.

The collective is latency-bound (<= 512 bytes); it is a plain NCCL all_gather because there is no compute to overlap
it with -- the exchange happens once, after the last local level.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .device import dev_u64, dptr


def log2_strict(n):
    if n <= 0 or n & (n - 1):
        raise _lib.PmtError(_lib.PMT_E_NOT_POW2, "log2_strict: %d is not a power of two" % n)
    return n.bit_length() - 1


def shard_range(n, world, rank):
    """leaves [start, start + count) owned by `rank`."""
    log2_strict(n), log2_strict(world)
    if world > n:
        raise _lib.PmtError(_lib.PMT_E_RANGE, "more ranks (%d) than leaves (%d)" % (world, n))
    return rank * (n // world), n // world


def digest_index(level, k):
    """index of node (level, k) inside ONE cap subtree's digest slice [UPSTREAM hash/merkle_tree.rs prove()]."""
    return 2 * (((k >> 1) << (level + 1)) + (1 << level) - 1) + (k & 1)


def global_digest_index(n, cap_height, level, k):
    """index into upstream's global `digests` of node (level, k), level < log2 n - cap_height."""
    L = log2_strict(n) - cap_height
    per = L - level
    c, kk = k >> per, k & ((1 << per) - 1)
    return c * ((2 << L) - 2) + digest_index(level, kk)


class CudaEngine:
    """The product engine: libpmt on this rank's GPU."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.device = torch.device("cuda:%d" % ctx.device)

    def build_local(self, d_leaves, cap_height):
        n, w = d_leaves.shape
        ncap = 1 << cap_height
        d_digests = dev_u64((2 * (n - ncap), 4), self.device)
        d_cap = dev_u64((ncap, 4), self.device)
        self.ctx.call("pmt_merkle_tree_build_dev", dptr(d_leaves), n, w, cap_height, dptr(d_digests), dptr(d_cap))
        return d_digests, d_cap

    def top_levels(self, d_roots, cap_height):
        g = d_roots.shape[0]
        d_top = dev_u64((max(g - (1 << cap_height), 1), 4), self.device)
        self.ctx.call("pmt_top_levels_dev", dptr(d_roots), g, cap_height, dptr(d_top))
        return d_top[:g - (1 << cap_height)]

    def sync(self):
        self.ctx.sync()


class ShardedMerkleTree:
    """Result on one rank: its chunk of `digests`, the replicated top digests and the replicated cap."""

    def __init__(self, n, width, cap_height, world, rank, local_digests, roots, top, cap):
        self.n, self.width, self.cap_height, self.world, self.rank = n, width, cap_height, world, rank
        self.local_digests = local_digests   # (2 (n/G - max(1, 2^(h-g))), 4)
        self.roots = roots                   # (G, 4) gathered subtree roots (None when h >= g)
        self.top = top                       # level-major digests above the roots, incl. the cap (None when h >= g)
        self.cap = cap                       # (2^h, 4)

    def local_offset(self):
        """position of this rank's chunk inside upstream's global `digests`."""
        g = log2_strict(self.world)
        if self.cap_height >= g:
            total = 2 * (self.n - (1 << self.cap_height))
            return self.rank * (total // self.world)
        return global_digest_index(self.n, self.cap_height, 0, self.rank * (self.n // self.world))

    def assemble_global(self, gathered_local):
        """host-side helper (tests / small trees): upstream's full `digests` from every rank's chunk (list of numpy)."""
        g = log2_strict(self.world)
        total = 2 * (self.n - (1 << self.cap_height))
        out = np.zeros((total, 4), np.uint64)
        if self.cap_height >= g:
            return np.concatenate(gathered_local, axis=0) if total else out
        per_rank = self.n // self.world
        Lr = log2_strict(per_rank)
        for r, chunk in enumerate(gathered_local):
            off = global_digest_index(self.n, self.cap_height, 0, r * per_rank)
            out[off:off + chunk.shape[0]] = chunk
        roots = self.roots.cpu().numpy().view(np.uint64) if torch.is_tensor(self.roots) else self.roots
        top = self.top.cpu().numpy().view(np.uint64) if torch.is_tensor(self.top) else self.top
        L = log2_strict(self.n) - self.cap_height
        level_nodes, level = roots, Lr
        pos = 0
        while level < L:
            for k in range(level_nodes.shape[0]):
                out[global_digest_index(self.n, self.cap_height, level, k)] = level_nodes[k]
            m = level_nodes.shape[0] // 2
            level_nodes = top[pos:pos + m]
            pos += m
            level += 1
        return out


def build_sharded_tree(d_local_leaves, n_total, cap_height, engine, group=None):
    """Collective over `group`: every rank passes its (n_total / world, width) leaf shard."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lg, g = log2_strict(n_total), log2_strict(world)
    if cap_height > lg:
        raise _lib.PmtError(_lib.PMT_E_RANGE, "cap_height=%d should be at most log2(leaves.len())=%d" % (cap_height, lg))
    start, count = shard_range(n_total, world, rank)
    if d_local_leaves.shape[0] != count:
        raise ValueError("rank %d: expected %d leaf rows, got %d" % (rank, count, d_local_leaves.shape[0]))
    width = d_local_leaves.shape[1]
    local_h = max(cap_height - g, 0)
    d_digests, d_local_cap = engine.build_local(d_local_leaves, local_h)
    if world == 1:
        engine.sync()
        return ShardedMerkleTree(n_total, width, cap_height, 1, 0, d_digests, None, None, d_local_cap)
    gathered = torch.empty((world * d_local_cap.shape[0], 4), dtype=d_local_cap.dtype, device=d_local_cap.device)
    engine.sync()  # the local cap must be complete before NCCL (a different stream) reads it
    dist.all_gather_into_tensor(gathered, d_local_cap.contiguous(), group=group)
    if cap_height >= g:
        return ShardedMerkleTree(n_total, width, cap_height, world, rank, d_digests, None, None, gathered)
    d_top = engine.top_levels(gathered, cap_height)
    engine.sync()
    ncap = 1 << cap_height
    return ShardedMerkleTree(n_total, width, cap_height, world, rank, d_digests, gathered, d_top, d_top[d_top.shape[0] - ncap:])

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
        self.has_comm = False
        self.peer_memory = False      # set by comm_init: the ranks exchange through peer-memory mailboxes (k_exchange_top)

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

    def top_levels_batch(self, d_roots):
        """d_roots: (batch, G, 4) -> (batch, G - 1, 4): the levels above every set of G roots, one launch."""
        b, g = d_roots.shape[0], d_roots.shape[1]
        d_top = dev_u64((b, max(g - 1, 1), 4), self.device)
        self.ctx.call("pmt_top_levels_batch_dev", dptr(d_roots), b, g, 0, dptr(d_top))
        return d_top[:, :g - 1]

    def sync(self):
        self.ctx.sync()

    def comm_init(self, group=None):
        """Give the ctx its own communicator (pmt_comm_init): rank 0 creates the unique id, torch.distributed carries it to the
        other ranks (plumbing), every rank joins and maps its peers' mailboxes (CUDA IPC).  After this build_sharded_tree runs
        as ONE library call on ONE stream: local build -> k_exchange_top (the roots stored straight into the peers' memory,
        the top levels in the same launch); where peers cannot map each other: ncclAllGather -> top levels."""
        import ctypes as C
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        buf = (C.c_char * 128)()
        if rank == 0:
            self.ctx.call("pmt_nccl_unique_id", C.cast(buf, C.c_void_p))
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=self.device)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(t.cpu().tolist())
        self.ctx.call("pmt_comm_init", C.cast(C.create_string_buffer(raw, 128), C.c_void_p), rank, world)
        self.has_comm = True
        self.peer_memory = bool(self.ctx.lib.pmt_comm_uses_peer_memory(self.ctx.h))

    def build_sharded(self, d_local_leaves, n_total, cap_height, world):
        """pmt_merkle_tree_build_sharded_dev -> (local digests, roots or None, top or None, cap); enqueued only"""
        per, w = d_local_leaves.shape
        g = world.bit_length() - 1
        ncap = 1 << cap_height
        local_cap = ncap // world if cap_height >= g else 1
        d_dig = dev_u64((2 * (per - local_cap), 4), self.device)
        d_roots, d_top = dev_u64((world, 4), self.device), dev_u64((max(world - ncap, 1), 4), self.device)
        d_cap = dev_u64((ncap, 4), self.device)
        self.ctx.call("pmt_merkle_tree_build_sharded_dev", dptr(d_local_leaves), n_total, w, cap_height, dptr(d_dig), dptr(d_roots),
                      dptr(d_top), dptr(d_cap))
        if cap_height >= g:
            return d_dig, None, None, d_cap
        return d_dig, d_roots, d_top[:world - ncap], d_cap

    def build_sharded_mmr(self, d_local_leaves, n_total, world, rank):
        """pmt_mmr_build_sharded_dev -> (local MMR, tail MMR or None, roots (rounds, G, 4), tops (rounds, G - 1, 4), peaks);
        enqueued only.  The plan comes from the library too (pmt_mmr_shard_plan)."""
        import ctypes as C
        from .mmr import MMR
        k, ms, t = C.c_uint32(0), (C.c_size_t * 64)(), C.c_size_t(0)
        self.ctx.check(self.ctx.lib.pmt_mmr_shard_plan(n_total, world, C.byref(k), ms, C.byref(t)))
        k, t = k.value, t.value
        n_main, tp = sum(ms[:k]), bin(t).count("1")
        slots = k + tp
        last = rank == world - 1
        d_local = dev_u64((max(mmr_size(n_main), 1), 4), self.device)
        d_tail = dev_u64((max(mmr_size(t), 1), 4), self.device) if (t and last) else None
        d_gathered = dev_u64((world, max(slots, 1), 4), self.device)
        d_tops = dev_u64((max(k, 1), max(world - 1, 1), 4), self.device)
        d_peaks = dev_u64((max(slots, 1), 4), self.device)
        self.ctx.call("pmt_mmr_build_sharded_dev", dptr(d_local_leaves), n_total, dptr(d_local), dptr(d_tail) if d_tail is not None else None,
                      dptr(d_gathered), dptr(d_tops), dptr(d_peaks))
        local = MMR.adopt(self.ctx, d_local, n_main)
        tail = MMR.adopt(self.ctx, d_tail, t) if d_tail is not None else None
        # (rounds, G, 4) view of the gathered matrix: round i = column i; no copy, the host reads it only on demand
        d_roots = d_gathered[:, :k].permute(1, 0, 2)
        return local, tail, d_roots, d_tops[:k, :world - 1], d_peaks[:slots]

    def publish(self):
        """order torch's current stream (the one NCCL synchronises with) after the ctx stream: no host round trip"""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(torch.cuda.ExternalStream(self.ctx.stream, device=self.device))


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
    if world > 1 and getattr(engine, "has_comm", False):      # NCCL inside libpmt: one call, one stream, no host sync
        d_digests, d_roots, d_top, d_cap = engine.build_sharded(d_local_leaves, n_total, cap_height, world)
        return ShardedMerkleTree(n_total, width, cap_height, world, rank, d_digests, d_roots, d_top, d_cap)
    local_h = max(cap_height - g, 0)
    d_digests, d_local_cap = engine.build_local(d_local_leaves, local_h)
    if world == 1:
        engine.sync()
        return ShardedMerkleTree(n_total, width, cap_height, 1, 0, d_digests, None, None, d_local_cap)
    gathered = torch.empty((world * d_local_cap.shape[0], 4), dtype=d_local_cap.dtype, device=d_local_cap.device)
    # the local cap must be complete before NCCL (a different stream) reads it
    engine.publish() if hasattr(engine, "publish") else engine.sync()
    dist.all_gather_into_tensor(gathered, d_local_cap.contiguous(), group=group)
    if cap_height >= g:
        return ShardedMerkleTree(n_total, width, cap_height, world, rank, d_digests, None, None, gathered)
    d_top = engine.top_levels(gathered, cap_height)
    engine.sync()
    ncap = 1 << cap_height
    return ShardedMerkleTree(n_total, width, cap_height, world, rank, d_digests, gathered, d_top, d_top[d_top.shape[0] - ncap:])


# ---------------------------------------------------------------------------------------------------------------------
# Subtree-sharded MMR (/root/reference/src/mmr/merkle_mountain_ranges.rs)
# ---------------------------------------------------------------------------------------------------------------------
def mmr_size(n_leaves):
    return 2 * n_leaves - bin(n_leaves).count("1")


def mmr_pos(level, k):
    """post-order position of node (level, k) in `elements` (closed form of add_leaf, :89-120)."""
    last = ((k + 1) << level) - 1
    return 2 * last - bin(last).count("1") + level


def _cuda_mmr_engine_methods():
    # CudaEngine grows the three MMR calls the sharded form needs; kept here so the plonky2-tree engine above stays small
    from . import hasher, mmr as mmr_mod

    def new_mmr(self):
        return mmr_mod.MMR(self.ctx)

    def hash_or_noop(self, felts):
        return hasher.hash_or_noop(np.asarray(felts, dtype=np.uint64).reshape(1, -1), self.ctx)[0]

    CudaEngine.new_mmr = new_mmr
    CudaEngine.hash_or_noop = hash_or_noop


_cuda_mmr_engine_methods()


def _host(x):
    """numpy uint64 view of a digest array that may still live on a device"""
    if torch.is_tensor(x):
        return x.detach().cpu().numpy().view(np.uint64)
    return np.asarray(x, dtype=np.uint64)


class ShardedMMR:
    """One rank's part of an MMR over n leaves, G = world size (a power of two).

    Plan (mmr_shard_plan): every set bit 2^b >= G of n is one mountain, cut into G equal perfect sub-mountains of
    m_i = 2^b / G leaves, one per rank ("round" i); the bits below G (< G leaves in total) are the tail, kept by the last
    rank.  So rank r owns the leaves  [S_i + r m_i, S_i + (r + 1) m_i)  of every round (S_i = leaves before round i) --
    perfectly balanced for any n -- and its local MMR over those leaves, fed in round order, has exactly these
    sub-mountains as its own mountains (m_1 > m_2 > ...): ONE batch append builds them all.

      local      the rank's MMR; sub-mountain i is the contiguous slice of the global post-order `elements` that starts at
                 mmr_size(S_i + r m_i)  (popcounts add: S_i, r m_i and the local index occupy disjoint bit ranges)
      rounds     per round: m_i, the G gathered sub-mountain roots and the log2 G levels above them (replicated;
                 global positions mmr_pos(level, S_i / 2^level + k))
      tail       MMR over the last < G leaves (last rank only)
      peaks      global get_peaks(): one per round, then the tail's (replicated)

    The build leaves roots, tops and peaks ON THE DEVICE (no host round trip inside build_sharded_mmr); `rounds` and `peaks`
    download them on first use.
    """

    def __init__(self, n_total, world, rank, plan, local, d_roots, d_tops, tail, d_peaks, engine):
        self.n_total, self.world, self.rank, self.plan = n_total, world, rank, plan
        self.local, self.tail, self.engine = local, tail, engine
        self.d_roots, self.d_tops, self.d_peaks = d_roots, d_tops, d_peaks      # (rounds, G, 4), (rounds, G - 1, 4), (peaks, 4)
        self._rounds = self._peaks = None

    def __len__(self):
        return mmr_size(self.n_total)

    @property
    def rounds(self):
        if self._rounds is None:
            if hasattr(self.engine, "sync"):
                self.engine.sync()
            roots, tops = _host(self.d_roots), _host(self.d_tops)
            self._rounds = [(m, roots[i], np.ascontiguousarray(tops[i])) for i, m in enumerate(self.plan[0])]
        return self._rounds

    @property
    def peaks(self):
        if self._peaks is None:
            if hasattr(self.engine, "sync"):
                self.engine.sync()
            self._peaks = np.ascontiguousarray(_host(self.d_peaks))
        return self._peaks

    def get_peaks(self):
        return self.peaks

    def bagging_the_peaks(self):
        """merkle_mountain_ranges.rs:122-127 on the replicated peaks: the same digest on every rank."""
        return self.engine.hash_or_noop(self.peaks.reshape(-1))

    def locate(self, normal_index):
        """-> (rank, round or None for the tail, index local to that rank's MMR / tail MMR)."""
        start, before = 0, 0
        for i, m in enumerate(self.plan[0]):
            if normal_index < start + self.world * m:
                r, x = divmod(normal_index - start, m)
                return r, i, before + x
            start += self.world * m
            before += m
        if normal_index >= self.n_total:
            raise IndexError("leaf index out of range")
        return self.world - 1, None, normal_index - start

    def owner(self, normal_index):
        return self.locate(normal_index)[0]

    def get_proof_normal_index(self, normal_index):
        """merkle_mountain_ranges.rs:203-223 for a leaf this rank owns -> MMR_proof (same fields as the reference)."""
        from .mmr import MMR_proof
        r, rnd, x = self.locate(normal_index)
        if r != self.rank:
            raise IndexError("leaf %d lives on rank %d" % (normal_index, r))
        src = self.tail if rnd is None else self.local
        sib, left, ln = src.prove_batch([x])
        path = [(sib[0, j].copy(), bool(left[0, j])) for j in range(int(ln[0]))]
        if rnd is not None:
            m, roots, top = self.rounds[rnd]
            k = self.rank                       # index of the sub-mountain among the round's G roots
            level_nodes, pos = roots, 0
            for _ in range(log2_strict(self.world)):
                path.append((np.array(level_nodes[k ^ 1], dtype=np.uint64), bool(k & 1)))
                cnt = level_nodes.shape[0] // 2
                level_nodes = top[pos:pos + cnt]
                pos += cnt
                k >>= 1
        return MMR_proof(len(self), path, self.peaks)

    def assemble_global(self, gathered_local, tail_elements):
        """host-side helper (tests / small MMRs): the reference's full `elements` from every rank's local elements."""
        out = np.zeros((len(self), 4), np.uint64)
        start, before = 0, 0     # global leaves / local leaves before the round
        for m, roots, top in self.rounds:
            lg_m = log2_strict(m)
            for r, chunk in enumerate(gathered_local):
                src = mmr_size(before)
                dst = mmr_size(start + r * m)
                out[dst:dst + 2 * m - 1] = chunk[src:src + 2 * m - 1]
            pos, cnt, level = 0, self.world // 2, lg_m + 1
            while cnt >= 1:
                for k in range(cnt):
                    out[mmr_pos(level, (start >> level) + k)] = top[pos + k]
                pos += cnt
                cnt //= 2
                level += 1
            start += self.world * m
            before += m
        if tail_elements is not None and tail_elements.shape[0]:
            off = mmr_size(start)
            out[off:off + tail_elements.shape[0]] = tail_elements
        return out


def mmr_shard_plan(n_total, world):
    """-> ([m_1, m_2, ...], t): sub-mountain sizes per round (decreasing powers of two) and the tail length (< world)."""
    log2_strict(world)
    ms, rest = [], n_total
    while rest >= world:
        m = 1 << ((rest // world).bit_length() - 1)
        ms.append(m)
        rest -= world * m
    return ms, rest


def mmr_shard_ranges(n_total, world, rank):
    """global leaf ranges [(start, count), ...] that `rank` owns, in the order it must feed them to build_sharded_mmr."""
    ms, t = mmr_shard_plan(n_total, world)
    out, start = [], 0
    for m in ms:
        out.append((start + rank * m, m))
        start += world * m
    if t and rank == world - 1:
        out.append((start, t))
    return out


def _extend_nosync(m, d_leaves):
    try:
        m.extend_dev(d_leaves, sync=False)      # the CUDA MMR: enqueue only
    except TypeError:
        m.extend_dev(d_leaves)                  # test doubles


def _peaks_into(m, out):
    """the peaks of MMR m into the device rows `out` (k, 4) without a host round trip where the engine can"""
    if out.shape[0] == 0:
        return
    if hasattr(m, "get_peaks_dev"):
        m.get_peaks_dev(out)
    else:
        out.copy_(torch.from_numpy(np.ascontiguousarray(m.get_peaks()).view(np.int64)))


def build_sharded_mmr(d_local_leaves, n_total, engine, group=None):
    """Collective over `group`: rank r passes the leaves of mmr_shard_ranges(n_total, world, r), concatenated, as single
    felts.  Equivalent to n_total calls of MMR::add_leaf (:89-120) on one machine.

    Everything stays on the device and nothing waits for the host: one batch append builds all of the rank's sub-mountains,
    their roots (= the local peaks) and the tail's peaks are gathered by a device kernel into one buffer, ONE all_gather moves
    them, one batched launch finishes the log2 G levels above every round's roots (pmt_top_levels_batch_dev).  The host
    sees digests only when it asks (ShardedMMR.peaks / .rounds)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ms, t = mmr_shard_plan(n_total, world)
    n_main = sum(ms)
    want = n_main + (t if rank == world - 1 else 0)
    d_local_leaves = d_local_leaves.reshape(-1)
    if d_local_leaves.numel() != want:
        raise ValueError("rank %d: expected %d leaves, got %d" % (rank, want, d_local_leaves.numel()))
    if world > 1 and getattr(engine, "has_comm", False):      # NCCL inside libpmt: one call, one stream, no host sync
        local, tail, d_roots, d_tops, d_peaks = engine.build_sharded_mmr(d_local_leaves, n_total, world, rank)
        return ShardedMMR(n_total, world, rank, (ms, t), local, d_roots, d_tops, tail, d_peaks, engine)
    dev = d_local_leaves.device
    k, n_tail_peaks = len(ms), bin(t).count("1")
    slots = max(k + n_tail_peaks, 1)
    mine = torch.zeros((slots, 4), dtype=torch.int64, device=dev)     # [k sub-mountain roots | the tail's peaks (last rank)]
    local = engine.new_mmr()
    if n_main:
        _extend_nosync(local, d_local_leaves[:n_main])
        _peaks_into(local, mine[:k])
    tail = None
    if t and rank == world - 1:
        tail = engine.new_mmr()
        _extend_nosync(tail, d_local_leaves[n_main:])
        _peaks_into(tail, mine[k:k + n_tail_peaks])
    if world > 1:
        if hasattr(engine, "publish"):
            engine.publish()                    # NCCL (torch's stream) after the ctx stream: an event, not a host sync
        gathered = torch.empty((world, slots, 4), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered.view(world * slots, 4), mine, group=group)
        d_roots = gathered[:, :k].permute(1, 0, 2).contiguous()        # (rounds, world, 4)
        if k and hasattr(engine, "top_levels_batch"):                    # all rounds' finishes in one launch
            d_tops = engine.top_levels_batch(d_roots)
        elif k:
            d_tops = torch.stack([engine.top_levels(d_roots[i], 0) for i in range(k)])
        else:
            d_tops = torch.zeros((0, world - 1, 4), dtype=torch.int64, device=dev)
        if hasattr(engine, "publish"):
            engine.publish()                    # the tops were written on the ctx stream; torch's stream reads them next
        d_peaks = torch.cat([d_tops[:, -1], gathered[world - 1, k:k + n_tail_peaks]], dim=0) if k else gathered[world - 1, k:k + n_tail_peaks]
    else:
        d_roots = mine[:k].view(k, 1, 4)
        d_tops = torch.zeros((k, 0, 4), dtype=torch.int64, device=dev)
        d_peaks = mine[:k + n_tail_peaks]
    return ShardedMMR(n_total, world, rank, (ms, t), local, d_roots, d_tops, tail, d_peaks, engine)

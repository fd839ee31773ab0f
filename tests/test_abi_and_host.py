"""CPU-side tests: the C ABI library loads and exports every symbol include/pmt.h declares, the host-side index math
matches the reference's tables, and the sharded build's collective logic works across 2 gloo ranks (oracle engine)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, splitmix_felts


def test_library_builds_and_exports_every_declared_symbol():
    from plonky2_merkle_trees_b200 import _lib, build
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "pmt.h")).read()
    declared = set(re.findall(r"\b(pmt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.exported_symbols())
    assert lib.pmt_version().startswith(b"pmt")
    # pure index helpers need no GPU
    assert [lib.pmt_mmr_size(n) for n in (0, 1, 2, 3, 4, 7, 8)] == [0, 1, 3, 4, 7, 11, 15]


def _c_prototypes(header):
    """name -> list of 'p' (pointer) / 's' (scalar) per parameter, from the prototypes of include/pmt.h"""
    text = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    out = {}
    for name, args in re.findall(r"\b(pmt_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text):
        args = " ".join(args.split())
        kinds = [] if args in ("", "void") else ["p" if "*" in a else "s" for a in args.split(",")]
        out[name] = kinds
    return out


def test_header_ctypes_table_and_rust_extern_block_agree():
    """The ABI is declared three times -- include/pmt.h (the contract), _lib.py's ctypes signatures (what the tests and the
    bench call through) and rust/src/pmt_ffi.rs (the crate a maintainer links; it cannot be compiled in this image) --:
    every function they share must have the same number of parameters and the same pointer / scalar kind in each position."""
    import ctypes as C
    from plonky2_merkle_trees_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pmt.h")).read()
    protos = _c_prototypes(header)
    assert len(protos) >= 60
    # ctypes table: every declared function, exactly
    sigs = _lib.signatures()
    assert set(sigs) == set(protos)
    for name, (_, argtypes) in sigs.items():
        kinds = []
        for t in argtypes:
            is_ptr = t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or (isinstance(t, type) and issubclass(t, C._Pointer))
            kinds.append("p" if is_ptr else "s")
        # the table lists the ctx handle first for the functions that take one, like the header
        assert kinds == protos[name], (name, kinds, protos[name])
    # Rust extern block: a subset of the header (what the wrappers use), same shapes
    rs = open(os.path.join(ROOT, "rust", "src", "pmt_ffi.rs")).read()
    rs = re.sub(r"//[^\n]*", " ", rs)
    fns = re.findall(r"pub fn (pmt_[a-z0-9_]+)\s*\(([^;]*?)\)\s*(?:->\s*[^;]+)?;", rs, flags=re.S)
    assert len(fns) >= 35
    for name, args in fns:
        assert name in protos, "rust/src/pmt_ffi.rs declares %s, include/pmt.h does not" % name
        args = " ".join(args.split())
        kinds = ["p" if "*" in a else "s" for a in args.split(",") if a.strip()]
        assert kinds == protos[name], (name, kinds, protos[name])


def test_no_cpu_fallback_without_device():
    import torch
    from plonky2_merkle_trees_b200 import PmtError, _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(PmtError):
        _lib.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "plonky2_merkle_trees_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), os.path.join(dirpath, f)


def test_mmr_index_math_matches_reference_tables(golden, oracle):
    from plonky2_merkle_trees_b200 import mmr
    t = golden["reference"]["mmr_tables"]
    for size, bitmap in t["heights_bitmap"]:       # merkle_mountain_ranges.rs:280-297
        assert mmr.get_heights_bitmap_for_mmr_size(size) == (bitmap, 0)
    for normal, idx in t["mmr_index"]:             # :307-324
        assert mmr.get_mmr_index(normal) == idx
        assert mmr._normal_index(idx) == normal
    for size in range(0, 200):
        assert mmr.get_heights_bitmap_for_mmr_size(size) == oracle.mmr_heights_bitmap(size)
    with pytest.raises(ValueError):
        mmr._normal_index(2)                       # position 2 is an internal node
    with pytest.raises(OverflowError):
        mmr.get_mmr_index(1 << 30)


def test_sharding_index_math(oracle):
    from plonky2_merkle_trees_b200 import sharded
    import py_oracle as po
    assert sharded.shard_range(1 << 10, 8, 3) == (384, 128)
    for lg, h in [(3, 0), (4, 1), (5, 0), (6, 2)]:
        n = 1 << lg
        seen = set()
        for l in range(lg - h):
            for k in range(n >> l):
                seen.add(sharded.global_digest_index(n, h, l, k))
        assert seen == set(range(2 * (n - (1 << h))))
    assert [sharded.digest_index(l, k) for l, k in [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (0, 3), (2, 0), (2, 1)]] == list(range(8))
    assert sharded.digest_index(1, 3) == po.digest_index(1, 3)


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle")); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import oracle as orc
from conftest import splitmix_felts
from plonky2_merkle_trees_b200 import sharded

class OracleEngine:   # TEST DOUBLE: stands in for the GPU engine so the collective logic runs on CPU/gloo
    def build_local(self, leaves, h):
        dg, cap = orc.merkle_tree_new(leaves.numpy().view(np.uint64), h, threads=2, fast=True)
        return torch.from_numpy(dg.view(np.int64)), torch.from_numpy(cap.view(np.int64))
    def top_levels(self, roots, h):
        cur = roots.numpy().view(np.uint64); out = []
        while cur.shape[0] > (1 << h):
            cur = orc.two_to_one_batch(cur[0::2], cur[1::2]); out.append(cur)
        return torch.from_numpy(np.concatenate(out).view(np.int64))
    def sync(self): pass

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
for lg, w, h in [(6, 4, 0), (5, 9, 1), (5, 135, 3)]:
    n = 1 << lg
    rows = splitmix_felts(lg * 13 + w, n * w).reshape(n, w)
    s, c = sharded.shard_range(n, 2, rank)
    t = sharded.build_sharded_tree(torch.from_numpy(rows[s:s + c].view(np.int64)), n, h, OracleEngine())
    chunks = [torch.empty_like(t.local_digests) for _ in range(2)]
    dist.all_gather(chunks, t.local_digests)
    odg, ocap = orc.merkle_tree_new(rows, h)
    assert np.array_equal(t.cap.numpy().view(np.uint64), ocap), (lg, w, h)
    full = t.assemble_global([x.numpy().view(np.uint64) for x in chunks])
    assert np.array_equal(full, odg), (lg, w, h)
    assert t.local_offset() == (0 if rank == 0 else [i for i in range(full.shape[0]) if np.array_equal(full[i], chunks[1][0].numpy().view(np.uint64))][0])

# ---- sharded MMR: balanced rounds + tail, elements / peaks / bag / proofs against the sequential add_leaf oracle --------
class OracleMMR:          # TEST DOUBLE with the interface of plonky2_merkle_trees_b200.mmr.MMR
    def __init__(self): self.elements = np.zeros((0, 4), np.uint64); self.n_leaves = 0
    def extend_dev(self, t):
        x = t.numpy().view(np.uint64).reshape(-1)
        self.elements = orc.mmr_extend(self.elements if self.n_leaves else None, x); self.n_leaves += x.size
    def get_peaks(self): return orc.mmr_peaks(self.elements)
    def prove_batch(self, idx):
        sib = np.zeros((len(idx), 32, 4), np.uint64); left = np.zeros((len(idx), 32), np.uint8); ln = np.zeros(len(idx), np.uint32)
        for q, i in enumerate(idx):
            s_, l_ = orc.mmr_subtree_proof(self.elements, orc.mmr_index(int(i)))
            sib[q, :len(l_)] = s_; left[q, :len(l_)] = l_; ln[q] = len(l_)
        return sib, left, ln
OracleEngine.new_mmr = lambda self: OracleMMR()
OracleEngine.hash_or_noop = lambda self, felts: orc.hash_or_noop(np.asarray(felts, dtype=np.uint64))

for n in [2, 3, 8, 13, 31, 64, 77, 100, 255]:
    leaves = splitmix_felts(1000 + n, n)
    mine = np.concatenate([leaves[a:a + c] for a, c in sharded.mmr_shard_ranges(n, 2, rank)]) if sharded.mmr_shard_ranges(n, 2, rank) else np.zeros(0, np.uint64)
    sm = sharded.build_sharded_mmr(torch.from_numpy(mine.view(np.int64)), n, OracleEngine())
    want = orc.mmr_extend(None, leaves)
    assert len(sm) == want.shape[0], n
    assert np.array_equal(sm.get_peaks(), orc.mmr_peaks(want)), n
    assert np.array_equal(sm.bagging_the_peaks(), orc.mmr_bag(want)), n
    loc = torch.from_numpy(np.ascontiguousarray(sm.local.elements).view(np.int64))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(2)]
    dist.all_gather(sizes, torch.tensor([loc.shape[0]]))
    assert sizes[0].item() == sizes[1].item(), "rounds are balanced"
    chunks = [torch.empty_like(loc) for _ in range(2)]
    if loc.numel(): dist.all_gather(chunks, loc)
    tail_el = sm.tail.elements if sm.tail is not None else None
    if rank == 1:   # the last rank owns the tail and can assemble everything
        full = sm.assemble_global([c.numpy().view(np.uint64) for c in chunks], tail_el)
        assert np.array_equal(full, want), n
    root = orc.mmr_bag(want)
    for i in range(n):
        if sm.owner(i) != rank: continue
        pr = sm.get_proof_normal_index(i)
        es, el = orc.mmr_subtree_proof(want, orc.mmr_index(i))
        assert pr.mmr_size == want.shape[0] and len(pr.merkle_proof) == len(el), (n, i)
        for (d, on_left), ed, eo in zip(pr.merkle_proof, es, el):
            assert np.array_equal(d, ed) and bool(on_left) == bool(eo), (n, i)
        assert orc.mmr_verify(leaves[i], root, np.array([d for d, _ in pr.merkle_proof]).reshape(-1, 4), [int(b) for _, b in pr.merkle_proof], pr.peaks) == 1
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_build_two_gloo_ranks(oracle, tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


# ---- pmt_mmr_extend_multi: the cut (real index math of libpmt, no device) and the construction it relies on ------------------
def _mmr_size(n):
    return 2 * n - bin(n).count("1")


def _mmr_pos(l, k):
    last = ((k + 1) << l) - 1
    return 2 * last - bin(last).count("1") + l


@pytest.mark.parametrize("n0,m,g", [(0, 1 << 15, 2), (0, 1 << 16, 8), (12345, (1 << 15) + 77, 2), (4096, 3 * 4096, 3), (4095, 2 * 4096 + 2, 2),
                                    (1 << 14, (1 << 16) - 1, 4), (7, 40000, 4), (3 << 12, 1 << 13, 2), (100, 5000, 2)])
def test_mmr_multi_plan_and_block_construction(oracle, n0, m, g):
    """The plan pmt_mmr_extend_multi computes (libpmt's own pmt_mmr_multi_plan) and the construction behind it, replayed on the
    CPU with the oracle standing in for the contexts: head, aligned blocks as fresh MMRs written into their slices, the
    coarse MMR over the block roots, tail -- must equal the sequential MMR element by element."""
    from plonky2_merkle_trees_b200 import mmr
    leaves = splitmix_felts(4000 + n0 + m + g, n0 + m)
    want = oracle.mmr_extend(None, leaves)
    plan = mmr.multi_plan(n0, m, g)
    if m // g < 4096:
        assert plan is None
        return
    assert plan is not None
    b, a, z = plan
    big = 1 << b
    assert b >= 12 and a % big == 0 and z % big == 0 and n0 <= a < z <= n0 + m and a - n0 < big and n0 + m - z < big
    blocks = (z - a) >> b
    assert 1 <= blocks <= 8 * g
    out = np.zeros((_mmr_size(n0 + m), 4), np.uint64)
    out[:_mmr_size(n0)] = want[:_mmr_size(n0)]                     # the MMR before the append
    if a > n0:                                                     # head, by context 0
        out[:_mmr_size(a)] = oracle.mmr_extend(out[:_mmr_size(n0)], leaves[n0:a])
    for j in range(blocks):                                        # blocks, from empty, into their slices
        start = a + j * big
        out[_mmr_size(start):_mmr_size(start) + 2 * big - 1] = oracle.mmr_extend(None, leaves[start:start + big])
    s0, s1 = a >> b, (a >> b) + blocks                             # coarse MMR over the block roots
    sup = np.zeros((_mmr_size(s1), 4), np.uint64)
    base = 0
    for bit in range(63, -1, -1):
        if (s0 >> bit) & 1:
            base += 1 << bit
            sup[_mmr_size(base) - 1] = out[_mmr_size(base << b) - 1]
    for s_ in range(s0, s1):
        sup[_mmr_pos(0, s_)] = out[_mmr_pos(b, s_)]
    l = 1
    while (s1 >> l) > 0:
        for k in range(s0 >> l, s1 >> l):
            p = _mmr_pos(l, k)
            sup[p] = oracle.two_to_one(sup[p - (1 << l)], sup[p - 1])
            out[_mmr_pos(l + b, k)] = sup[p]
        l += 1
    if n0 + m > z:                                                 # tail
        out = oracle.mmr_extend(out[:_mmr_size(z)], leaves[z:])
    assert np.array_equal(out, want)


def test_mmr_shard_plan_in_the_library_equals_the_host_plan():
    """pmt_mmr_shard_plan (what pmt_mmr_build_sharded_dev cuts the leaves by) against sharded.mmr_shard_plan (what the gloo
    world-of-2 tests below replay with the oracle), 5 000 random (n, world); the rounds tile n exactly; bad worlds are refused."""
    import ctypes as C
    import random
    from plonky2_merkle_trees_b200 import _lib, sharded
    L = _lib.load()
    rnd = random.Random(11)
    k, ms, t = C.c_uint32(0), (C.c_size_t * 64)(), C.c_size_t(0)
    for _ in range(5000):
        w = 1 << rnd.randrange(0, 5)
        n = rnd.choice([rnd.randrange(0, 1 << 30), rnd.randrange(0, 4 * w), (1 << rnd.randrange(0, 31)) - rnd.randrange(0, 2)])
        assert L.pmt_mmr_shard_plan(n, w, C.byref(k), ms, C.byref(t)) == 0
        got = list(ms[:k.value])
        assert (got, t.value) == sharded.mmr_shard_plan(n, w)
        assert w * sum(got) + t.value == n and t.value < w and all(a > b for a, b in zip(got, got[1:]))
    for w in (0, 3, 12):
        assert L.pmt_mmr_shard_plan(100, w, C.byref(k), ms, C.byref(t)) == _lib.PMT_E_NOT_POW2


def test_mmr_multi_plan_invariants_random():
    """pmt_mmr_multi_plan over 20 000 random (n_before, m, contexts): block size, alignment, head / tail bounds, block count, and
    that every node the plan assigns (head, blocks, coarse nodes, tail) is a distinct new position of the post-order array."""
    import random
    from plonky2_merkle_trees_b200 import mmr
    rnd = random.Random(7)
    split = 0
    for it in range(20000):
        g = rnd.choice([1, 2, 3, 4, 5, 8, 16])
        m = rnd.choice([rnd.randrange(1, 1 << 14), rnd.randrange(1, 1 << 20), 1 << rnd.randrange(10, 24), (1 << rnd.randrange(12, 24)) - 1])
        n0 = rnd.choice([0, rnd.randrange(0, 1 << 22), 1 << rnd.randrange(0, 24), rnd.randrange(0, 1 << 12)])
        plan = mmr.multi_plan(n0, m, g)
        if g == 1 or m // g < 4096:
            assert plan is None
            continue
        if plan is None:        # no aligned block of 2^b leaves inside [n0, n0 + m): only possible for the smallest b
            fl = (m // g).bit_length() - 1
            b = max(fl - 2, 12)
            assert ((n0 + m) >> b) <= ((n0 + (1 << b) - 1) >> b)
            continue
        split += 1
        b, a, z = plan
        big = 1 << b
        assert b == max((m // g).bit_length() - 1 - 2, 12)
        assert a % big == 0 and z % big == 0 and n0 <= a < z <= n0 + m and a - n0 < big and n0 + m - z < big
        blocks = (z - a) >> b
        assert 1 <= blocks <= 8 * g
        if it % 50 == 0:        # the positions written by the four parts tile [mmr_size(n0), mmr_size(n0 + m)) exactly
            new = _mmr_size(n0 + m) - _mmr_size(n0)
            head = _mmr_size(a) - _mmr_size(n0)
            tail = _mmr_size(n0 + m) - _mmr_size(z)
            coarse = set()
            s0, s1 = a >> b, (a >> b) + blocks
            l = 1
            while (s1 >> l) > 0:
                for k in range(s0 >> l, s1 >> l):
                    p = _mmr_pos(l + b, k)
                    assert _mmr_size(a) <= p < _mmr_size(z) and p not in coarse
                    coarse.add(p)
                l += 1
            for j in range(blocks):
                lo = _mmr_size(a + j * big)
                assert not any(lo <= p < lo + 2 * big - 1 for p in coarse)
            assert head + blocks * (2 * big - 1) + len(coarse) + tail == new
    assert split > 3000

# ---- the wavefront launch's block arithmetic, as a model ------------------------------------------------------------------
def _wave_runs(l_first, l_last, n0, n1, coop_max, resident_nodes, wave_min=1 << 14):
    """restates launch_level_span (plonky2_merkle_trees_b200/csrc/pmt_api.cu): -> [(la, n_levels)] of the wavefront launches"""
    runs, l = [], l_first
    while l <= l_last:
        k0, k1 = n0 >> l, n1 >> l
        if k1 == 0:
            break
        if k1 <= k0:
            l += 1
            continue
        count = k1 - k0
        if count <= coop_max:
            break                                    # the cooperative tail takes over (never followed by a big level again)
        l2 = l
        while (count > resident_nodes and l2 < l_last and not ((n0 >> l2) & 1)
               and (n1 >> (l2 + 1)) - (n0 >> (l2 + 1)) > coop_max and (n1 >> (l2 + 1)) - (n0 >> (l2 + 1)) >= wave_min):
            l2 += 1
        if l2 > l:
            runs.append((l, l2 - l + 1))
        l = l2 + 1
    return runs


def _wave_block(logical_id, la, n_levels, n0, n1, B=128):
    """restates the head of k_levels_wave (merkle_kernels.cuh): logical block number -> (level index j, block of the level,
    first node, node count of the level, flags offset of the level, blocks of the level below)"""
    idx, off, prev, j = logical_id, 0, 0, 0
    while True:
        k0 = n0 >> (la + j)
        cnt = (n1 >> (la + j)) - k0
        nb = (cnt + B - 1) // B
        if idx < nb or j + 1 == n_levels:
            return j, idx, k0, cnt, off, prev
        idx -= nb
        off += nb
        prev = nb
        j += 1


def test_wavefront_block_arithmetic_model():
    """k_levels_wave: every node of every level of a run is computed by exactly one block; the children of a block's nodes lie
    in the two blocks of the level below whose flags it waits for; those blocks have SMALLER logical numbers (they started
    earlier: no deadlock) and distinct flag slots; a run only crosses levels that start at an even node.  Random perfect and
    ragged ranges, incl. appends to a non-empty MMR (n0 > 0)."""
    import random
    rnd = random.Random(5)
    B, coop_max, resident = 128, 1 << 13, 148 * 5 * 128
    checked = 0
    for it in range(400):
        kind = it % 4
        if kind == 0:                                   # a perfect tree
            lg = rnd.randrange(15, 23)
            n0, n1 = 0, 1 << lg
        elif kind == 1:                                 # ragged from empty
            n0, n1 = 0, rnd.randrange(1 << 17, 1 << 22)
        elif kind == 2:                                 # append to a non-empty MMR, even start
            n0 = 2 * rnd.randrange(1, 1 << 20)
            n1 = n0 + rnd.randrange(1 << 17, 1 << 21)
        else:                                           # an aligned pipeline chunk
            b = rnd.randrange(18, 22)
            n0 = rnd.randrange(0, 16) << b
            n1 = n0 + (1 << b)
        for la, n_levels in _wave_runs(1, 40, n0, n1, coop_max, resident):
            checked += 1
            for l in range(la, la + n_levels - 1):
                assert not ((n0 >> l) & 1)
            blocks = [((n1 >> (la + j)) - (n0 >> (la + j)) + B - 1) // B for j in range(n_levels)]
            total = sum(blocks)
            # sample logical ids: all block boundaries of every level and a few random ones
            ids = set()
            acc = 0
            for nb in blocks:
                ids.update({acc, acc + nb - 1, acc + nb // 2})
                acc += nb
            ids.update(rnd.randrange(total) for _ in range(20))
            seen_flags = {}
            for lid in sorted(ids):
                j, idx, k0, cnt, off, prev = _wave_block(lid, la, n_levels, n0, n1)
                assert 0 <= j < n_levels and idx < blocks[j] and off == sum(blocks[:j]) and prev == (blocks[j - 1] if j else 0)
                assert k0 == n0 >> (la + j) and cnt == (n1 >> (la + j)) - k0
                lo, hi = idx * B, min(idx * B + B, cnt)             # this block's nodes, local to the level
                assert lo < hi
                assert seen_flags.setdefault(off + idx, lid) == lid  # one flag slot per block
                if j:
                    # children of local node i are the local nodes 2 i, 2 i + 1 of the level below (it starts at 2 k0)
                    assert (n0 >> (la + j - 1)) == 2 * k0
                    cnt_below = (n1 >> (la + j - 1)) - (n0 >> (la + j - 1))
                    assert 2 * (hi - 1) + 1 < cnt_below
                    child_blocks = {(2 * lo) // B, (2 * (hi - 1) + 1) // B}
                    waited = {c for c in (2 * idx, 2 * idx + 1) if c < prev}
                    assert child_blocks <= waited
                    for c in waited:                                 # they were numbered (= started) before this block
                        assert (off - prev) + c < lid
            # the blocks of a level tile its nodes exactly
            for j, nb in enumerate(blocks):
                cnt = (n1 >> (la + j)) - (n0 >> (la + j))
                assert (nb - 1) * B < cnt <= nb * B
    assert checked > 300

# ---- the mailbox exchange's ring / epoch protocol, as a model ---------------------------------------------------------------
def _mailbox_run(world, exchanges, ring, rnd):
    """Random interleaving of k_exchange_top's steps on `world` ranks (merkle_kernels.cuh): per exchange s a rank pushes
    (data, then flag = s + 1) into slot s % ring of every peer's mailbox -- peer by peer, in any order relative to the other
    ranks --, then waits until all peers' flags of s are in its own mailbox, then reads the peers' data.  Returns the number of
    reads that saw another exchange's data (0 = the protocol held)."""
    data = [[[None] * world for _ in range(ring)] for _ in range(world)]      # data[owner][slot][from]
    flag = [[[0] * world for _ in range(ring)] for _ in range(world)]
    # per rank: (exchange, phase, progress); phases: 0 push to peers one at a time, 1 wait, 2 read
    state = [[0, 0, 0] for _ in range(world)]
    order = [rnd.sample([p for p in range(world) if p != r], world - 1) for r in range(world)]
    bad = 0
    live = set(range(world))
    while live:
        r = rnd.choice(sorted(live))
        s, phase, k = state[r]
        slot = s % ring
        if phase == 0:
            p = order[r][k]
            data[p][slot][r] = (r, s)
            flag[p][slot][r] = s + 1
            state[r][2] += 1
            if state[r][2] == world - 1:
                state[r][1:] = [1, 0]
        elif phase == 1:
            if any(flag[r][slot][p] > s + 1 for p in range(world) if p != r):
                bad += 1                                 # a later exchange overwrote the flag: the real kernel would time out
                state[r][1] = 2
            elif all(flag[r][slot][p] == s + 1 for p in range(world) if p != r):
                state[r][1] = 2
        else:
            bad += sum(data[r][slot][p] != (p, s) for p in range(world) if p != r)
            if s + 1 == exchanges:
                live.discard(r)
            else:
                state[r] = [s + 1, 0, 0]
                order[r] = rnd.sample([p for p in range(world) if p != r], world - 1)
    return bad


def test_mailbox_ring_protocol_model():
    """A rank cannot get more than one exchange ahead of its slowest peer (it needs all flags of exchange s before it leaves s),
    so slots are never overwritten while a reader still needs them -- with the library's ring of 4 and already with 2; a ring of
    ONE slot must fail under some interleaving, which shows that the model can see the hazard."""
    import random
    rnd = random.Random(17)
    for world in (2, 4, 8):
        for ring in (4, 2):
            for _ in range(30):
                assert _mailbox_run(world, 12, ring, rnd) == 0
    assert any(_mailbox_run(4, 12, 1, rnd) > 0 for _ in range(200))

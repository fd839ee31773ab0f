"""GPU parity: the CUDA path, called through the C ABI (libpmt.so), against the CPU oracle -- bit-exact.

Run on the B200 box: python -m pytest tests -m gpu.  Nothing here reads /root/reference.
"""
import ctypes as C
import random

import numpy as np
import pytest

from conftest import EDGE, P, splitmix_felts, u64

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from plonky2_merkle_trees_b200 import _lib
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _lib.default_context(0)


@pytest.fixture(scope="module")
def api(ctx):
    from plonky2_merkle_trees_b200 import hasher, merkle_tree, mmr, simple_merkle_tree
    class NS: pass
    ns = NS()
    ns.hasher, ns.mt, ns.mmr, ns.smt = hasher, merkle_tree, mmr, simple_merkle_tree
    return ns


def edge_states(count, seed):
    rnd = random.Random(seed)
    out = np.zeros((count, 12), np.uint64)
    for i in range(count):
        for j in range(12):
            out[i, j] = rnd.choice(EDGE) if rnd.random() < 0.35 else rnd.getrandbits(64)
    return out


# ---- permutation / Hasher ------------------------------------------------------------------------------------------
def test_permute_upstream_vectors_and_random(api, oracle, golden):
    st = np.concatenate([np.zeros((1, 12), np.uint64), np.arange(12, dtype=np.uint64).reshape(1, 12), edge_states(3000, 1),
                         np.full((1, 12), 2**64 - 1, np.uint64), np.full((1, 12), P - 1, np.uint64)])
    got = api.hasher.permute(st)
    assert got[0, :4].tolist() == golden["upstream"]["perm_zero"]
    assert got[1, :4].tolist() == golden["upstream"]["perm_range12"]
    for i in range(st.shape[0]):
        assert got[i].tolist() == oracle.permute(st[i]).tolist(), i
    assert (got < np.uint64(P)).all()


def test_two_to_one_reference_vectors(api, golden):
    # every two_to_one pair pinned by simple_merkle_tree.rs:136-140, :181-190
    for name in ("test_build_merkle_tree_4_leaves", "test_build_merkle_tree_16_leaves"):
        g = golden["reference"]["simple_tree"][name]
        levels = [u64(l) for l in g["levels"]] + [u64([g["root"]])]
        for a, b in zip(levels[:-1], levels[1:]):
            assert api.hasher.two_to_one(a[0::2], a[1::2]).tolist() == b.tolist()


def test_two_to_one_random_and_noncanonical(api, oracle):
    st = edge_states(2000, 2)
    l, r = st[:, :4].copy(), st[:, 4:8].copy()
    got = api.hasher.two_to_one(l, r)
    assert got.tolist() == oracle.two_to_one_batch(l, r).tolist()


@pytest.mark.parametrize("w", [1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 17, 24, 25, 135])
def test_hash_or_noop_widths(api, oracle, golden, w):
    rows = splitmix_felts(w, 257 * w).reshape(257, w)
    rows[0] = np.arange(w, dtype=np.uint64)
    rows[1, :] = np.uint64(2**64 - 1)
    rows[2, :] = np.uint64(P)
    got = api.hasher.hash_or_noop(rows)
    assert got.tolist() == oracle.hash_or_noop_rows(rows).tolist()
    key = "hash_no_pad_range_%d" % w
    if key in golden["derived_unpinned"]:
        assert got[0].tolist() == golden["derived_unpinned"][key]
    if w <= 4:  # the no-op rule (simple_merkle_tree.rs:210): never permuted, only canonicalised
        assert got[0, :w].tolist() == list(range(w)) and got[1, :w].tolist() == [2**32 - 2] * w and got[2].tolist() == [0] * 4
        assert api.hasher.hash_no_pad(rows)[0].tolist() == oracle.hash_no_pad(rows[0]).tolist()


def test_empty_batches(api):
    assert api.hasher.two_to_one(np.zeros((0, 4), np.uint64), np.zeros((0, 4), np.uint64)).shape == (0, 4)
    assert api.hasher.permute(np.zeros((0, 12), np.uint64)).shape == (0, 12)


# ---- simple tree ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["test_build_merkle_tree_4_leaves", "test_build_merkle_tree_16_leaves"])
def test_simple_tree_reference_vectors(api, golden, name):
    g = golden["reference"]["simple_tree"][name]
    t = api.smt.MerkleTree.build(g["leaves"])
    assert t.count_levels == len(g["levels"])
    for got, want in zip(t.tree, g["levels"]):
        assert got.tolist() == want
    assert t.root.tolist() == g["root"]


def test_simple_tree_asserted_proof(api, golden):
    g = golden["reference"]["simple_tree"]["test_merkle_proof_small_tree"]
    t = api.smt.MerkleTree.build(g["leaves"])
    proof = t.get_merkle_proof(0)
    assert proof.tolist() == g["proof"]                      # simple_merkle_tree.rs:210-211
    assert api.smt.verify_merkle_proof(g["leaves"][0], 0, t.root, proof)


def test_simple_tree_verify_roundtrip_and_negatives(api, oracle, golden):
    leaves = golden["reference"]["simple_tree"]["test_build_merkle_tree_16_leaves"]["leaves"]
    t = api.smt.MerkleTree.build(leaves)
    proofs = t.get_merkle_proofs(list(range(16)))
    assert api.smt.verify_merkle_proofs(leaves, list(range(16)), t.root, proofs).all()
    for i in range(16):
        assert oracle.simple_tree_verify(leaves[i], i, t.root, proofs[i])
        assert t.get_in_between_hashes(i).tolist() == oracle.simple_tree_in_between(np.concatenate(t.tree), t.root, 16, i).tolist()
    p3 = proofs[3]
    assert not api.smt.verify_merkle_proof(leaves[4], 3, t.root, p3)          # wrong leaf  (:298)
    assert not api.smt.verify_merkle_proof(leaves[3], 2, t.root, p3)          # wrong index (:300)
    bad = p3.copy(); bad[1, 0] ^= np.uint64(1)
    assert not api.smt.verify_merkle_proof(leaves[3], 3, t.root, bad)         # wrong proof (:303)
    assert not api.smt.verify_merkle_proof(leaves[3], 3, t.tree[0][0], p3)    # wrong root  (:306)
    with pytest.raises(IndexError):
        t.get_merkle_proof(16)
    from plonky2_merkle_trees_b200 import PmtError
    with pytest.raises(PmtError):
        api.smt.MerkleTree.build([1, 2, 3])
    with pytest.raises(PmtError):
        api.smt.MerkleTree.build([1])


@pytest.mark.parametrize("lg", [1, 2, 5, 8, 9, 10, 13])
def test_simple_tree_vs_oracle(api, oracle, lg):
    n = 1 << lg
    leaves = splitmix_felts(100 + lg, n)
    leaves[:min(n, 8)] = u64(EDGE)[:min(n, 8)]
    t = api.smt.MerkleTree.build(leaves)
    levels, root = oracle.simple_tree_build(leaves)
    assert np.array_equal(np.concatenate(t.tree), levels)
    assert np.array_equal(t.root, root)
    idx = sorted({0, 1, n // 2, n - 1})
    proofs = t.get_merkle_proofs(idx)
    for i, pr in zip(idx, proofs):
        assert np.array_equal(pr, oracle.simple_tree_proof(levels, n, i))


def test_simple_tree_host_buffer_abi(ctx, oracle):
    from plonky2_merkle_trees_b200._lib import ptr
    n = 1 << 10  # BASELINE config C1
    leaves = splitmix_felts(0x706d745f62323030, n)
    levels = np.zeros((2 * n - 2, 4), np.uint64); root = np.zeros(4, np.uint64)
    ctx.call("pmt_simple_tree_build", ptr(leaves), n, ptr(levels), ptr(root))
    ol, orr = oracle.simple_tree_build(leaves)
    assert np.array_equal(levels, ol) and np.array_equal(root, orr)


# ---- plonky2 MerkleTree::new ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("lg,w,h", [(0, 4, 0), (1, 1, 0), (1, 4, 1), (3, 4, 0), (4, 5, 2), (5, 9, 5), (6, 135, 4), (6, 3, 0),
                                    (10, 4, 0), (11, 7, 3), (9, 135, 4), (12, 1, 0), (12, 4, 12)])
def test_merkle_tree_new_vs_oracle(api, oracle, lg, w, h):
    n = 1 << lg
    rows = splitmix_felts(7 * lg + w, n * w).reshape(n, w)
    rows[0, :] = np.uint64(2**64 - 1)
    t = api.mt.MerkleTree.new(rows, h)
    dg, cap = oracle.merkle_tree_new(rows, h, threads=4, fast=True)
    assert np.array_equal(t.digests, dg)
    assert np.array_equal(t.cap, cap)
    idx = sorted({0, n // 3, n - 1})
    proofs = t.prove_batch(idx)
    for i, pr in zip(idx, proofs):
        assert np.array_equal(pr, oracle.merkle_prove(dg, n, h, i))
    ok = api.mt.verify_merkle_proofs_to_cap(rows[idx], idx, cap, h, proofs)
    assert ok.all()
    if proofs.shape[1]:
        bad = proofs.copy(); bad[0, 0, 0] ^= np.uint64(1)
        assert not api.mt.verify_merkle_proofs_to_cap(rows[idx], idx, cap, h, bad)[0]
    wrong_row = rows[idx].copy(); wrong_row[0, 0] ^= np.uint64(1)
    assert not api.mt.verify_merkle_proofs_to_cap(wrong_row, idx, cap, h, proofs)[0]


def test_merkle_tree_errors(api):
    from plonky2_merkle_trees_b200 import PmtError
    with pytest.raises(PmtError):
        api.mt.MerkleTree.new(np.zeros((3, 4), np.uint64), 0)
    with pytest.raises(PmtError):
        api.mt.MerkleTree.new(np.zeros((4, 4), np.uint64), 3)


def test_merkle_tree_host_buffer_abi(ctx, oracle):
    from plonky2_merkle_trees_b200._lib import ptr
    n, w, h = 1 << 9, 135, 4
    rows = splitmix_felts(3, n * w).reshape(n, w)
    dg = np.zeros((2 * (n - 16), 4), np.uint64); cap = np.zeros((16, 4), np.uint64)
    ctx.call("pmt_merkle_tree_build", ptr(rows), n, w, h, ptr(dg), ptr(cap))
    odg, ocap = oracle.merkle_tree_new(rows, h, threads=4, fast=True)
    assert np.array_equal(dg, odg) and np.array_equal(cap, ocap)


@pytest.mark.parametrize("lg,w,h", [(18, 4, 0), (18, 4, 3), (16, 135, 4), (19, 5, 17)])
def test_merkle_tree_host_buffer_pipelined_path(ctx, api, oracle, lg, w, h):
    """large host-buffer builds take the chunked H2D / hash / D2H pipeline: must equal the one-shot device build + oracle"""
    from plonky2_merkle_trees_b200._lib import ptr
    n = 1 << lg
    rows = splitmix_felts(lg + w + h, n * w).reshape(n, w)
    ncap = 1 << h
    dg = np.zeros((2 * (n - ncap), 4), np.uint64); cap = np.zeros((ncap, 4), np.uint64)
    ctx.call("pmt_merkle_tree_build", ptr(rows), n, w, h, ptr(dg), ptr(cap))
    t = api.mt.MerkleTree.new(rows, h)
    assert np.array_equal(dg, t.digests) and np.array_equal(cap, t.cap)
    if lg <= 18:
        odg, ocap = oracle.merkle_tree_new(rows, h, threads=oracle.max_threads(), fast=True)
        assert np.array_equal(dg, odg) and np.array_equal(cap, ocap)
    # a second call reuses the arenas and streams
    dg2 = np.zeros_like(dg); cap2 = np.zeros_like(cap)
    ctx.call("pmt_merkle_tree_build", ptr(rows), n, w, h, ptr(dg2), ptr(cap2))
    assert np.array_equal(dg, dg2) and np.array_equal(cap, cap2)


def test_full_size_properties_2p20(api, oracle):
    """BASELINE C2 (2^20 x 4, cap 0) at full size: structure checks that do not need a full CPU rebuild."""
    n, w = 1 << 20, 4
    rows = splitmix_felts(0x706d745f62323030, n * w).reshape(n, w)
    t0 = api.mt.MerkleTree.new(rows, 0)
    t4 = api.mt.MerkleTree.new(rows, 4)
    # (1) the 16 cap entries of the cap-4 tree hash up to the cap-0 root
    lvl = t4.cap
    while lvl.shape[0] > 1:
        lvl = api.hasher.two_to_one(lvl[0::2], lvl[1::2])
    assert np.array_equal(lvl[0], t0.cap[0])
    # (2) leaf digests are the canonical no-op copy, at digest positions 4j, 4j+1
    d = t0.digests
    assert np.array_equal(d[0::4][:n // 2], rows[0::2]) and np.array_equal(d[1::4][:n // 2], rows[1::2])
    # (3) one random cap-4 subtree (2^16 leaves) rebuilt by the multi-threaded oracle is byte-identical
    c = 11
    sub = rows[c << 16:(c + 1) << 16]
    odg, ocap = oracle.merkle_tree_new(sub, 0, threads=oracle.max_threads(), fast=True)
    per = t4.digests.shape[0] // 16
    assert np.array_equal(t4.digests[c * per:(c + 1) * per], odg) and np.array_equal(t4.cap[c], ocap[0])
    # (4) 1024 proofs verify against the cap; a corrupted one does not
    idx = (splitmix_felts(5, 1024) % np.uint64(n)).astype(np.uint64)
    proofs = t0.prove_batch(idx)
    assert api.mt.verify_merkle_proofs_to_cap(rows[idx.astype(np.int64)], idx, t0.cap, 0, proofs).all()


def test_full_size_properties_2p24(api, oracle):
    """The headline workload (2^24 x 4 felts, cap 0) and BASELINE C3 (MMR of 2^24 and 2^24 - 1 leaves) at FULL size, through
    properties that do not need a CPU rebuild of 16.7 M permutations:
    (1) every one of 6 000 randomly sampled inner nodes, spread over all 24 levels, equals the ORACLE's two_to_one of its two
        children as stored in the digests array (so a wrong node anywhere on a sampled path cannot hide);
    (2) the root is the fold of the 16 cap-4 subtree roots of a second, independent build (checksum of checksums);
    (3) 1 024 proofs verify against the cap and fail for the neighbouring leaf;
    (4) the MMR over the same first-column leaves is ONE mountain whose nodes are the width-1 tree's, its 2^24 - 1 sibling
        has 24 peaks whose bag equals the oracle's sponge over the GPU's peaks, and 1 024 proofs verify."""
    lg, w = 24, 4
    n = 1 << lg
    rows = splitmix_felts(0x706d745f62323030, n * w).reshape(n, w)
    t0 = api.mt.MerkleTree.new(rows, 0)
    d = t0.digests
    assert d.shape == (2 * n - 2, 4)
    node = lambda l, k: 2 * (((k >> 1) << (l + 1)) + (1 << l) - 1) + (k & 1)      # upstream's interleaved layout, cap 0
    rnd = random.Random(24)
    ls, ks = [], []
    for _ in range(6000):
        l = rnd.randrange(1, lg)
        ls.append(l); ks.append(rnd.randrange(n >> l))
    par = np.array([node(l, k) for l, k in zip(ls, ks)])
    c0 = np.array([node(l - 1, 2 * k) for l, k in zip(ls, ks)])
    c1 = np.array([node(l - 1, 2 * k + 1) for l, k in zip(ls, ks)])
    assert np.array_equal(oracle.two_to_one_batch(d[c0], d[c1]), d[par])
    assert np.array_equal(oracle.two_to_one(d[node(lg - 1, 0)], d[node(lg - 1, 1)]), t0.cap[0])
    assert np.array_equal(d[0::4][:n // 2], rows[0::2]) and np.array_equal(d[1::4][:n // 2], rows[1::2])   # no-op leaves
    assert (d < np.uint64(P)).all()                                                                          # canonical
    # (2)
    t4 = api.mt.MerkleTree.new(rows, 4)
    lvl = t4.cap
    while lvl.shape[0] > 1:
        lvl = oracle.two_to_one_batch(lvl[0::2], lvl[1::2])
    assert np.array_equal(lvl[0], t0.cap[0])
    # (3)
    idx = (splitmix_felts(5, 1024) % np.uint64(n)).astype(np.uint64)
    proofs = t0.prove_batch(idx)
    assert proofs.shape == (1024, lg, 4)
    assert api.mt.verify_merkle_proofs_to_cap(rows[idx.astype(np.int64)], idx, t0.cap, 0, proofs).all()
    assert not api.mt.verify_merkle_proofs_to_cap(rows[(idx ^ np.uint64(1)).astype(np.int64)], idx, t0.cap, 0, proofs).any()
    del t0, t4, d
    # (4)
    leaves = np.ascontiguousarray(rows[:, 0])
    m = api.mmr.MMR.new(); m.extend(leaves)
    t1 = api.mt.MerkleTree.new(leaves.reshape(n, 1), 0)
    assert len(m) == 2 * n - 1 and m.get_peaks().shape[0] == 1
    assert np.array_equal(m.bagging_the_peaks(), t1.cap[0])
    el, d1 = m.elements, t1.digests
    pos = lambda h, k: 2 * (((k + 1) << h) - 1) - bin(((k + 1) << h) - 1).count("1") + h
    for _ in range(2000):
        h = rnd.randrange(0, lg)
        k = rnd.randrange(n >> h)
        assert np.array_equal(el[pos(h, k)], d1[node(h, k)])
    del el, d1, t1
    m2 = api.mmr.MMR.new(); m2.extend(leaves[:n - 1])
    peaks = m2.get_peaks()
    assert peaks.shape[0] == 24 and len(m2) == 2 * (n - 1) - 24
    assert np.array_equal(m2.bagging_the_peaks(), oracle.hash_or_noop(peaks.reshape(-1)))
    idx = (splitmix_felts(6, 1024) % np.uint64(n - 1)).astype(np.uint64)
    sib, left, ln = m2.prove_batch(idx)
    st = api.mmr.verify_batch(leaves[idx.astype(np.int64)], sib, left, ln, peaks, m2.bagging_the_peaks())
    assert (st == 1).all()
    # one proof end to end against the oracle's verifier
    i = int(idx[0])
    assert oracle.mmr_verify(int(leaves[i]), m2.bagging_the_peaks(), sib[0, :ln[0]], left[0, :ln[0]], peaks) == 1


# ---- MMR -------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", list(range(1, 36)) + [63, 64, 65, 70, 100, 255, 1000, 4097])
def test_mmr_extend_vs_sequential_add_leaf(api, oracle, n):
    leaves = splitmix_felts(1000 + n, n)
    m = api.mmr.MMR.new()
    m.extend(leaves)
    want = oracle.mmr_extend(None, leaves)
    assert len(m) == want.shape[0] == 2 * n - bin(n).count("1")
    assert np.array_equal(m.elements, want)
    assert np.array_equal(m.get_peaks(), oracle.mmr_peaks(want))
    assert np.array_equal(m.bagging_the_peaks(), oracle.mmr_bag(want))


def test_mmr_incremental_and_add_leaf(api, oracle):
    leaves = splitmix_felts(77, 300)
    m = api.mmr.MMR.new()
    cuts = [0, 1, 2, 5, 6, 64, 65, 130, 299, 300]
    for a, b in zip(cuts[:-1], cuts[1:]):
        m.extend(leaves[a:b])
        assert np.array_equal(m.elements, oracle.mmr_extend(None, leaves[:b]))
    m2 = api.mmr.MMR.new()
    for x in leaves[:20]:
        m2.add_leaf(int(x))
    assert np.array_equal(m2.elements, oracle.mmr_extend(None, leaves[:20]))


@pytest.mark.parametrize("n0,m", [(0, 250001), (0, (1 << 18) - 1), (1 << 18, 200001), ((1 << 17) + 6, 210000), (3 << 12, (1 << 18) + (1 << 12))])
def test_big_ragged_mmr_appends_through_the_wavefront_kernel(api, oracle, n0, m):
    """k_levels_wave on ranges that are not perfect trees: appends big enough for the wavefront launch (level 1 more than fills
    the GPU) whose levels have ragged ends (last blocks partial, a level-(l+1) block with one child block) and, for n0 > 0, start
    at nodes that are even only up to some level (the wave stops where bit l of n0 is set and the remaining levels run one
    launch each).  Device-resident and host-buffer (compact MmrAppend) forms against the sequential add_leaf oracle; both wave
    settings must agree (PMT_WAVE is read when the context is created)."""
    import os
    from plonky2_merkle_trees_b200 import _lib
    leaves = splitmix_felts(5000 + (n0 % 1000) + m % 1000, n0 + m)
    want = oracle.mmr_extend(None, leaves)
    got = {}
    for wave in ("1", "0"):
        os.environ["PMT_WAVE"] = wave
        c = _lib.Context(0)
        try:
            mm = api.mmr.MMR.new(c)
            if n0:
                mm.extend(leaves[:n0])
            mm.extend(leaves[n0:])
            got[wave] = mm.elements.copy()
            if wave == "1":                       # the host-buffer append: only old peaks up, only new elements down
                import ctypes as C
                from plonky2_merkle_trees_b200._lib import u64p
                el = np.zeros((want.shape[0], 4), np.uint64)
                size0 = 2 * n0 - bin(n0).count("1")
                el[:size0] = want[:size0]
                new = np.ascontiguousarray(leaves[n0:])
                c.call("pmt_mmr_extend", el.ctypes.data_as(u64p), n0, new.ctypes.data_as(u64p), m)
                assert np.array_equal(el, want)
        finally:
            c.close()
            os.environ.pop("PMT_WAVE", None)
    assert np.array_equal(got["1"], want) and np.array_equal(got["0"], want)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 13, 16, 22, 31, 70, 1031])
def test_mmr_proofs_and_verify(api, oracle, n):
    leaves = splitmix_felts(2000 + n, n)
    m = api.mmr.MMR.new()
    m.extend(leaves)
    el = oracle.mmr_extend(None, leaves)
    root = m.bagging_the_peaks()
    peaks = m.get_peaks()
    idx = list(range(n)) if n <= 70 else [0, 1, 511, 512, 1023, 1024, 1027, 1028, 1029, 1030]
    sib, left, ln = m.prove_batch(idx)
    for q, i in enumerate(idx):
        osib, oleft = oracle.mmr_subtree_proof(el, api.mmr.get_mmr_index(i))
        assert int(ln[q]) == osib.shape[0]
        assert np.array_equal(sib[q, :ln[q]], osib) and np.array_equal(left[q, :ln[q]], oleft)
    st = api.mmr.verify_batch(leaves[idx], sib, left, ln, peaks, root)
    assert (st == 1).all()
    assert (api.mmr.verify_batch(leaves[idx] ^ np.uint64(1), sib, left, ln, peaks, root) == -1).all()   # assert! :245
    assert (api.mmr.verify_batch(leaves[idx], sib, left, ln, peaks, root ^ np.uint64(1)) == 0).all()
    # the reference-shaped single-proof API
    pr = m.get_proof(api.mmr.get_mmr_index(idx[-1]))
    assert pr.mmr_size == el.shape[0] and pr.verify(int(leaves[idx[-1]]), root)
    with pytest.raises(AssertionError):
        pr.verify(int(leaves[idx[-1]]) ^ 1, root)
    assert oracle.mmr_verify(leaves[idx[-1]], root, np.array([d for d, _ in pr.merkle_proof]).reshape(-1, 4),
                             [int(b) for _, b in pr.merkle_proof], pr.peaks) == 1


def test_mmr_host_buffer_abi(ctx, oracle):
    from plonky2_merkle_trees_b200._lib import ptr
    lib = ctx.lib
    n0, m = 37, 100
    leaves = splitmix_felts(9, n0 + m)
    el = np.zeros((lib.pmt_mmr_size(n0 + m), 4), np.uint64)
    ctx.call("pmt_mmr_extend", ptr(el), 0, ptr(leaves[:n0]), n0)
    ctx.call("pmt_mmr_extend", ptr(el), n0, ptr(leaves[n0:]), m)
    want = oracle.mmr_extend(None, leaves)
    assert np.array_equal(el, want)
    root = np.zeros(4, np.uint64); peaks = np.zeros((64, 4), np.uint64); k = C.c_uint32(0)
    ctx.call("pmt_mmr_bag", ptr(el), n0 + m, ptr(root))
    ctx.call("pmt_mmr_peaks", ptr(el), n0 + m, ptr(peaks), C.byref(k))
    assert np.array_equal(root, oracle.mmr_bag(want)) and np.array_equal(peaks[:k.value], oracle.mmr_peaks(want))
    assert lib.pmt_mmr_index(15) == 26


@pytest.mark.parametrize("n0,m", [(0, (1 << 20) + 777), (12345, 1 << 20), (3 * (1 << 16), (1 << 21) - 5)])
def test_mmr_host_buffer_pipeline(ctx, api, oracle, n0, m):
    """pmt_mmr_extend for big batches: old peaks only are uploaded, and the append runs as a pipeline of aligned chunks,
    each downloading one contiguous slice of `elements`.  Against the device-resident path (tied to the oracle above) for
    every element, and against the sequential add_leaf oracle for the first 2^17 + n0 leaves."""
    from plonky2_merkle_trees_b200._lib import ptr
    from plonky2_merkle_trees_b200.device import to_device
    lib = ctx.lib
    leaves = splitmix_felts(n0 + m, n0 + m)
    el = np.zeros((lib.pmt_mmr_size(n0 + m), 4), np.uint64)
    if n0:
        ctx.call("pmt_mmr_extend", ptr(el), 0, ptr(leaves[:n0]), n0)
    before = el[:lib.pmt_mmr_size(n0)].copy()
    ctx.call("pmt_mmr_extend", ptr(el), n0, ptr(leaves[n0:]), m)
    assert np.array_equal(el[:before.shape[0]], before)            # the old elements are never written
    ref = api.mmr.MMR.new(ctx)
    ref.extend_dev(to_device(leaves, "cuda:0"))
    assert np.array_equal(el, ref.elements)
    k = n0 + (1 << 17)
    want = oracle.mmr_extend(None, leaves[:k])
    # elements of the first k leaves that are not touched by later leaves = everything below the last peak merge: compare
    # the prefix that the k-leaf MMR and the full MMR share (all of it: post-order is append-only)
    assert np.array_equal(el[:want.shape[0]], want)


def test_mmr_2p20_against_tree(api):
    """an MMR with 2^20 leaves is one mountain: its nodes are the simple tree's levels in post-order (size check)."""
    n = 1 << 20
    leaves = splitmix_felts(0x706d745f62323030, n)
    m = api.mmr.MMR.new(); m.extend(leaves)
    t = api.smt.MerkleTree.build(leaves)
    assert np.array_equal(m.bagging_the_peaks(), t.root)      # 1 peak => bag = identity
    el = m.elements
    assert np.array_equal(el[-1], t.root)
    pos = lambda h, k: 2 * (((k + 1) << h) - 1) - bin(((k + 1) << h) - 1).count("1") + h
    for h in (0, 1, 7, 19):
        ks = [0, 1, (n >> h) - 1]
        for k in ks:
            assert np.array_equal(el[pos(h, k)], t.tree[h][k])
    # ragged: 2^20 - 1 leaves => 20 peaks, multi-block sponge bag; proofs for random leaves verify
    m2 = api.mmr.MMR.new(); m2.extend(leaves[:n - 1])
    assert m2.get_peaks().shape[0] == 20
    idx = (splitmix_felts(6, 64) % np.uint64(n - 1)).astype(np.uint64)
    sib, left, ln = m2.prove_batch(idx)
    st = api.mmr.verify_batch(leaves[idx.astype(np.int64)], sib, left, ln, m2.get_peaks(), m2.bagging_the_peaks())
    assert (st == 1).all()


# ---- host-buffer proofs and verification (pmt_*_prove / pmt_*_verify on the caller's arrays) ------------------------
def test_host_buffer_prove_and_verify_abi(ctx, oracle):
    """The whole proof path with HOST buffers only, as the Rust shim would drive it: build -> prove (host gathers) -> verify
    (upload, fold on the GPU, download), each step against the oracle, incl. the reference's panics as PMT_E_RANGE."""
    import ctypes as C
    from plonky2_merkle_trees_b200._lib import PMT_E_RANGE, PmtError, ptr, u8p, u32p, i8p
    p8 = lambda a: a.ctypes.data_as(u8p)
    # simple tree
    n, lg = 1 << 9, 9
    leaves = splitmix_felts(3, n)
    levels = np.zeros((2 * n - 2, 4), np.uint64); root = np.zeros(4, np.uint64)
    ctx.call("pmt_simple_tree_build", ptr(leaves), n, ptr(levels), ptr(root))
    idx = np.array([0, 1, 2, 255, 256, n - 1, 77], np.uint64)
    sib = np.zeros((idx.size, lg, 4), np.uint64)
    ctx.call("pmt_simple_tree_prove", ptr(levels), n, ptr(idx), idx.size, ptr(sib))
    for q, i in enumerate(idx):
        assert np.array_equal(sib[q], oracle.simple_tree_proof(levels, n, int(i)))
    ok = np.zeros(idx.size, np.uint8)
    lv = np.ascontiguousarray(leaves[idx.astype(np.int64)])
    ctx.call("pmt_simple_tree_verify", ptr(lv), ptr(idx), idx.size, ptr(root), ptr(sib), lg, p8(ok))
    assert ok.all()
    bad = sib.copy(); bad[3, 4, 1] ^= np.uint64(1)
    ctx.call("pmt_simple_tree_verify", ptr(lv), ptr(idx), idx.size, ptr(root), ptr(bad), lg, p8(ok))
    assert list(ok) == [1, 1, 1, 0, 1, 1, 1]
    with pytest.raises(PmtError) as e:
        ctx.call("pmt_simple_tree_prove", ptr(levels), n, ptr(np.array([n], np.uint64)), 1, ptr(sib))   # :56
    assert e.value.code == PMT_E_RANGE
    # plonky2 tree, wide leaves, cap 3
    n, w, h = 1 << 8, 11, 3
    rows = splitmix_felts(4, n * w).reshape(n, w)
    dig = np.zeros((2 * (n - (1 << h)), 4), np.uint64); cap = np.zeros((1 << h, 4), np.uint64)
    ctx.call("pmt_merkle_tree_build", ptr(rows), n, w, h, ptr(dig), ptr(cap))
    idx = np.array([0, 31, 32, 200, n - 1], np.uint64)
    L = 8 - h
    sib = np.zeros((idx.size, L, 4), np.uint64)
    ctx.call("pmt_merkle_prove", ptr(dig), n, h, ptr(idx), idx.size, ptr(sib))
    for q, i in enumerate(idx):
        assert np.array_equal(sib[q], oracle.merkle_prove(dig, n, h, int(i)))
    ok = np.zeros(idx.size, np.uint8)
    rr = np.ascontiguousarray(rows[idx.astype(np.int64)])
    ctx.call("pmt_merkle_verify", ptr(rr), w, ptr(idx), idx.size, ptr(cap), h, ptr(sib), L, p8(ok))
    assert ok.all()
    ctx.call("pmt_merkle_verify", ptr(rr), w, ptr(idx ^ np.uint64(1)), idx.size, ptr(cap), h, ptr(sib), L, p8(ok))
    assert not ok.any()
    # MMR, ragged size
    nl = 1000
    leaves = splitmix_felts(5, nl)
    el = np.zeros((2 * nl - bin(nl).count("1"), 4), np.uint64)
    ctx.call("pmt_mmr_extend", ptr(el), 0, ptr(leaves), nl)
    assert np.array_equal(el, oracle.mmr_extend(None, leaves))
    idx = np.array([0, 511, 512, 900, 992, 999], np.uint64)
    sib = np.zeros((idx.size, 32, 4), np.uint64); left = np.zeros((idx.size, 32), np.uint8); ln = np.zeros(idx.size, np.uint32)
    ctx.call("pmt_mmr_prove", ptr(el), nl, ptr(idx), idx.size, ptr(sib), p8(left), ln.ctypes.data_as(u32p))
    for q, i in enumerate(idx):
        osib, oleft = oracle.mmr_subtree_proof(el, 2 * int(i) - bin(int(i)).count("1"))
        assert ln[q] == len(oleft) and np.array_equal(sib[q, :ln[q]], osib) and list(left[q, :ln[q]]) == list(oleft)
    peaks = np.zeros((64, 4), np.uint64); npk = C.c_uint32(0)
    ctx.call("pmt_mmr_peaks", ptr(el), nl, ptr(peaks), C.byref(npk))
    peaks = np.ascontiguousarray(peaks[:npk.value])
    bag = np.zeros(4, np.uint64)
    ctx.call("pmt_mmr_bag", ptr(el), nl, ptr(bag))
    st = np.zeros(idx.size, np.int8)
    lv = np.ascontiguousarray(leaves[idx.astype(np.int64)])
    ctx.call("pmt_mmr_verify", ptr(lv), idx.size, ptr(sib), p8(left), ln.ctypes.data_as(u32p), ptr(peaks), npk.value, ptr(bag),
             st.ctypes.data_as(i8p))
    assert (st == 1).all()
    lv2 = lv.copy(); lv2[2] ^= np.uint64(1)          # wrong leaf: the subtree root is not a peak -> the reference asserts (:245)
    wrong_root = bag.copy(); wrong_root[0] ^= np.uint64(1)
    ctx.call("pmt_mmr_verify", ptr(lv2), idx.size, ptr(sib), p8(left), ln.ctypes.data_as(u32p), ptr(peaks), npk.value, ptr(bag),
             st.ctypes.data_as(i8p))
    assert list(st) == [1, 1, -1, 1, 1, 1]
    ctx.call("pmt_mmr_verify", ptr(lv), idx.size, ptr(sib), p8(left), ln.ctypes.data_as(u32p), ptr(peaks), npk.value, ptr(wrong_root),
             st.ctypes.data_as(i8p))
    assert (st == 0).all()
    with pytest.raises(PmtError) as e:
        ctx.call("pmt_mmr_prove", ptr(el), nl, ptr(np.array([nl], np.uint64)), 1, ptr(sib), p8(left), ln.ctypes.data_as(u32p))
    assert e.value.code == PMT_E_RANGE


# ---- narrow leaves, big trees: the leaf copy is fused into level 1 (k_leaves_level1) ------------------------------------------
@pytest.mark.parametrize("w,h", [(1, 0), (3, 2), (4, 0), (4, 5)])
def test_fused_leaf_level_plonky2(ctx, api, oracle, w, h):
    n = 1 << 15
    rows = splitmix_felts(500 + w, n * w).reshape(n, w)
    rows[0, 0] = np.uint64(2**64 - 1); rows[n - 1, w - 1] = np.uint64(P)      # non-canonical inputs
    t = api.mt.MerkleTree.new(rows, h, ctx)
    dg, cap = oracle.merkle_tree_new(rows, h, threads=oracle.max_threads(), fast=True)
    assert np.array_equal(t.digests, dg) and np.array_equal(t.cap, cap)


def test_fused_leaf_level_simple_tree_and_mmr(ctx, api, oracle):
    n = 1 << 15
    leaves = splitmix_felts(600, n + 1)
    leaves[5] = np.uint64(P + 3)
    t = api.smt.MerkleTree.build(leaves[:n], ctx)
    levels, root = oracle.simple_tree_build(leaves[:n])
    assert np.array_equal(np.concatenate(t.tree), levels) and np.array_equal(t.root, root)
    m = api.mmr.MMR.new(ctx)
    m.extend(leaves[:6])            # even n_before, odd batch: fused pass + the unpaired last leaf
    m.extend(leaves[6:])
    assert np.array_equal(m.elements, oracle.mmr_extend(None, leaves))


# ---- batched verification: cooperative (<= 2^14 proofs) and thread-per-proof (larger) kernels agree ---------------------------
def test_verify_batches_cooperative_and_thread_kernels_agree(ctx, api, oracle):
    rnd = np.random.default_rng(5)
    # MMR, ragged size: several mountains, path lengths 0..9
    n = 1000 + 1
    leaves = splitmix_felts(77, n)
    m = api.mmr.MMR.new(ctx)
    m.extend(leaves)
    peaks, root = m.get_peaks(), m.bagging_the_peaks()
    big = 20000
    idx = rnd.integers(0, n, size=big).astype(np.uint64)
    idx[:3] = [0, n - 1, 512]                       # n - 1: the single-leaf mountain (empty path)
    sib, left, ln = m.prove_batch(idx)
    claimed = leaves[idx.astype(np.int64)].copy()
    bad_leaf = rnd.random(big) < 0.1
    claimed[bad_leaf] ^= np.uint64(1)
    st_big = api.mmr.verify_batch(claimed, sib, left, ln, peaks, root, ctx)                    # thread-per-proof kernel
    st_small = np.concatenate([api.mmr.verify_batch(claimed[a:a + 5000], sib[a:a + 5000], left[a:a + 5000], ln[a:a + 5000], peaks, root, ctx)
                               for a in range(0, big, 5000)])                                  # cooperative kernel
    assert np.array_equal(st_big, st_small)
    assert (st_big[~bad_leaf] == 1).all() and (st_big[bad_leaf] == -1).all()   # a wrong leaf misses every peak: assert! :245
    wrong_root = root.copy(); wrong_root[0] ^= np.uint64(1)
    assert (api.mmr.verify_batch(claimed[:100], sib[:100], left[:100], ln[:100], peaks, wrong_root, ctx)[~bad_leaf[:100]] == 0).all()
    for q in (0, 1, 2, 7):
        want = oracle.mmr_verify(claimed[q], root, sib[q, :ln[q]], left[q, :ln[q]], peaks)
        assert st_big[q] == want
    # plonky2 tree with a cap and sponge-hashed leaves
    lg, w, h = 10, 9, 2
    rows = splitmix_felts(78, (1 << lg) * w).reshape(1 << lg, w)
    t = api.mt.MerkleTree.new(rows, h, ctx)
    idx = rnd.integers(0, 1 << lg, size=big).astype(np.uint64)
    proofs = t.prove_batch(idx)
    claimed = rows[idx.astype(np.int64)].copy()
    bad = rnd.random(big) < 0.1
    claimed[bad, 3] ^= np.uint64(2)
    ok_big = api.mt.verify_merkle_proofs_to_cap(claimed, idx, t.cap, h, proofs, ctx)
    ok_small = np.concatenate([api.mt.verify_merkle_proofs_to_cap(claimed[a:a + 4000], idx[a:a + 4000], t.cap, h, proofs[a:a + 4000], ctx)
                               for a in range(0, big, 4000)])
    assert np.array_equal(ok_big, ok_small) and np.array_equal(ok_big, ~bad)
    for q in range(4):
        assert bool(ok_big[q]) == oracle.merkle_verify_to_cap(claimed[q], int(idx[q]), t.cap, h, proofs[q])
    # wrong index / index beyond the tree
    assert not api.mt.verify_merkle_proofs_to_cap(rows[5:6], [4], t.cap, h, proofs[:1] * 0 + t.prove_batch([5]), ctx)[0]
    assert not api.mt.verify_merkle_proofs_to_cap(rows[5:6], [5 + (1 << lg)], t.cap, h, t.prove_batch([5]), ctx)[0]


# ---- sharded build (virtual ranks on one GPU) ---------------------------------------------------------------------------
@pytest.mark.parametrize("lg,w,h,G", [(10, 4, 0, 8), (10, 4, 1, 4), (9, 135, 4, 8), (8, 4, 3, 8), (6, 1, 0, 2)])
def test_sharded_virtual_ranks(ctx, api, oracle, lg, w, h, G):
    from plonky2_merkle_trees_b200 import sharded
    from plonky2_merkle_trees_b200.device import to_device, to_host
    n = 1 << lg
    rows = splitmix_felts(31 * lg + G, n * w).reshape(n, w)
    eng = sharded.CudaEngine(ctx)
    g = G.bit_length() - 1
    chunks, caps = [], []
    for r in range(G):
        s, c = sharded.shard_range(n, G, r)
        dg, cp = eng.build_local(to_device(rows[s:s + c], eng.device), max(h - g, 0))
        eng.sync()
        chunks.append(to_host(dg)); caps.append(to_host(cp))
    gathered = np.concatenate(caps)
    odg, ocap = oracle.merkle_tree_new(rows, h, threads=4, fast=True)
    if h >= g:
        t = sharded.ShardedMerkleTree(n, w, h, G, 0, None, None, None, gathered)
        assert np.array_equal(gathered, ocap)
    else:
        d_top = eng.top_levels(to_device(gathered, eng.device), h)
        eng.sync()   # the ctx stream is non-blocking with respect to torch's stream: sync before the D2H copy
        top = to_host(d_top)
        t = sharded.ShardedMerkleTree(n, w, h, G, 0, None, gathered, top, top[top.shape[0] - (1 << h):])
        assert np.array_equal(t.cap, ocap)
    assert np.array_equal(t.assemble_global(chunks), odg)


# ---- the tail in one launch (k_tree_coop: block-local 64-node subtrees + ticket rounds) against one cooperative launch per level ----
@pytest.mark.parametrize("lg,w,h", [(5, 4, 0), (6, 4, 1), (9, 4, 0), (10, 1, 3), (12, 4, 0), (14, 4, 0), (14, 9, 6), (15, 4, 11), (17, 4, 2)])
def test_fused_subtree_blocks_equal_per_level_launches(api, oracle, lg, w, h, monkeypatch):
    """all levels of at most 2^13 nodes of a perfect tree run in ONE k_tree_coop launch (block-local 64-node subtrees, then two
    rounds of tickets); PMT_FUSE_SUBTREES=0 keeps one cooperative launch per level: both must give the oracle's tree (plonky2
    layout, simple tree), the fused plan with fewer launches"""
    from plonky2_merkle_trees_b200 import _lib
    n = 1 << lg
    rows = splitmix_felts(300 + lg + w + h, n * w).reshape(n, w)
    monkeypatch.setenv("PMT_FUSE_SUBTREES", "0")
    plain = _lib.Context(0)
    monkeypatch.delenv("PMT_FUSE_SUBTREES")
    fused = _lib.Context(0)
    try:
        l0 = plain.launches
        a = api.mt.MerkleTree.new(rows, h, plain)
        l_plain = plain.launches - l0
        l0 = fused.launches
        b = api.mt.MerkleTree.new(rows, h, fused)
        l_fused = fused.launches - l0
        assert np.array_equal(a.digests, b.digests) and np.array_equal(a.cap, b.cap)
        if lg <= 15:
            odg, ocap = oracle.merkle_tree_new(rows, h, threads=4, fast=True)
            assert np.array_equal(b.digests, odg) and np.array_equal(b.cap, ocap)
        if lg - h >= 7:
            assert l_fused < l_plain, (l_fused, l_plain)
        if w == 1:
            s1 = api.smt.MerkleTree.build(rows[:, 0], plain)
            s2 = api.smt.MerkleTree.build(rows[:, 0], fused)
            assert np.array_equal(np.concatenate(s1.tree), np.concatenate(s2.tree)) and np.array_equal(s1.root, s2.root)
    finally:
        plain.close(); fused.close()


@pytest.mark.parametrize("n0,m", [(0, 1 << 14), (3 << 14, 1 << 14), (1 << 9, 3 << 9), (48, 1 << 10), (5, 40000), (1 << 15, (1 << 15) + 16)])
def test_fused_subtree_blocks_in_mmr_appends(api, oracle, n0, m, monkeypatch):
    """batch appends whose level ranges are aligned to 16 nodes take the fused blocks too; ragged ones stay per level"""
    from plonky2_merkle_trees_b200 import _lib
    leaves = splitmix_felts(900 + n0 + m, n0 + m)
    monkeypatch.setenv("PMT_FUSE_SUBTREES", "0")
    plain = _lib.Context(0)
    monkeypatch.delenv("PMT_FUSE_SUBTREES")
    fused = _lib.Context(0)
    try:
        got = []
        for c in (plain, fused):
            mm = api.mmr.MMR.new(c)
            if n0:
                mm.extend(leaves[:n0])
            mm.extend(leaves[n0:])
            got.append(mm.elements.copy())
        assert np.array_equal(got[0], got[1])
        assert np.array_equal(got[1], oracle.mmr_extend(None, leaves))
    finally:
        plain.close(); fused.close()


# ---- single-process multi-GPU build (one ctx per device; here: several ctxs on the devices that exist) ---------------------
@pytest.fixture(scope="module")
def ctx_pool():
    """8 distinct contexts spread over the visible devices (all on cuda:0 on a one-GPU box: the sharding, the slices of
    upstream's layout and the finish are the same code either way)"""
    from plonky2_merkle_trees_b200 import _lib
    ndev = torch.cuda.device_count()
    pool = [_lib.Context(i % ndev) for i in range(8)]
    yield pool
    for c in pool:
        c.close()


@pytest.mark.parametrize("lg,w,h,G", [(10, 4, 0, 8), (10, 4, 1, 4), (10, 4, 3, 8), (9, 135, 4, 8), (8, 4, 5, 8), (6, 1, 0, 2),
                                      (3, 4, 0, 8), (3, 4, 3, 8), (4, 9, 2, 4), (5, 4, 0, 1), (14, 7, 0, 4)])
def test_multi_context_build_equals_oracle(ctx_pool, api, oracle, lg, w, h, G):
    n = 1 << lg
    rows = splitmix_felts(100 + lg + w + h + G, n * w).reshape(n, w)
    rows[0, 0] = np.uint64(2**64 - 1)   # non-canonical input
    dg, cap = api.mt.MerkleTree.new_multi(rows, h, ctx_pool[:G])
    odg, ocap = oracle.merkle_tree_new(rows, h, threads=4, fast=True)
    assert np.array_equal(cap, ocap)
    assert np.array_equal(dg, odg)


@pytest.mark.parametrize("lg,w,h,G", [(20, 4, 0, 4), (21, 4, 1, 8), (19, 4, 5, 4), (18, 135, 2, 2)])
def test_multi_context_build_pipelined_sizes(ctx_pool, api, lg, w, h, G):
    """every context takes the chunked copy / hash / copy pipeline; result = the one-piece device build (tied to the oracle above)"""
    n = 1 << lg
    rows = splitmix_felts(7 + lg + w, n * w).reshape(n, w)
    dg, cap = api.mt.MerkleTree.new_multi(rows, h, ctx_pool[:G])
    t = api.mt.MerkleTree.new(rows, h)
    assert np.array_equal(cap, t.cap) and np.array_equal(dg, t.digests)
    dg2, cap2 = api.mt.MerkleTree.new_multi(rows, h, ctx_pool[:G])   # arenas and streams are reused
    assert np.array_equal(dg, dg2) and np.array_equal(cap, cap2)


def test_multi_context_build_errors(ctx_pool, api):
    from plonky2_merkle_trees_b200._lib import PmtError, PMT_E_NOT_POW2, PMT_E_RANGE, PMT_E_INVALID_ARG
    rows = np.zeros((8, 4), np.uint64)
    for bad, code in ((ctx_pool[:3], PMT_E_NOT_POW2), ([ctx_pool[0], ctx_pool[0]], PMT_E_INVALID_ARG)):
        with pytest.raises(PmtError) as e:
            api.mt.MerkleTree.new_multi(rows, 0, bad)
        assert e.value.code == code
    with pytest.raises(PmtError) as e:
        api.mt.MerkleTree.new_multi(rows[:4], 0, ctx_pool[:8])      # more contexts than leaves
    assert e.value.code == PMT_E_RANGE
    with pytest.raises(PmtError) as e:
        api.mt.MerkleTree.new_multi(rows[:6], 0, ctx_pool[:2])      # log2_strict
    assert e.value.code == PMT_E_NOT_POW2
    with pytest.raises(PmtError) as e:
        api.mt.MerkleTree.new_multi(rows, 4, ctx_pool[:2])          # cap_height > log2 n
    assert e.value.code == PMT_E_RANGE


@pytest.mark.parametrize("n0,m,G", [(0, 1 << 15, 2), (0, 1 << 17, 8), (12345, (1 << 15) + 77, 2), (4096, 3 * 4096, 3), (4095, 2 * 4096 + 2, 2),
                                    (1 << 14, (1 << 16) - 1, 4), (7, 40000, 4), (3 << 12, 1 << 13, 2), (100, 5000, 2), (0, 1 << 13, 1),
                                    (0, (1 << 18) + 5, 8), (999, 1 << 18, 5)])
def test_multi_context_mmr_extend_equals_oracle(ctx_pool, api, oracle, n0, m, G):
    """pmt_mmr_extend_multi: head on context 0, aligned blocks as sub-mountains built from empty by context j mod G, the coarse
    MMR over the block roots finished on context 0, tail -- element by element the sequential MMR"""
    leaves = splitmix_felts(5000 + n0 + m + G, n0 + m)
    leaves[n0] = np.uint64(2**64 - 1)     # non-canonical input
    if n0 + m <= 1 << 16:
        want = oracle.mmr_extend(None, leaves)
    else:                                  # the device-resident single-context append (tied to the oracle by the tests above)
        one = api.mmr.MMR.new(ctx_pool[0])
        one.extend(leaves)
        want = one.elements
    before = want[:2 * n0 - bin(n0).count("1")].copy()
    got = api.mmr.extend_multi(before if n0 else None, n0, leaves[n0:], ctx_pool[:G])
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert (api.mmr.multi_plan(n0, m, G) is not None) == (G > 1 and m // G >= 4096 and ((n0 + m) >> 12) > ((n0 + 4095) >> 12))


def test_multi_context_mmr_extend_errors(ctx_pool, api):
    from plonky2_merkle_trees_b200._lib import PmtError, PMT_E_INVALID_ARG, PMT_E_RANGE
    with pytest.raises(PmtError) as e:
        api.mmr.extend_multi(None, 0, np.arange(10, dtype=np.uint64), [ctx_pool[1], ctx_pool[1]])
    assert e.value.code == PMT_E_INVALID_ARG
    assert api.mmr.extend_multi(None, 0, [], ctx_pool[:2]).shape == (0, 4)
    import ctypes as C
    handles = (C.c_void_p * 2)(ctx_pool[0].h, ctx_pool[1].h)
    buf = np.zeros((4, 4), np.uint64)
    leaves = np.zeros(1, np.uint64)
    rc = ctx_pool[0].lib.pmt_mmr_extend_multi(handles, 2, buf.ctypes.data_as(C.POINTER(C.c_uint64)), (1 << 30), leaves.ctypes.data_as(C.POINTER(C.c_uint64)), 1)
    assert rc == PMT_E_RANGE


@pytest.mark.parametrize("n_roots,h", [(2, 0), (2, 1), (8, 0), (8, 2), (64, 0), (1024, 3), (4096, 0), (8192, 1)])
def test_top_levels_above_gathered_roots(ctx, oracle, n_roots, h):
    """pmt_top_levels_dev (TopRoots layout): the levels above 2 .. 8192 gathered subtree roots, one k_tree_coop launch while a
    level has at most 2^13 nodes"""
    from plonky2_merkle_trees_b200 import sharded
    from plonky2_merkle_trees_b200.device import to_device, to_host
    roots = splitmix_felts(n_roots + h, n_roots * 4).reshape(n_roots, 4)
    eng = sharded.CudaEngine(ctx)
    d_top = eng.top_levels(to_device(roots, eng.device), h)
    eng.sync()
    got = to_host(d_top)
    cur, want = roots, []
    while cur.shape[0] > (1 << h):
        cur = oracle.two_to_one_batch(cur[0::2], cur[1::2])
        want.append(cur)
    want = np.concatenate(want) if want else np.zeros((0, 4), np.uint64)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_top_levels_batch(ctx, oracle):
    from plonky2_merkle_trees_b200 import sharded
    from plonky2_merkle_trees_b200.device import to_device, to_host
    eng = sharded.CudaEngine(ctx)
    for b, g in [(1, 2), (21, 8), (5, 64), (3, 1)]:
        roots = splitmix_felts(b * 100 + g, b * g * 4).reshape(b, g, 4)
        d_top = eng.top_levels_batch(to_device(roots, eng.device).view(b, g, 4))
        eng.sync()
        got = to_host(d_top.contiguous()).reshape(b, g - 1, 4)
        for i in range(b):
            cur, want = roots[i], []
            while cur.shape[0] > 1:
                cur = oracle.two_to_one_batch(cur[0::2], cur[1::2]); want.append(cur)
            want = np.concatenate(want) if want else np.zeros((0, 4), np.uint64)
            assert np.array_equal(got[i], want), (b, g, i)


def test_sharded_mmr_virtual_world_of_one(ctx, api, oracle):
    """build_sharded_mmr without a process group (world 1): rounds = the set bits of n, no tail; elements, peaks, bag and
    proofs against the sequential add_leaf oracle.  (world > 1: tests/test_abi_and_host.py on gloo, tools/multigpu_check.py)"""
    from plonky2_merkle_trees_b200 import sharded
    from plonky2_merkle_trees_b200.device import to_device
    eng = sharded.CudaEngine(ctx)
    for n in (1, 2, 77, 1000, 4097):
        leaves = splitmix_felts(n, n)
        sm = sharded.build_sharded_mmr(to_device(leaves, eng.device), n, eng)
        want = oracle.mmr_extend(None, leaves)
        assert np.array_equal(sm.assemble_global([sm.local.elements], None), want)
        assert np.array_equal(sm.get_peaks(), oracle.mmr_peaks(want))
        assert np.array_equal(sm.bagging_the_peaks(), oracle.mmr_bag(want))
        for i in {0, n // 2, n - 1}:
            assert sm.get_proof_normal_index(i).verify(int(leaves[i]), sm.bagging_the_peaks(), ctx)


def test_sharded_mmr_with_nccl_inside_the_library_world_of_one(api, oracle):
    """pmt_mmr_build_sharded_dev on a communicator of ONE rank (several ranks: tools/multigpu_check.py under torchrun): one
    library call builds the sub-mountains, gathers their roots and finishes the peaks; elements, peaks, bag and proofs against
    the sequential add_leaf oracle.  Also the call's argument checks."""
    import ctypes as C
    from plonky2_merkle_trees_b200 import _lib, sharded
    from plonky2_merkle_trees_b200.device import to_device
    c = _lib.Context(0)
    try:
        eng = sharded.CudaEngine(c)
        with pytest.raises(_lib.PmtError):                               # no communicator yet
            eng.build_sharded_mmr(to_device(splitmix_felts(1, 8), eng.device), 8, 1, 0)
        uid = C.create_string_buffer(128)
        c.call("pmt_nccl_unique_id", C.cast(uid, C.c_void_p))
        c.call("pmt_comm_init", C.cast(uid, C.c_void_p), 0, 1)
        eng.has_comm = True
        for n in (1, 2, 77, 1000, 4097, (1 << 15) + 5):
            leaves = splitmix_felts(n + 3, n)
            local, tail, d_roots, d_tops, d_peaks = eng.build_sharded_mmr(to_device(leaves, eng.device), n, 1, 0)
            assert tail is None and d_tops.shape[1] == 0
            sm = sharded.ShardedMMR(n, 1, 0, sharded.mmr_shard_plan(n, 1), local, d_roots, d_tops, tail, d_peaks, eng)
            want = oracle.mmr_extend(None, leaves)
            assert np.array_equal(sm.assemble_global([sm.local.elements], None), want)
            assert np.array_equal(sm.get_peaks(), oracle.mmr_peaks(want))
            assert np.array_equal(sm.bagging_the_peaks(), oracle.mmr_bag(want))
            assert len(sm.rounds) == bin(n).count("1")
            for i in {0, n // 2, n - 1}:
                assert sm.get_proof_normal_index(i).verify(int(leaves[i]), sm.bagging_the_peaks(), c)
        with pytest.raises(_lib.PmtError):                               # more than 2^30 leaves
            c.call("pmt_mmr_build_sharded_dev", None, (1 << 30) + 1, None, None, None, None, None)
    finally:
        c.close()


# ---- tree fed by the prover's column-major LDE values -----------------------------------------------------
def _reverse_index_bits(rows):
    """[UPSTREAM] plonky2_util::reverse_index_bits: out[i] = in[bit_reverse(i, log2 n)]"""
    n = rows.shape[0]
    lg = n.bit_length() - 1
    idx = np.array([int(format(i, "0%db" % lg)[::-1], 2) if lg else 0 for i in range(n)], dtype=np.int64)
    return rows[idx]


@pytest.mark.parametrize("lg,w,h,rev", [(1, 1, 0, True), (4, 4, 0, True), (6, 5, 2, True), (8, 135, 4, True), (8, 135, 4, False),
                                        (10, 9, 0, True), (12, 4, 3, True), (5, 3, 5, True), (0, 7, 0, True)])
def test_tree_from_columns_equals_transpose_then_build(ctx, api, oracle, lg, w, h, rev):
    from plonky2_merkle_trees_b200.device import to_device
    n = 1 << lg
    cols = splitmix_felts(97 * lg + w, w * n).reshape(w, n)
    cols[0, 0] = np.uint64(2**64 - 1)          # non-canonical inputs are canonicalised in the no-op leaf rule
    if n > 1:
        cols[w - 1, n - 1] = np.uint64(P)
    rows = np.ascontiguousarray(cols.T)
    if rev:
        rows = _reverse_index_bits(rows)
    dg, cap = oracle.merkle_tree_new(rows, h, threads=2, fast=True)
    t = api.mt.MerkleTree.from_columns_dev(to_device(cols, "cuda:0"), h, bit_reverse=rev, keep_leaves=True, ctx=ctx)
    assert np.array_equal(t.cap, cap)
    assert np.array_equal(t.digests, dg)
    assert np.array_equal(t.leaves, rows)      # MerkleTree.leaves as upstream keeps them (raw felts, row-major)
    t2 = api.mt.MerkleTree.from_columns_dev(to_device(cols, "cuda:0"), h, bit_reverse=rev, keep_leaves=False, ctx=ctx)
    assert np.array_equal(t2.digests, dg) and np.array_equal(t2.cap, cap)
    if lg > h:                                 # openings verify against the cap
        i = n // 3
        assert api.mt.verify_merkle_proof_to_cap(rows[i], i, t.cap, h, t.prove(i), ctx)


def test_tree_from_columns_errors(ctx, api):
    from plonky2_merkle_trees_b200 import _lib
    from plonky2_merkle_trees_b200.device import to_device
    with pytest.raises(_lib.PmtError) as e:
        api.mt.MerkleTree.from_columns_dev(to_device(np.zeros((4, 6), np.uint64), "cuda:0"), 0, ctx=ctx)
    assert e.value.code == _lib.PMT_E_NOT_POW2
    with pytest.raises(_lib.PmtError) as e:
        api.mt.MerkleTree.from_columns_dev(to_device(np.zeros((4, 8), np.uint64), "cuda:0"), 4, ctx=ctx)
    assert e.value.code == _lib.PMT_E_RANGE


# ---- stream ordering between torch and the ctx stream ---------------------------------------------------------------
def test_dev_entry_points_wait_for_torch_stream(api, ctx, oracle):
    """Leaves produced by torch kernels on torch's current stream, handed to a *_dev entry point with no host sync in
    between: the ctx stream (non-blocking) must be ordered after torch's stream (plonky2_merkle_trees_b200/device.py
    order_after_torch).  Found by tools/multigpu_check.py: without the wait the build read half-written leaves."""
    import bench
    n, w = 1 << 16, 4
    want = None
    for rep in range(4):
        # a long-running torch kernel queue in front of the generator makes the race (if any) certain
        junk = torch.randn(4096, 4096, device="cuda")
        for _ in range(8):
            junk = junk @ junk * 1e-3
        d_leaves = bench.splitmix_torch(rep * n * w, n * w, torch.device("cuda", 0)).view(n, w)
        t = api.mt.MerkleTree.new_dev(d_leaves, 0, ctx)
        rows = bench.splitmix_numpy(rep * n * w, n * w).reshape(n, w)
        dg, cap = oracle.merkle_tree_new(rows, 0, threads=oracle.max_threads(), fast=True)
        assert np.array_equal(t.cap, cap), rep
        assert np.array_equal(t.digests, dg), rep


# =====================================================================================================================
# round 2: the four-threads-per-state cooperative kernels, the one-launch tree tail, ABI hardening, full-size parity
# =====================================================================================================================
def test_external_stream_and_switching_back(api, oracle):
    """pmt_set_stream: builds enqueued on a caller-owned CUDA stream (a torch stream), then on the ctx's own stream again;
    changing the stream drains the old one first, so results are complete and equal the oracle's either way."""
    from plonky2_merkle_trees_b200 import _lib
    from plonky2_merkle_trees_b200.device import dev_u64, dptr, to_device, to_host
    c = _lib.Context(0)
    try:
        side = torch.cuda.Stream(device="cuda:0")
        for lg, use_side in [(12, True), (16, False), (18, True), (10, False)]:
            n = 1 << lg
            rows = splitmix_felts(40 + lg, n * 4).reshape(n, 4)
            d_rows = to_device(rows, "cuda:0")
            torch.cuda.synchronize()
            c.set_stream(side.cuda_stream if use_side else None)
            d_dig, d_cap = dev_u64((2 * n - 2, 4), "cuda:0"), dev_u64((1, 4), "cuda:0")
            torch.cuda.synchronize()
            c.call("pmt_merkle_tree_build_dev", dptr(d_rows), n, 4, 0, dptr(d_dig), dptr(d_cap))
            c.set_stream(None if use_side else side.cuda_stream)          # leaves the stream the build runs on: must drain it
            odg, ocap = oracle.merkle_tree_new(rows, 0, threads=4, fast=True)
            assert np.array_equal(to_host(d_dig), odg) and np.array_equal(to_host(d_cap), ocap)
        c.set_stream(None)
    finally:
        c.close()


def test_hasher_small_and_big_batches_agree(api, oracle):
    """Hasher batches of <= 2^12 rows run by quads (k_permute_coop / k_rows_coop), bigger ones one row per thread: both against
    the oracle and against each other, incl. non-canonical inputs"""
    st = edge_states(6000, 11)
    big = api.hasher.permute(st)                       # thread-per-state
    small = np.concatenate([api.hasher.permute(st[a:a + 1500]) for a in range(0, 6000, 1500)])   # quads
    assert np.array_equal(big, small)
    for i in range(0, 6000, 37):
        assert big[i].tolist() == oracle.permute(st[i]).tolist()
    for w in (5, 8, 9, 135):
        rows = splitmix_felts(40 + w, 5000 * w).reshape(5000, w)
        rows[7, :] = np.uint64(2**64 - 1)
        big = api.hasher.hash_or_noop(rows)
        small = np.concatenate([api.hasher.hash_or_noop(rows[a:a + 1000]) for a in range(0, 5000, 1000)])
        assert np.array_equal(big, small)
        assert np.array_equal(api.hasher.hash_no_pad(rows[:100]), big[:100])
        for i in (0, 7, 63, 64, 999, 4999):
            assert big[i].tolist() == oracle.hash_no_pad(rows[i]).tolist()
    rows = splitmix_felts(3, 100 * 3).reshape(100, 3)
    assert api.hasher.hash_no_pad(rows)[5].tolist() == oracle.hash_no_pad(rows[5]).tolist()     # sponge also for width <= 4


@pytest.mark.parametrize("coop_log2", [6, 7, 10, 15, 17])
def test_tree_tail_for_every_cooperative_threshold(api, oracle, coop_log2, monkeypatch):
    """k_tree_coop (block-local subtrees of 64 nodes, then the last block to finish takes the levels above) for trees whose
    cooperative part starts at 2^6 .. 2^17 nodes: 1, 2, 16, 512, 2048 blocks in the ticket.  Plonky2 layout with and without
    a cap, the simple tree, MMR appends (aligned, ragged, onto old peaks) -- all against the oracle."""
    from plonky2_merkle_trees_b200 import _lib
    monkeypatch.setenv("PMT_COOP_MAX_LOG2", str(coop_log2))
    c = _lib.Context(0)
    monkeypatch.delenv("PMT_COOP_MAX_LOG2")
    try:
        for lg, w, h in [(3, 4, 0), (7, 4, 0), (8, 4, 1), (12, 4, 0), (14, 3, 5), (16, 4, 0), (18, 1, 0), (11, 9, 0)]:
            n = 1 << lg
            rows = splitmix_felts(1000 * coop_log2 + lg + w, n * w).reshape(n, w)
            t = api.mt.MerkleTree.new(rows, h, c)
            dg, cap = oracle.merkle_tree_new(rows, h, threads=oracle.max_threads(), fast=True)
            assert np.array_equal(t.digests, dg) and np.array_equal(t.cap, cap), (lg, w, h)
            t2 = api.mt.MerkleTree.new(rows, h, c)      # the tickets are reset by the launch that used them
            assert np.array_equal(t2.digests, dg) and np.array_equal(t2.cap, cap)
        leaves = splitmix_felts(coop_log2, 1 << 13)
        s = api.smt.MerkleTree.build(leaves, c)
        levels, root = oracle.simple_tree_build(leaves)
        assert np.array_equal(np.concatenate(s.tree), levels) and np.array_equal(s.root, root)
        for n0, m in [(0, 1 << 13), (1 << 12, 1 << 12), (3 << 10, 5 << 10), (77, 9000), (1 << 14, 64), (4032, 64 + 4096)]:
            lv = splitmix_felts(n0 + m + coop_log2, n0 + m)
            mm = api.mmr.MMR.new(c)
            if n0:
                mm.extend(lv[:n0])
            mm.extend(lv[n0:])
            assert np.array_equal(mm.elements, oracle.mmr_extend(None, lv)), (n0, m)
    finally:
        c.close()


def test_mmr_verify_rejects_overlong_path(ctx, api):
    """a proof is the verifier's untrusted input: path_len beyond the 32 entries of the batch layout is status -1 (the
    reference would fail its assert! :245 on such a path), not an out-of-bounds walk -- both kernels"""
    n = 1000
    leaves = splitmix_felts(8, n)
    m = api.mmr.MMR.new(ctx); m.extend(leaves)
    peaks, root = m.get_peaks(), m.bagging_the_peaks()
    for q in (64, 20000):      # cooperative / thread-per-proof kernel
        idx = (splitmix_felts(q, q) % np.uint64(n)).astype(np.uint64)
        sib, left, ln = m.prove_batch(idx)
        ln = ln.copy()
        ln[3] = 33; ln[5] = 0xFFFFFFFF; ln[q - 1] = 1 << 20
        st = api.mmr.verify_batch(leaves[idx.astype(np.int64)], sib, left, ln, peaks, root, ctx)
        bad = np.zeros(q, bool); bad[[3, 5, q - 1]] = True
        assert (st[bad] == -1).all() and (st[~bad] == 1).all()


def test_simple_tree_verify_index_semantics_and_prove_bounds(ctx, api, oracle):
    """verify_merkle_proof (simple_merkle_tree.rs:91-109) folds by the parities of the low path_len index bits and never
    looks above them: leaf_index + k 2^path_len verifies in the reference, so it does here (device and host forms), while
    upstream's verify_merkle_proof_to_cap rejects it.  A device prove with an index >= n returns zeros, never reads outside."""
    from plonky2_merkle_trees_b200._lib import ptr, u8p
    from plonky2_merkle_trees_b200.device import dev_u64, dptr, to_device, to_host
    n, lg = 1 << 6, 6
    leaves = splitmix_felts(21, n)
    t = api.smt.MerkleTree.build(leaves, ctx)
    proof = t.get_merkle_proof(9)
    for idx in (9, 9 + n, 9 + 5 * n, 9 + (1 << 40)):
        assert api.smt.verify_merkle_proof(int(leaves[9]), idx, t.root, proof, ctx)
        assert oracle.simple_tree_verify(int(leaves[9]), idx, t.root, proof)
    assert not api.smt.verify_merkle_proof(int(leaves[9]), 8, t.root, proof, ctx)
    ok = np.zeros(2, np.uint8)
    ctx.call("pmt_simple_tree_verify", ptr(np.array([leaves[9], leaves[9]], np.uint64)), ptr(np.array([9 + n, 8], np.uint64)), 2, ptr(t.root),
             ptr(np.concatenate([proof, proof])), lg, ok.ctypes.data_as(u8p))
    assert list(ok) == [1, 0]
    levels = to_device(np.concatenate(t.tree), "cuda:0")
    d_idx = to_device(np.array([5, n, 2**63], np.uint64), "cuda:0")
    d_out = dev_u64((3, lg, 4), "cuda:0")
    ctx.call("pmt_simple_tree_prove_dev", dptr(levels), n, dptr(d_idx), 3, dptr(d_out)); ctx.sync()
    out = to_host(d_out)
    assert np.array_equal(out[0], t.get_merkle_proof(5)) and not out[1].any() and not out[2].any()
    rows = splitmix_felts(22, n * 4).reshape(n, 4)
    pt = api.mt.MerkleTree.new(rows, 2, ctx)
    d_out = dev_u64((2, lg - 2, 4), "cuda:0")
    ctx.call("pmt_merkle_prove_dev", dptr(pt.d_digests), n, 2, dptr(to_device(np.array([7, n + 7], np.uint64), "cuda:0")), 2, dptr(d_out)); ctx.sync()
    out = to_host(d_out)
    assert np.array_equal(out[0], pt.prove(7)) and not out[1].any()
    m = api.mmr.MMR.new(ctx); m.extend(leaves[:50])
    d_sib, d_left, d_len = dev_u64((2, 32, 4), "cuda:0"), torch.zeros((2, 32), dtype=torch.uint8, device="cuda:0"), torch.zeros(2, dtype=torch.int32, device="cuda:0")
    ctx.call("pmt_mmr_prove_dev", dptr(m.d_elements), 50, dptr(to_device(np.array([49, 50], np.uint64), "cuda:0")), 2, dptr(d_sib), dptr(d_left), dptr(d_len)); ctx.sync()
    assert d_len.cpu().tolist() == [1, 0]


@pytest.mark.parametrize("n0,m", [((1 << 20) + 123, 1), ((1 << 20) + 123, 5), ((1 << 20) - 1, 1), (3 << 18, 70000), ((1 << 21) - 7, 7 + (1 << 20))])
def test_mmr_host_append_touches_only_peaks_and_new_elements(ctx, api, n0, m):
    """pmt_mmr_extend on a HOST array: the device holds the old peaks and the new elements only (MmrAppend), so appending one
    leaf to a 2^20-leaf MMR is a few hundred bytes of device memory -- same elements as the device-resident append"""
    from plonky2_merkle_trees_b200._lib import ptr
    leaves = splitmix_felts(n0 + m, n0 + m)
    ref = api.mmr.MMR.new(ctx)
    ref.extend(leaves)
    want = ref.elements
    s0 = ctx.lib.pmt_mmr_size(n0)
    el = np.zeros_like(want)
    el[:s0] = want[:s0]
    ctx.call("pmt_mmr_extend", ptr(el), n0, ptr(leaves[n0:]), m)
    assert np.array_equal(el, want)


def test_full_size_headline_tree_against_full_oracle_rebuild(api, oracle):
    """The headline workload, 2^24 leaves x 4 felts, cap 0: EVERY digest (33.5 M) against a full rebuild by the oracle on all
    host cores (16.7 M permutations, a few seconds) -- and the same for the host-buffer, pipelined entry point"""
    from plonky2_merkle_trees_b200._lib import ptr
    import bench
    lg, w = 24, 4
    n = 1 << lg
    rows = bench.splitmix_numpy(0, n * w).reshape(n, w)
    odg, ocap = oracle.merkle_tree_new(rows, 0, threads=oracle.max_threads(), fast=True)
    t = api.mt.MerkleTree.new(rows, 0)
    assert np.array_equal(t.cap, ocap)
    assert np.array_equal(t.digests, odg)
    del t
    dg = np.zeros_like(odg); cap = np.zeros_like(ocap)
    api.mt._lib.default_context().call("pmt_merkle_tree_build", ptr(rows), n, w, 0, ptr(dg), ptr(cap))
    assert np.array_equal(cap, ocap) and np.array_equal(dg, odg)


def test_full_size_c4_fri_commitment_against_full_oracle_rebuild(api, oracle):
    """BASELINE C4 at FULL size: 2^20 rows x 135 columns, cap_height 4 (18.9 M permutations): every digest and the cap
    against the oracle, from row-major leaves and from the prover's column-major LDE values"""
    from plonky2_merkle_trees_b200.device import to_device
    import bench
    lg, w, h = 20, 135, 4
    n = 1 << lg
    rows = bench.splitmix_numpy(0, n * w).reshape(n, w)
    odg, ocap = oracle.merkle_tree_new(rows, h, threads=oracle.max_threads(), fast=True)
    t = api.mt.MerkleTree.new(rows, h)
    assert np.array_equal(t.cap, ocap) and np.array_equal(t.digests, odg)
    del t
    cols = np.ascontiguousarray(rows.T)
    t = api.mt.MerkleTree.from_columns_dev(to_device(cols, "cuda:0"), h, bit_reverse=False)
    assert np.array_equal(t.cap, ocap) and np.array_equal(t.digests, odg)


def test_full_size_c5_2p28_sampled_against_oracle(api, oracle):
    """BASELINE C5 on one GPU: 2^28 leaves x 4 felts (8 GiB of leaves, 16 GiB of digests).  A full CPU rebuild is 40 s, so:
    (1) the subtree over the first 2^20 leaves (a contiguous slice of upstream's layout) equals the oracle's tree, digest by
    digest; (2) 20 000 nodes sampled over all 28 levels equal the oracle's two_to_one of their stored children; (3) the root
    is the oracle's fold of the 16 subtree roots of an independent cap-4 build; (4) leaf digests are the no-op copy"""
    import bench
    from plonky2_merkle_trees_b200.device import to_host
    lg, w = 28, 4
    n = 1 << lg
    dev = torch.device("cuda", 0)
    d_leaves = bench.splitmix_torch(0, n * w, dev).view(n, w)
    t = api.mt.MerkleTree.new_dev(d_leaves, 0)
    d = t.d_digests
    node = lambda l, k: 2 * (((k >> 1) << (l + 1)) + (1 << l) - 1) + (k & 1)
    # (1)
    sub = 1 << 20
    rows = bench.splitmix_numpy(0, sub * w).reshape(sub, w)
    odg, ocap = oracle.merkle_tree_new(rows, 0, threads=oracle.max_threads(), fast=True)
    assert np.array_equal(to_host(d[:2 * sub - 2]), odg)
    assert np.array_equal(to_host(d[node(20, 0)]), ocap[0])
    # (2)
    rnd = random.Random(28)
    ls = [rnd.randrange(1, lg) for _ in range(20000)]
    ks = [rnd.randrange(n >> l) for l in ls]
    par = torch.tensor([node(l, k) for l, k in zip(ls, ks)], device=dev)
    c0 = torch.tensor([node(l - 1, 2 * k) for l, k in zip(ls, ks)], device=dev)
    assert np.array_equal(oracle.two_to_one_batch(to_host(d[c0]), to_host(d[c0 + 1])), to_host(d[par]))
    top = to_host(d[torch.tensor([node(lg - 1, 0), node(lg - 1, 1)], device=dev)])
    root = to_host(t.d_cap)[0]
    assert np.array_equal(oracle.two_to_one(top[0], top[1]), root)
    # (4)
    pick = torch.tensor([rnd.randrange(n // 2) for _ in range(5000)], device=dev)
    assert torch.equal(d[4 * pick], d_leaves[2 * pick]) and torch.equal(d[4 * pick + 1], d_leaves[2 * pick + 1])
    del t, d
    torch.cuda.empty_cache()
    # (3)
    t4 = api.mt.MerkleTree.new_dev(d_leaves, 4)
    lvl = to_host(t4.d_cap)
    while lvl.shape[0] > 1:
        lvl = oracle.two_to_one_batch(lvl[0::2], lvl[1::2])
    assert np.array_equal(lvl[0], root)
    del t4, d_leaves
    torch.cuda.empty_cache()


def test_multi_gpu_nccl_parity_under_torchrun():
    """the one-process-per-GPU path (NCCL all_gather of the roots, sharded MMR) against the single-GPU build, bit for bit:
    tools/multigpu_check.py under torchrun on every visible device -- skipped on a one-GPU box"""
    import os, subprocess, sys
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    world = 1 << (ndev.bit_length() - 1)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "multigpu_check.py")], capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": false' not in r.stdout


def test_peer_memory_exchange_on_one_device():
    """k_exchange_top -- the subtree roots pushed into the peers' mailboxes with plain stores + flags, the bounded wait, the top
    levels, all in one launch per context -- exercised on a one-GPU box by 2 .. 8 contexts that share cuda:0 (a hook: such
    contexts normally keep the copy path), 11 exchanges in a row (the ring has 4 slots), against the oracle and the copy path.
    With >= 2 GPUs the same kernel runs over NVLink in test_multi_context_device_resident_build... and under torchrun."""
    import os, subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_mailbox_same_device.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"ok": false' not in r.stdout and r.stdout.count('"exchange": "mailbox"') >= 9


@pytest.mark.parametrize("lg,w,h,G", [(10, 4, 0, 8), (10, 4, 1, 4), (10, 4, 3, 8), (9, 135, 4, 8), (8, 4, 5, 8), (6, 1, 0, 2), (5, 4, 0, 1),
                                      (15, 7, 0, 4), (16, 4, 2, 2)])
def test_multi_context_device_resident_build_with_peer_copies(ctx_pool, api, oracle, lg, w, h, G):
    """pmt_merkle_tree_build_multi_dev: every context builds its subtree from leaves on ITS device, the roots (or cap
    entries) travel with cudaMemcpyPeerAsync to ctxs[0]'s device, ctxs[0]'s stream waits on the other streams' events and
    finishes the top.  Slices, roots, top levels and cap against the oracle (on a one-GPU box the contexts share cuda:0)."""
    from plonky2_merkle_trees_b200 import sharded
    from plonky2_merkle_trees_b200.device import to_device, to_host
    n = 1 << lg
    rows = splitmix_felts(900 + lg + w + h + G, n * w).reshape(n, w)
    ctxs = ctx_pool[:G]
    per = n // G
    d_leaves = [to_device(rows[r * per:(r + 1) * per], "cuda:%d" % c.device) for r, c in enumerate(ctxs)]
    for rep in range(2):      # the second call reuses the events and staging
        d_dig, d_roots, d_top, d_cap = api.mt.MerkleTree.new_multi_dev(d_leaves, n, h, ctxs)
        ctxs[0].sync()
    for c in ctxs:
        c.sync()
    odg, ocap = oracle.merkle_tree_new(rows, h, threads=4, fast=True)
    assert np.array_equal(to_host(d_cap), ocap)
    g = G.bit_length() - 1
    chunks = [to_host(t) for t in d_dig]
    if h >= g:
        t = sharded.ShardedMerkleTree(n, w, h, G, 0, None, None, None, to_host(d_cap))
    else:
        t = sharded.ShardedMerkleTree(n, w, h, G, 0, None, to_host(d_roots), to_host(d_top)[:G - (1 << h)], to_host(d_cap))
    assert np.array_equal(t.assemble_global(chunks), odg)


def test_sharded_build_with_nccl_inside_the_library_world_of_one(ctx, api, oracle):
    """pmt_nccl_unique_id / pmt_comm_init / pmt_merkle_tree_build_sharded_dev with a communicator of ONE rank (the box the
    driver tests on has one GPU; tools/multigpu_check.py and bench.py --gpus N run the same call over 2 .. 8 ranks): libnccl
    is loaded with dlopen, the communicator hangs off the ctx, the build equals the plain one."""
    import ctypes as C
    from plonky2_merkle_trees_b200 import _lib
    from plonky2_merkle_trees_b200.device import dev_u64, dptr, to_device, to_host
    c = _lib.Context(0)
    try:
        uid = C.create_string_buffer(128)
        c.call("pmt_nccl_unique_id", C.cast(uid, C.c_void_p))
        assert any(uid.raw)
        c.call("pmt_comm_init", C.cast(uid, C.c_void_p), 0, 1)
        for lg, w, h in [(10, 4, 0), (9, 9, 3), (14, 4, 0)]:
            n = 1 << lg
            rows = splitmix_felts(lg + w + h, n * w).reshape(n, w)
            d_dig, d_roots, d_top, d_cap = dev_u64((2 * (n - (1 << h)), 4), "cuda:0"), dev_u64((1, 4), "cuda:0"), dev_u64((1, 4), "cuda:0"), dev_u64((1 << h, 4), "cuda:0")
            c.call("pmt_merkle_tree_build_sharded_dev", dptr(to_device(rows, "cuda:0")), n, w, h, dptr(d_dig), dptr(d_roots), dptr(d_top), dptr(d_cap))
            c.sync()
            odg, ocap = oracle.merkle_tree_new(rows, h, threads=4, fast=True)
            assert np.array_equal(to_host(d_dig), odg) and np.array_equal(to_host(d_cap), ocap)
        with pytest.raises(_lib.PmtError):
            c.call("pmt_comm_init", C.cast(uid, C.c_void_p), 0, 3)       # not a power of two
        c.call("pmt_comm_destroy")
        with pytest.raises(_lib.PmtError):                               # no communicator any more
            c.call("pmt_merkle_tree_build_sharded_dev", dptr(d_dig), 1 << 10, 4, 0, dptr(d_dig), dptr(d_roots), dptr(d_top), dptr(d_cap))
    finally:
        c.close()

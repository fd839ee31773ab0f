"""Pin the CPU oracle (C + pure Python) to every known answer the reference holds for this path (CPU only)."""
import os
import random

import numpy as np
import pytest

from conftest import EDGE, P, splitmix_felts, u64

import py_oracle as po


def test_upstream_permutation_vectors(oracle, golden):
    assert oracle.permute([0] * 12)[:4].tolist() == golden["upstream"]["perm_zero"]
    assert oracle.permute(list(range(12)))[:4].tolist() == golden["upstream"]["perm_range12"]
    assert oracle.permute([0] * 12, fast=True)[:4].tolist() == golden["upstream"]["perm_zero"]
    assert po.permute([0] * 12)[:4] == golden["upstream"]["perm_zero"]


def test_fast_partial_rounds_equal_naive(oracle):
    rnd = random.Random(7)
    for t in range(200):
        st = [rnd.choice(EDGE) if rnd.random() < 0.3 else rnd.getrandbits(64) for _ in range(12)]
        a, b = oracle.permute(st), oracle.permute(st, fast=True)
        assert a.tolist() == b.tolist()
        if t < 20:
            assert a.tolist() == po.permute(st)
        assert all(x < P for x in a.tolist())


def test_field_mul_edges(oracle):
    for a in EDGE:
        for b in EDGE:
            assert oracle.lib().pmt_oracle_mul(a, b) == (a * b) % P


# ---- reference: simple_merkle_tree.rs:131-141, :174-191, :210-211 -------------------------------------------------
@pytest.mark.parametrize("name", ["test_build_merkle_tree_4_leaves", "test_build_merkle_tree_16_leaves"])
def test_reference_simple_tree_vectors(oracle, golden, name):
    g = golden["reference"]["simple_tree"][name]
    levels, root = oracle.simple_tree_build(g["leaves"])
    want = [d for lvl in g["levels"] for d in lvl]
    assert levels.tolist() == want
    assert root.tolist() == g["root"]
    pl, pr = po.simple_tree_build(g["leaves"])
    assert [d for lvl in pl for d in lvl] == want and pr == g["root"]


def test_reference_asserted_proof(oracle, golden):
    g = golden["reference"]["simple_tree"]["test_merkle_proof_small_tree"]
    levels, root = oracle.simple_tree_build(g["leaves"])
    proof = oracle.simple_tree_proof(levels, 4, 0)
    assert proof.tolist() == g["proof"]          # simple_merkle_tree.rs:210-211
    assert oracle.simple_tree_verify(g["leaves"][0], 0, root, proof)


def test_reference_verify_roundtrip_and_negatives(oracle, golden):
    # mirrors test_verify_* at simple_merkle_tree.rs:215-309
    g = golden["reference"]["simple_tree"]["test_build_merkle_tree_16_leaves"]
    leaves = g["leaves"]
    levels, root = oracle.simple_tree_build(leaves)
    for i in range(16):
        proof = oracle.simple_tree_proof(levels, 16, i)
        assert oracle.simple_tree_verify(leaves[i], i, root, proof)
        assert po.simple_tree_verify(leaves[i], i, root.tolist(), proof.tolist())
        inb = oracle.simple_tree_in_between(levels, root, 16, i)
        assert inb[-1].tolist() == root.tolist()
    proof = oracle.simple_tree_proof(levels, 16, 3)
    assert not oracle.simple_tree_verify(leaves[4], 3, root, proof)        # wrong leaf
    assert not oracle.simple_tree_verify(leaves[3], 2, root, proof)        # wrong index
    bad = proof.copy(); bad[1, 0] ^= np.uint64(1)
    assert not oracle.simple_tree_verify(leaves[3], 3, root, bad)          # wrong proof
    assert not oracle.simple_tree_verify(leaves[3], 3, levels[0], proof)   # wrong root
    with pytest.raises(IndexError):
        oracle.simple_tree_proof(levels, 16, 16)                           # assert :56
    with pytest.raises(ValueError):
        oracle.simple_tree_build([1, 2, 3])                                # log2_strict :30
    with pytest.raises(ValueError):
        oracle.simple_tree_build([1])                                      # :38 underflow


# ---- reference: merkle_mountain_ranges.rs:280-297, :307-324 ---------------------------------------------------------
def test_reference_mmr_index_tables(oracle, golden):
    t = golden["reference"]["mmr_tables"]
    for size, bitmap in t["heights_bitmap"]:
        assert oracle.mmr_heights_bitmap(size) == (bitmap, 0)
    for normal, mmr in t["mmr_index"]:
        assert oracle.mmr_index(normal) == mmr
        assert mmr == 2 * normal - bin(normal).count("1")


# ---- MMR: sequential add_leaf restatement vs closed form (two independent oracles) ----------------------------------
@pytest.mark.parametrize("n", list(range(1, 41)) + [63, 64, 65, 70, 100])
def test_mmr_sequential_equals_closed_form(oracle, n):
    leaves = splitmix_felts(n, n)
    el = oracle.mmr_extend(None, leaves)
    assert el.shape[0] == 2 * n - bin(n).count("1")
    if n <= 40:
        assert el.tolist() == po.mmr_build(leaves.tolist())
        assert oracle.mmr_bag(el).tolist() == po.mmr_bag(el.tolist(), n)
    peaks = oracle.mmr_peaks(el)
    assert [el[p].tolist() for p in po.mmr_peak_positions(n)] == peaks.tolist()
    root = oracle.mmr_bag(el)
    for i in sorted({0, n // 2, n - 1}):
        sib, left = oracle.mmr_subtree_proof(el, oracle.mmr_index(i))
        want = po.mmr_proof(el.tolist(), n, i)
        assert [(s.tolist(), bool(b)) for s, b in zip(sib, left)] == want
        assert oracle.mmr_verify(leaves[i], root, sib, left, peaks) == 1
        assert oracle.mmr_verify(int(leaves[i]) ^ 1, root, sib, left, peaks) == -1   # assert! :245 panics
        assert oracle.mmr_verify(leaves[i], peaks[0] ^ np.uint64(1), sib, left, peaks) == 0


def test_mmr_incremental_extend(oracle):
    leaves = splitmix_felts(3, 50)
    a = oracle.mmr_extend(None, leaves)
    b = oracle.mmr_extend(oracle.mmr_extend(None, leaves[:19]), leaves[19:])
    assert a.tolist() == b.tolist()


# ---- derived (unpinned) anchors + plonky2 layout ---------------------------------------------------------------------
def test_derived_anchors(oracle, golden):
    d = golden["derived_unpinned"]
    assert oracle.two_to_one([1, 2, 3, 4], [5, 6, 7, 8]).tolist() == d["two_to_one_1234_5678"]
    for n in (5, 8, 9, 12, 16, 17, 135):
        assert oracle.hash_or_noop(list(range(n))).tolist() == d["hash_no_pad_range_%d" % n]
    assert oracle.hash_or_noop([P, P + 1, 2**64 - 1]).tolist() == d["hash_or_noop_noncanonical"] == [0, 1, 2**32 - 2, 0]
    for n, want in d["mmr_bag_leaves_1_to_n"].items():
        el = oracle.mmr_extend(None, list(range(1, int(n) + 1)))
        assert oracle.mmr_bag(el).tolist() == want
    g = d["merkle_tree_new_8x5_cap1"]
    dg, cap = oracle.merkle_tree_new(g["rows"], 1)
    assert dg.tolist() == g["digests"] and cap.tolist() == g["cap"]


@pytest.mark.parametrize("lg,w,h", [(0, 4, 0), (1, 1, 0), (1, 4, 1), (3, 4, 0), (4, 5, 2), (5, 9, 5), (6, 135, 4), (6, 3, 0)])
def test_merkle_tree_new_layout_and_proofs(oracle, lg, w, h):
    n = 1 << lg
    rows = splitmix_felts(11, n * w).reshape(n, w)
    dg, cap = oracle.merkle_tree_new(rows, h)
    dg2, cap2 = oracle.merkle_tree_new(rows, h, threads=4, fast=True)
    assert dg.tolist() == dg2.tolist() and cap.tolist() == cap2.tolist()
    pdg, pcap = po.merkle_tree_new(rows.tolist(), h)
    assert dg.tolist() == pdg and cap.tolist() == pcap
    for i in range(n):
        sib = oracle.merkle_prove(dg, n, h, i)
        assert oracle.merkle_verify_to_cap(rows[i], i, cap, h, sib)
        if sib.shape[0]:
            bad = sib.copy(); bad[0, 0] ^= np.uint64(1)
            assert not oracle.merkle_verify_to_cap(rows[i], i, cap, h, bad)
    if h == 0 and w == 1 and n >= 2:
        pass


def test_simple_tree_equals_plonky2_tree_cap0(oracle):
    # the simple tree (w = 1, all levels) and MerkleTree::new(cap_height 0) hold the same digests, different order
    n = 32
    leaves = splitmix_felts(5, n)
    levels, root = oracle.simple_tree_build(leaves)
    dg, cap = oracle.merkle_tree_new(leaves.reshape(n, 1), 0)
    assert cap[0].tolist() == root.tolist()
    off, m, l = 0, n, 0
    while m >= 2:
        for k in range(m):
            assert dg[po.digest_index(l, k)].tolist() == levels[off + k].tolist()
        off += m; m //= 2; l += 1


def test_frequency_domain_mds_is_exact_on_the_host(tmp_path):
    """The production MDS layer (csrc/poseidon_freq.cuh: fp64, frequency-domain convolution) compiled for the HOST with g++
    and checked by tests/cpp/check_freq.cpp: every layer output at all 4096 corners of the input cube against the integer
    matrix form for every constant set, and 200 000 whole permutations (random + edge states) against the oracle's
    specification-form permutation.  tools/gen_freq_constants.py re-derives the tables and proves the bounds."""
    import subprocess
    import oracle as orc
    orc.build()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "check_freq")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe, os.path.join(root, "tests", "cpp", "check_freq.cpp"),
                           "-L" + os.path.join(root, "oracle"), "-lpmt_oracle", "-Wl,-rpath," + os.path.join(root, "oracle")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


def test_frequency_domain_tables_are_reproducible(tmp_path):
    """the committed poseidon_freq_constants.cuh is what tools/gen_freq_constants.py derives (and its proofs pass)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "plonky2_merkle_trees_b200", "csrc", "poseidon_freq_constants.cuh")
    before = open(path).read()
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_freq_constants.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "exactness: worst" in out.stdout, out.stdout + out.stderr
    assert open(path).read() == before

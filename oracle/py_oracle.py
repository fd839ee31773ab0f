"""Pure-Python restatement of the hot path (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

A second, independent oracle: Python ints and ``pow(x, 7, p)`` instead of the C oracle's u128 / reduce128
arithmetic, and the MMR / plonky2 layouts written from the *closed forms* rather than from the reference's
sequential code, so that the two oracles cross-check each other (tests/test_oracle_kat.py).  Small cases only.

Citations: simple tree ``/root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:21-109``; MMR
``/root/reference/src/mmr/merkle_mountain_ranges.rs:39-270``; Poseidon / Hasher / MerkleTree::new are the
un-vendored plonky2 v0.1.3 @ 3b21b87d (see oracle/pmt_oracle.h).
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import gen_constants as _g  # noqa: E402

P = _g.P
RC = _g.generate_round_constants()
_MDS = _g.mds_matrix()


def permute(state):
    return _g.poseidon_naive(state, RC)


def two_to_one(l, r):
    return permute(list(l) + list(r) + [0, 0, 0, 0])[:4]


def hash_no_pad(xs):
    s = [0] * 12
    for off in range(0, len(xs), 8):
        chunk = xs[off:off + 8]
        s[:len(chunk)] = [x % P for x in chunk]
        s = permute(s)
    return s[:4]


def hash_or_noop(xs):
    if len(xs) <= 4:
        return [x % P for x in xs] + [0] * (4 - len(xs))
    return hash_no_pad(xs)


# ---- simple tree ----------------------------------------------------------------------------------------
def simple_tree_build(leaves):
    n = len(leaves)
    assert n >= 2 and n & (n - 1) == 0
    levels = [[hash_or_noop([x]) for x in leaves]]
    while len(levels[-1]) > 2:
        cur = levels[-1]
        levels.append([two_to_one(cur[2 * i], cur[2 * i + 1]) for i in range(len(cur) // 2)])
    root = two_to_one(levels[-1][0], levels[-1][1])
    return levels, root


def simple_tree_proof(levels, idx):
    out = []
    for lvl in levels:
        out.append(lvl[idx ^ 1])
        idx >>= 1
    return out


def simple_tree_verify(leaf, idx, root, hashes):
    cur = hash_or_noop([leaf])
    for h in hashes:
        cur = two_to_one(cur, h) if idx % 2 == 0 else two_to_one(h, cur)
        idx >>= 1
    return cur == [x % P for x in root]


# ---- MMR, closed form: node of height h covering leaves [k 2^h, (k+1) 2^h) exists iff (k+1) 2^h <= n and sits
#      at post-order position 2*last - popcount(last) + h with last = (k+1) 2^h - 1 ---------------------------
def mmr_pos(h, k):
    last = ((k + 1) << h) - 1
    return 2 * last - bin(last).count("1") + h


def mmr_size(n):
    return 2 * n - bin(n).count("1")


def mmr_build(leaves):
    n = len(leaves)
    elements = [None] * mmr_size(n)
    level = [hash_or_noop([x]) for x in leaves]
    h = 0
    while level:
        for k, d in enumerate(level):
            elements[mmr_pos(h, k)] = d
        level = [two_to_one(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
        h += 1
    assert all(e is not None for e in elements)
    return elements


def mmr_peak_positions(n):
    out, base = [], 0
    for b in range(n.bit_length() - 1, -1, -1):
        if (n >> b) & 1:
            k = base >> b
            out.append(mmr_pos(b, k))
            base += 1 << b
    return out


def mmr_bag(elements, n):
    flat = [x for p in mmr_peak_positions(n) for x in elements[p]]
    return hash_or_noop(flat)


def mmr_proof(elements, n, leaf):
    """-> [(sibling digest, sibling_on_left)] for leaf `leaf` (normal index)."""
    # mountain containing the leaf
    base = 0
    for b in range(n.bit_length() - 1, -1, -1):
        if (n >> b) & 1:
            if leaf < base + (1 << b):
                height = b
                break
            base += 1 << b
    out = []
    k = leaf
    for h in range(height):
        out.append((elements[mmr_pos(h, k ^ 1)], bool(k & 1)))
        k >>= 1
    return out


# ---- plonky2 MerkleTree::new layout, closed form -------------------------------------------------------------
def digest_index(level, k):
    """index inside one cap-subtree's digest slice of the node at `level` (0 = leaf digests), position k."""
    return 2 * (((k >> 1) << (level + 1)) + (1 << level) - 1) + (k & 1)


def merkle_tree_new(rows, cap_height):
    n = len(rows)
    lg = n.bit_length() - 1
    assert n == 1 << lg and cap_height <= lg
    L = lg - cap_height
    sub = (1 << (L + 1)) - 2
    digests = [None] * (2 * (n - (1 << cap_height)))
    level = [hash_or_noop(r) for r in rows]
    for l in range(L):
        per = 1 << (L - l)  # nodes per subtree at this level
        for g, d in enumerate(level):
            c, k = divmod(g, per)
            digests[c * sub + digest_index(l, k)] = d
        level = [two_to_one(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
    assert all(d is not None for d in digests)
    return digests, level

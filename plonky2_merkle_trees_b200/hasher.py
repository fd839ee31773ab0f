"""PoseidonHash (plonky2 `Hasher`) batch entry points over host numpy buffers -- thin wrappers of the C ABI.

Mirrors the three calls the reference makes: `hash_or_noop`, `two_to_one`, `hash_no_pad`
(/root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:23,33; src/mmr/merkle_mountain_ranges.rs:96,111,125).
"""
import numpy as np

from . import _lib
from ._lib import as_u64, ptr


def permute(states, ctx=None):
    ctx = ctx or _lib.default_context()
    s = as_u64(states, (-1, 12))
    out = np.empty_like(s)
    ctx.call("pmt_permute", ptr(s), s.shape[0], ptr(out))
    return out


def two_to_one(left, right, ctx=None):
    ctx = ctx or _lib.default_context()
    l, r = as_u64(left, (-1, 4)), as_u64(right, (-1, 4))
    if l.shape != r.shape:
        raise ValueError("two_to_one: shape mismatch")
    out = np.empty_like(l)
    ctx.call("pmt_hash_two_to_one", ptr(l), ptr(r), l.shape[0], ptr(out))
    return out


def hash_or_noop(rows, ctx=None):
    ctx = ctx or _lib.default_context()
    rows = as_u64(rows)
    if rows.ndim == 1:
        rows = rows.reshape(1, -1)
    out = np.empty((rows.shape[0], 4), np.uint64)
    ctx.call("pmt_hash_or_noop", ptr(rows), rows.shape[0], rows.shape[1], ptr(out))
    return out


def hash_no_pad(rows, ctx=None):
    ctx = ctx or _lib.default_context()
    rows = as_u64(rows)
    if rows.ndim == 1:
        rows = rows.reshape(1, -1)
    out = np.empty((rows.shape[0], 4), np.uint64)
    ctx.call("pmt_hash_no_pad", ptr(rows), rows.shape[0], rows.shape[1], ptr(out))
    return out

"""ctypes front-end of the C oracle, oracle/libpmt_oracle.so (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
All arrays are numpy uint64, digests are rows of 4.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpmt_oracle.so")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("pmt_oracle.c", "pmt_oracle.h", "poseidon_constants.h", "Makefile")]
    stale = lambda: not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale():
        import fcntl
        with open(_SO + ".lock", "w") as lock:      # several ranks may get here at once: one runs make, the others wait
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or stale():
                    subprocess.check_call(["make", "-C", _HERE, "-s"])
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


_lib = None
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        L = _lib
        L.pmt_oracle_canonical.restype = C.c_uint64
        L.pmt_oracle_canonical.argtypes = [C.c_uint64]
        L.pmt_oracle_mul.restype = C.c_uint64
        L.pmt_oracle_mul.argtypes = [C.c_uint64, C.c_uint64]
        L.pmt_oracle_mmr_index.restype = C.c_size_t
        L.pmt_oracle_mmr_index.argtypes = [C.c_size_t]
        L.pmt_oracle_mmr_peaks.restype = C.c_size_t
        L.pmt_oracle_mmr_subtree_proof.restype = C.c_size_t
    return _lib


def _p(a):
    return a.ctypes.data_as(_u64p)


def _arr(x, shape=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64))
    return a if shape is None else a.reshape(shape)


def permute(state, fast=False):
    s = _arr(state).copy()
    assert s.size == 12
    (lib().pmt_oracle_permute_fast if fast else lib().pmt_oracle_permute)(_p(s))
    return s


def two_to_one(l, r):
    l, r = _arr(l), _arr(r)
    out = np.zeros(4, np.uint64)
    lib().pmt_oracle_two_to_one(_p(l), _p(r), _p(out))
    return out


def two_to_one_batch(l, r):
    l, r = _arr(l, (-1, 4)), _arr(r, (-1, 4))
    out = np.zeros_like(l)
    for i in range(l.shape[0]):
        lib().pmt_oracle_two_to_one(_p(l[i]), _p(r[i]), _p(out[i]))
    return out


def hash_no_pad(x):
    x = _arr(x)
    out = np.zeros(4, np.uint64)
    lib().pmt_oracle_hash_no_pad(_p(x), C.c_size_t(x.size), _p(out))
    return out


def hash_or_noop(x):
    x = _arr(x)
    out = np.zeros(4, np.uint64)
    lib().pmt_oracle_hash_or_noop(_p(x), C.c_size_t(x.size), _p(out))
    return out


def hash_or_noop_rows(rows):
    rows = _arr(rows)
    out = np.zeros((rows.shape[0], 4), np.uint64)
    for i in range(rows.shape[0]):
        lib().pmt_oracle_hash_or_noop(_p(rows[i]), C.c_size_t(rows.shape[1]), _p(out[i]))
    return out


# ---- simple tree ----------------------------------------------------------------------------------------
def simple_tree_build(leaves):
    leaves = _arr(leaves)
    n = leaves.size
    levels = np.zeros((max(2 * n - 2, 0), 4), np.uint64)
    root = np.zeros(4, np.uint64)
    rc = lib().pmt_oracle_simple_tree_build(_p(leaves), C.c_size_t(n), _p(levels), _p(root))
    if rc != 0:
        raise ValueError("simple_tree_build: n must be a power of two >= 2")
    return levels, root


def simple_tree_proof(levels, n, idx):
    lg = n.bit_length() - 1
    out = np.zeros((lg, 4), np.uint64)
    if lib().pmt_oracle_simple_tree_proof(_p(_arr(levels)), C.c_size_t(n), C.c_size_t(idx), _p(out)) != 0:
        raise IndexError("leaf_index out of range")
    return out


def simple_tree_in_between(levels, root, n, idx):
    lg = n.bit_length() - 1
    out = np.zeros((lg, 4), np.uint64)
    if lib().pmt_oracle_simple_tree_in_between(_p(_arr(levels)), _p(_arr(root)), C.c_size_t(n), C.c_size_t(idx),
                                               _p(out)) != 0:
        raise IndexError("leaf_index out of range")
    return out


def simple_tree_verify(leaf, idx, root, hashes):
    hashes = _arr(hashes, (-1, 4))
    return bool(lib().pmt_oracle_simple_tree_verify(C.c_uint64(int(leaf)), C.c_size_t(idx), _p(_arr(root)),
                                                    _p(hashes), C.c_size_t(hashes.shape[0])))


# ---- MMR ---------------------------------------------------------------------------------------------------
def mmr_heights_bitmap(size):
    pk, rem = C.c_uint64(0), C.c_size_t(0)
    lib().pmt_oracle_mmr_heights_bitmap(C.c_size_t(size), C.byref(pk), C.byref(rem))
    return pk.value, rem.value


def mmr_index(i):
    return lib().pmt_oracle_mmr_index(C.c_size_t(i))


def mmr_extend(elements, leaves):
    """sequential add_leaf loop (merkle_mountain_ranges.rs:89-120). elements: (len, 4) or None."""
    leaves = _arr(leaves)
    old = 0 if elements is None else elements.shape[0]
    # new elements = 2 m + popcount(n_before) - popcount(n_before + m) <= 2 m + 63 (an append can complete up to
    # popcount(n_before) old mountains on top of its own nodes)
    buf = np.zeros((old + 2 * leaves.size + 64, 4), np.uint64)
    if old:
        buf[:old] = elements
    ln = C.c_size_t(old)
    add = lib().pmt_oracle_mmr_add_leaf
    for x in leaves.tolist():
        add(_p(buf), C.byref(ln), C.c_uint64(x))
    return buf[:ln.value].copy()


def mmr_peaks(elements):
    elements = _arr(elements, (-1, 4))
    out = np.zeros((64, 4), np.uint64)
    k = lib().pmt_oracle_mmr_peaks(_p(elements), C.c_size_t(elements.shape[0]), _p(out))
    return out[:k].copy()


def mmr_bag(elements):
    elements = _arr(elements, (-1, 4))
    out = np.zeros(4, np.uint64)
    lib().pmt_oracle_mmr_bag(_p(elements), C.c_size_t(elements.shape[0]), _p(out))
    return out


def mmr_subtree_proof(elements, mmr_idx):
    elements = _arr(elements, (-1, 4))
    sib = np.zeros((64, 4), np.uint64)
    left = np.zeros(64, np.uint8)
    k = lib().pmt_oracle_mmr_subtree_proof(_p(elements), C.c_size_t(elements.shape[0]), C.c_size_t(mmr_idx), _p(sib),
                                           left.ctypes.data_as(_u8p))
    return sib[:k].copy(), left[:k].copy()


def mmr_verify(leaf, root, siblings, on_left, peaks):
    siblings = _arr(siblings, (-1, 4))
    on_left = np.ascontiguousarray(np.asarray(on_left, dtype=np.uint8))
    peaks = _arr(peaks, (-1, 4))
    return lib().pmt_oracle_mmr_verify(C.c_uint64(int(leaf)), _p(_arr(root)), _p(siblings),
                                       on_left.ctypes.data_as(_u8p), C.c_size_t(siblings.shape[0]), _p(peaks),
                                       C.c_size_t(peaks.shape[0]))


# ---- plonky2 MerkleTree::new ------------------------------------------------------------------------------
def merkle_tree_new(rows, cap_height, threads=1, fast=False):
    rows = _arr(rows)
    n, w = rows.shape
    ncap = 1 << cap_height
    digests = np.zeros((2 * (n - ncap), 4), np.uint64)
    cap = np.zeros((ncap, 4), np.uint64)
    rc = lib().pmt_oracle_merkle_tree_new(_p(rows), C.c_size_t(n), C.c_size_t(w), C.c_uint(cap_height), _p(digests),
                                          _p(cap), C.c_int(threads), C.c_int(int(fast)))
    if rc != 0:
        raise ValueError("merkle_tree_new: n must be a power of two and cap_height <= log2 n")
    return digests, cap


def merkle_prove(digests, n, cap_height, idx):
    lg = n.bit_length() - 1
    out = np.zeros((lg - cap_height, 4), np.uint64)
    if lib().pmt_oracle_merkle_prove(_p(_arr(digests)), C.c_size_t(n), C.c_uint(cap_height), C.c_size_t(idx),
                                     _p(out)) != 0:
        raise IndexError("bad index")
    return out


def merkle_verify_to_cap(leaf_row, idx, cap, cap_height, siblings):
    leaf_row = _arr(leaf_row)
    siblings = _arr(siblings, (-1, 4))
    return bool(lib().pmt_oracle_merkle_verify_to_cap(_p(leaf_row), C.c_size_t(leaf_row.size), C.c_size_t(idx),
                                                      _p(_arr(cap)), C.c_uint(cap_height), _p(siblings),
                                                      C.c_size_t(siblings.shape[0])))


def max_threads():
    """host cores this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, which would
    silently time the CPU baseline on one thread; the thread count is passed to the library explicitly instead."""
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or lib().pmt_oracle_max_threads())

#!/usr/bin/env python3
"""Tuning aid: end-to-end time of the host-buffer pmt_merkle_tree_build (pinned leaves up, every digest down) for several
pipeline depths, next to the raw PCIe copy times of the same buffers.  One JSON line per measurement."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plonky2_merkle_trees_b200 import _lib  # noqa: E402
from plonky2_merkle_trees_b200._lib import u64p  # noqa: E402


def host_memory_kinds():
    """pmt_merkle_tree_build on the headline workload from four kinds of host memory: torch-pinned, PAGEABLE (numpy = what a
    plain Rust Vec is), the same pageable buffers page-locked with pmt_host_register, and pmt_host_alloc'ed buffers."""
    lg, w = 24, 4
    n = 1 << lg
    ctx = _lib.default_context(0)
    dev = torch.device("cuda", 0)
    src = bench.splitmix_torch(0, n * w, dev).view(n, w).cpu().numpy().view(np.uint64)
    nd = 2 * (n - 1)

    def run(name, p_leaves, p_dig, p_cap, check, extra=None):
        ts = []
        for _ in range(5):
            ctx.sync(); t0 = time.perf_counter()
            ctx.call("pmt_merkle_tree_build", C.cast(p_leaves, u64p), n, w, 0, C.cast(p_dig, u64p), C.cast(p_cap, u64p))
            ts.append(time.perf_counter() - t0)
        out = {"e2e_host_memory": name, "ms_best": 1e3 * min(ts[1:]), "ms_median": 1e3 * sorted(ts[1:])[2], "ms_first_call": 1e3 * ts[0],
               "M_leaves_per_s": n / min(ts[1:]) / 1e6, "checksum": "%016x" % int(np.bitwise_xor.reduce(check().reshape(-1)[::997]))}
        out.update(extra or {})
        print(json.dumps(out), flush=True)

    h_l = torch.empty((n, w), dtype=torch.int64).pin_memory(); h_l.copy_(torch.from_numpy(src.view(np.int64)))
    h_d = torch.empty((nd, 4), dtype=torch.int64).pin_memory(); h_c = torch.empty((1, 4), dtype=torch.int64).pin_memory()
    run("pinned (torch pin_memory)", h_l.data_ptr(), h_d.data_ptr(), h_c.data_ptr(), lambda: h_d.numpy().view(np.uint64))
    del h_l, h_d
    pl, pd, pc = src.copy(), np.zeros((nd, 4), np.uint64), np.zeros((1, 4), np.uint64)
    run("pageable (numpy / Vec)", pl.ctypes.data, pd.ctypes.data, pc.ctypes.data, lambda: pd)
    t0 = time.perf_counter()
    ctx.call("pmt_host_register", C.c_void_p(pl.ctypes.data), pl.nbytes)
    ctx.call("pmt_host_register", C.c_void_p(pd.ctypes.data), pd.nbytes)
    reg_ms = 1e3 * (time.perf_counter() - t0)
    pd[:] = 0
    run("pageable + pmt_host_register", pl.ctypes.data, pd.ctypes.data, pc.ctypes.data, lambda: pd, {"register_ms_once": reg_ms})
    ctx.call("pmt_host_unregister", C.c_void_p(pl.ctypes.data)); ctx.call("pmt_host_unregister", C.c_void_p(pd.ctypes.data))
    del pl, pd
    a, b = C.c_void_p(), C.c_void_p()
    t0 = time.perf_counter()
    ctx.call("pmt_host_alloc", n * w * 8, C.byref(a)); ctx.call("pmt_host_alloc", nd * 32, C.byref(b))
    alloc_ms = 1e3 * (time.perf_counter() - t0)
    al = np.ctypeslib.as_array(C.cast(a, u64p), shape=(n, w)); ad = np.ctypeslib.as_array(C.cast(b, u64p), shape=(nd, 4))
    al[:] = src
    run("pmt_host_alloc", a.value, b.value, pc.ctypes.data, lambda: ad, {"alloc_ms_once": alloc_ms})
    ctx.call("pmt_host_free", a); ctx.call("pmt_host_free", b)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "mem":
        return host_memory_kinds()
    lg, w = 24, 4
    n = 1 << lg
    ctx = _lib.default_context(0)
    dev = torch.device("cuda", 0)
    h_leaves = torch.empty((n, w), dtype=torch.int64).pin_memory()
    h_leaves.copy_(bench.splitmix_torch(0, n * w, dev).view(n, w))
    h_dig = torch.empty((2 * (n - 1), 4), dtype=torch.int64).pin_memory()
    h_cap = torch.empty((1, 4), dtype=torch.int64).pin_memory()
    d_a = torch.empty_like(h_leaves, device=dev)
    d_b = torch.empty_like(h_dig, device=dev)
    for name, fn in [("h2d_512MiB", lambda: d_a.copy_(h_leaves, non_blocking=True)), ("d2h_1GiB", lambda: h_dig.copy_(d_b, non_blocking=True))]:
        ts = []
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(json.dumps({"copy": name, "ms": 1e3 * min(ts)}), flush=True)
    del d_a, d_b
    torch.cuda.empty_cache()
    # reference digests: the device-resident build (tied to the oracle by tests/test_gpu_parity.py)
    from plonky2_merkle_trees_b200 import merkle_tree
    t = merkle_tree.MerkleTree.new_dev(h_leaves.to(dev), 0, ctx)
    want = t.digests.copy()
    want_cap = t.cap.copy()
    del t
    torch.cuda.empty_cache()
    ref = None
    for chunks in (3, 4, 5, 6):
        os.environ["PMT_PIPELINE_LOG2_CHUNKS"] = str(chunks)
        ts = []
        for _ in range(5):
            ctx.sync(); t0 = time.perf_counter()
            ctx.call("pmt_merkle_tree_build", C.cast(h_leaves.data_ptr(), u64p), n, w, 0, C.cast(h_dig.data_ptr(), u64p), C.cast(h_cap.data_ptr(), u64p))
            ts.append(time.perf_counter() - t0)
        cs = int(np.bitwise_xor.reduce(h_dig.numpy().view(np.uint64).reshape(-1)[::997]))
        ref = cs if ref is None else ref
        print(json.dumps({"e2e_log2_chunks": chunks, "ms_best": 1e3 * min(ts[1:]), "ms_median": 1e3 * sorted(ts[1:])[2], "M_leaves_per_s": n / min(ts[1:]) / 1e6,
                          "same_digests": cs == ref,
                          "equals_device_build": bool(np.array_equal(h_dig.numpy().view(np.uint64), want) and
                                                      np.array_equal(h_cap.numpy().view(np.uint64), want_cap))}), flush=True)


if __name__ == "__main__":
    main()

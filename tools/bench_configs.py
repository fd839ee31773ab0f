#!/usr/bin/env python3
"""Time every BASELINE.json config on ONE GPU (device-resident inputs, CUDA events on the ctx stream) and print one JSON
line per config.  bench.py is the headline; this is the parity-sized side table."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plonky2_merkle_trees_b200 import _lib, merkle_tree, mmr, simple_merkle_tree  # noqa: E402
from plonky2_merkle_trees_b200.device import dev_u64, dptr  # noqa: E402


def timed(ctx, fn, reps=5, warm=2):
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for _ in range(warm):
        fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); ctx.sync()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    ctx = _lib.default_context(0)
    dev = "cuda:0"
    only = sys.argv[1:] or ["C1", "C2", "C3", "C3b", "C4", "C5"]
    out = []

    def emit(name, desc, units, unit_name, best, med, perms):
        line = {"config": name, "desc": desc, "ms_best": best, "ms_median": med, unit_name + "_per_s": units / (med * 1e-3),
                "permutations": perms, "Gperm_per_s": perms / (med * 1e-3) / 1e9}
        print(json.dumps(line)); out.append(line)

    if "C1" in only:
        n = 1 << 10
        d_leaves = bench.splitmix_torch(0, n, dev)
        d_levels, d_root = dev_u64((2 * n - 2, 4), dev), dev_u64((4,), dev)
        b, m = timed(ctx, lambda: ctx.call("pmt_simple_tree_build_dev", dptr(d_leaves), n, dptr(d_levels), dptr(d_root)))
        emit("C1", "simple tree 2^10 x 1 felt (build)", n, "leaves", b, m, n - 1)
    if "C2" in only:
        n, w = 1 << 20, 4
        d_leaves = bench.splitmix_torch(0, n * w, dev).view(n, w)
        d_dig, d_cap = dev_u64((2 * n - 2, 4), dev), dev_u64((1, 4), dev)
        b, m = timed(ctx, lambda: ctx.call("pmt_merkle_tree_build_dev", dptr(d_leaves), n, w, 0, dptr(d_dig), dptr(d_cap)))
        emit("C2", "plonky2 tree 2^20 x 4 felts, cap 0", n, "leaves", b, m, n - 1)
    for tag, n in (("C3", 1 << 24), ("C3b", (1 << 24) - 1)):
        if tag not in only:
            continue
        d_leaves = bench.splitmix_torch(0, n, dev)
        size = 2 * n - bin(n).count("1")
        d_el = dev_u64((size, 4), dev)
        b, m = timed(ctx, lambda: ctx.call("pmt_mmr_extend_dev", dptr(d_el), 0, dptr(d_leaves), n), reps=3, warm=1)
        emit(tag, "MMR extend by %d single-felt leaves from empty" % n, n, "leaves", b, m, n - bin(n).count("1"))
        if tag == "C3":   # end to end through host buffers: pinned leaves up, every element down (pipelined)
            import time
            import ctypes as C
            from plonky2_merkle_trees_b200._lib import u64p
            h_leaves = torch.empty(n, dtype=torch.int64).pin_memory(); h_leaves.copy_(d_leaves)
            h_el = torch.empty((size, 4), dtype=torch.int64).pin_memory()
            ts = []
            for _ in range(4):
                ctx.sync(); t0 = time.perf_counter()
                ctx.call("pmt_mmr_extend", C.cast(h_el.data_ptr(), u64p), 0, C.cast(h_leaves.data_ptr(), u64p), n)
                ts.append(1e3 * (time.perf_counter() - t0))
            ok = bool(torch.equal(h_el, d_el.cpu()))
            emit("C3-e2e", "MMR extend through HOST buffers (128 MiB up, 1 GiB down), equals device path: %s" % ok, n, "leaves",
                 min(ts[1:]), sorted(ts[1:])[1], n - bin(n).count("1"))
            del h_leaves, h_el
        d_root = dev_u64((4,), dev)
        b, m = timed(ctx, lambda: ctx.call("pmt_mmr_bag_dev", dptr(d_el), n, dptr(d_root)))
        emit(tag + "-bag", "peaks + bag (%d peaks)" % bin(n).count("1"), 1, "bags", b, m, (4 * bin(n).count("1") + 7) // 8 if bin(n).count("1") > 1 else 0)
        q = 1024
        idx = (bench.splitmix_numpy(7, q) % np.uint64(n)).astype(np.uint64)
        d_idx = torch.from_numpy(idx.view(np.int64)).to(dev)
        d_sib = torch.zeros((q, 32, 4), dtype=torch.int64, device=dev)
        d_left = torch.zeros((q, 32), dtype=torch.uint8, device=dev)
        d_len = torch.zeros(q, dtype=torch.int32, device=dev)
        b, m = timed(ctx, lambda: ctx.call("pmt_mmr_prove_dev", dptr(d_el), n, dptr(d_idx), q, dptr(d_sib), dptr(d_left), dptr(d_len)))
        emit(tag + "-prove", "1024 leaf proofs (gather)", q, "proofs", b, m, 0)
        d_peaks = dev_u64((64, 4), dev)
        import ctypes as C
        k = C.c_uint32(0)
        ctx.call("pmt_mmr_peaks_dev", dptr(d_el), n, dptr(d_peaks), C.byref(k)); ctx.sync()
        d_status = torch.empty(q, dtype=torch.int8, device=dev)
        d_lv = d_leaves[torch.from_numpy(idx.astype(np.int64)).to(dev)].contiguous()
        b, m = timed(ctx, lambda: ctx.call("pmt_mmr_verify_dev", dptr(d_lv), q, dptr(d_sib), dptr(d_left), dptr(d_len), dptr(d_peaks), k.value, dptr(d_root), dptr(d_status)))
        assert bool((d_status == 1).all()), "proofs must verify"
        emit(tag + "-verify", "verify the 1024 proofs", q, "proofs", b, m, int(d_len.sum().item()))
        del d_el
    if "C4" in only:
        n, w, h = 1 << 20, 135, 4
        d_leaves = bench.splitmix_torch(0, n * w, dev).view(n, w)
        d_dig, d_cap = dev_u64((2 * (n - 16), 4), dev), dev_u64((16, 4), dev)
        ctx.profile(True)
        b, m = timed(ctx, lambda: ctx.call("pmt_merkle_tree_build_dev", dptr(d_leaves), n, w, h, dptr(d_dig), dptr(d_cap)), reps=3, warm=1)
        prof = ctx.profile_read(); ctx.profile(False)
        emit("C4", "FRI-style 2^20 rows x 135 felts, cap 4 (1 GPU)", n, "leaves", b, m, 17 * n + n - 16)
        print(json.dumps({"config": "C4-kernels", "kernels": prof}))
        # the same commitment fed by the prover's column-major LDE values (transpose + reverse_index_bits fused into the
        # leaf kernel) against transposing first with torch and then building
        d_cols = d_leaves.t().contiguous()          # (w, n)
        b2, m2 = timed(ctx, lambda: ctx.call("pmt_merkle_tree_build_from_columns_dev", dptr(d_cols), n, w, 1, h, None, dptr(d_dig), dptr(d_cap)),
                       reps=3, warm=1)
        emit("C4-columns", "same tree from column-major LDE values, fused transpose + bit reversal", n, "leaves", b2, m2, 17 * n + n - 16)
        rev = torch.tensor([int(format(i, "020b")[::-1], 2) for i in range(n)], device=dev)

        def transpose_then_build():
            rows = d_cols.t().contiguous()[rev]     # what a prover without the fused feed does: 2 x 1 GiB round trips
            ctx.call("pmt_merkle_tree_build_dev", dptr(rows), n, w, h, dptr(d_dig), dptr(d_cap))
        b3, m3 = timed(ctx, transpose_then_build, reps=3, warm=1)
        emit("C4-transpose-first", "torch transpose + index_select, then pmt_merkle_tree_build_dev", n, "leaves", b3, m3, 17 * n + n - 16)
        del d_leaves, d_dig, d_cols
    if "C5" in only:
        n, w = 1 << 28, 4
        torch.cuda.empty_cache()
        d_leaves = bench.splitmix_torch(0, n * w, dev).view(n, w)
        d_dig, d_cap = dev_u64((2 * n - 2, 4), dev), dev_u64((1, 4), dev)
        b, m = timed(ctx, lambda: ctx.call("pmt_merkle_tree_build_dev", dptr(d_leaves), n, w, 0, dptr(d_dig), dptr(d_cap)), reps=2, warm=1)
        emit("C5", "2^28 x 4 felts, cap 0, ONE GPU (24 GiB resident)", n, "leaves", b, m, n - 1)
    return 0


if __name__ == "__main__":
    sys.exit(main())

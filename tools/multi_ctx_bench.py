#!/usr/bin/env python3
"""Tuning aid: end-to-end time of the single-process multi-GPU build (pmt_merkle_tree_build_multi: pinned leaves up, every
digest down, one ctx + one host thread per device) for 1, 2, 4 ... visible devices, and its equality with the one-device
build.  One JSON line per measurement.  usage: multi_ctx_bench.py [log2_leaves=24] [width=4] [cap_height=0]
       multi_ctx_bench.py mmr [log2_leaves=24] [ragged=0|1]   -- pmt_mmr_extend_multi from empty (2^k or 2^k - 1 single-felt leaves)"""
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


def main_mmr():
    lg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    m = (1 << lg) - (1 if len(sys.argv) > 3 and sys.argv[3] == "1" else 0)
    ndev = torch.cuda.device_count()
    ctxs = [_lib.Context(i) for i in range(ndev)]
    size = 2 * m - bin(m).count("1")
    h_leaves = torch.empty(m, dtype=torch.int64).pin_memory()
    h_leaves.copy_(bench.splitmix_torch(0, m, torch.device("cuda", 0)))
    h_el = torch.empty((size, 4), dtype=torch.int64).pin_memory()
    lib = ctxs[0].lib
    ref = None
    g = 1
    while g <= ndev:
        handles = (C.c_void_p * g)(*[c.h for c in ctxs[:g]])
        ts = []
        h_el.zero_()      # before the first (untimed) call only, see main()
        for _ in range(5):
            t0 = time.perf_counter()
            rc = lib.pmt_mmr_extend_multi(handles, g, C.cast(h_el.data_ptr(), u64p), 0, C.cast(h_leaves.data_ptr(), u64p), m)
            ts.append(time.perf_counter() - t0)
            ctxs[0].check(rc)
        el = h_el.numpy().view(np.uint64)
        cs = [int(np.bitwise_xor.reduce(el.reshape(-1)[k::997])) for k in range(3)] + el[-1].tolist()
        ref = cs if ref is None else ref
        print(json.dumps({"mmr_multi_ctx_devices": g, "leaves": m, "ms_best": 1e3 * min(ts[1:]), "ms_median": 1e3 * sorted(ts[1:])[2],
                          "M_leaves_per_s": m / min(ts[1:]) / 1e6, "equals_one_device_append": cs == ref,
                          "launches": [c.launches for c in ctxs[:g]]}), flush=True)
        g *= 2


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "mmr":
        return main_mmr()
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    n, ncap = 1 << lg, 1 << h
    ndev = torch.cuda.device_count()
    ctxs = [_lib.Context(i) for i in range(ndev)]
    h_leaves = torch.empty((n, w), dtype=torch.int64).pin_memory()
    h_leaves.copy_(bench.splitmix_torch(0, n * w, torch.device("cuda", 0)).view(n, w))
    h_dig = torch.empty((2 * (n - ncap), 4), dtype=torch.int64).pin_memory()
    h_cap = torch.empty((ncap, 4), dtype=torch.int64).pin_memory()
    lib = ctxs[0].lib
    ref = None
    g = 1
    while g <= ndev:
        handles = (C.c_void_p * g)(*[c.h for c in ctxs[:g]])
        ts = []
        h_dig.zero_()     # before the first (untimed) call only.  profiles/multi_ctx_2gpu_r1.jsonl zeroed before EVERY call: 27.4 ms
        for _ in range(5):  # on one device against 21.7 ms without (profiles/multi_ctx_1gpu_nozero_r1.jsonl): dirty host cache lines
            t0 = time.perf_counter()
            rc = lib.pmt_merkle_tree_build_multi(handles, g, C.cast(h_leaves.data_ptr(), u64p), n, w, h, C.cast(h_dig.data_ptr(), u64p),
                                                 C.cast(h_cap.data_ptr(), u64p))
            ts.append(time.perf_counter() - t0)
            ctxs[0].check(rc)
        dig = h_dig.numpy().view(np.uint64)
        cs = [int(np.bitwise_xor.reduce(dig.reshape(-1)[k::997])) for k in range(3)] + h_cap.numpy().view(np.uint64).reshape(-1)[:4].tolist()
        ref = cs if ref is None else ref
        print(json.dumps({"multi_ctx_devices": g, "log2_leaves": lg, "width": w, "cap_height": h, "ms_best": 1e3 * min(ts[1:]),
                          "ms_median": 1e3 * sorted(ts[1:])[2], "M_leaves_per_s": n / min(ts[1:]) / 1e6,
                          "equals_one_device_build": cs == ref, "launches": [c.launches for c in ctxs[:g]]}), flush=True)
        g *= 2


if __name__ == "__main__":
    main()

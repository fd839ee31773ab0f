#!/usr/bin/env python3
"""Strong scaling of the ONE-PROCESS multi-GPU builds a compiled host can call (no torch.distributed, no NCCL):

  pmt_merkle_tree_build_multi_dev   device resident: every ctx builds its subtree on its own device, 32-byte peer copies of the
                                    roots over NVLink to device 0, top levels there.  2^24 and 2^28 leaves x 4 felts.
  pmt_merkle_tree_build_multi       host buffers (pinned): the same partition, every ctx runs the pipelined H2D / hash / D2H
                                    build from its own host thread.  2^24 leaves x 4 felts, and C4 (2^20 x 135, cap 4).
  pmt_mmr_extend_multi              host buffers: 2^24 single-felt leaves from empty.

for 1, 2, 4, ... visible devices; wall clock around the call + pmt_sync (what the host sees), best / median of 7 after 2
warm-ups; the result of every device count must equal the one-device result.  One JSON line per measurement.
usage: multi_dev_bench.py [dev|host|mmr ...]   (default: all three)"""
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
from plonky2_merkle_trees_b200 import _lib, merkle_tree  # noqa: E402
from plonky2_merkle_trees_b200._lib import u64p  # noqa: E402


def timeit(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e3 * ts[0], 1e3 * ts[len(ts) // 2]


def dev_mode(ctxs):
    ndev = len(ctxs)
    for lg in (24, 28):
        n, w = 1 << lg, 4
        ref, t1 = None, None
        g = 1
        while g <= ndev:
            per = n // g
            leaves = [bench.splitmix_torch(r * per * w, per * w, torch.device("cuda", r)).view(per, w) for r in range(g)]
            for r in range(g):
                torch.cuda.synchronize(r)
            out = {}

            def run():
                out["r"] = merkle_tree.MerkleTree.new_multi_dev(leaves, n, 0, ctxs[:g])
                for c in ctxs[:g]:
                    c.sync()

            best, med = timeit(run)
            cap = out["r"][3].cpu().numpy().view(np.uint64).reshape(-1).tolist()
            ref = cap if ref is None else ref
            t1 = med if t1 is None else t1
            print(json.dumps({"multi_dev_devices": g, "log2_leaves": lg, "ms_best": best, "ms_median": med, "G_leaves_per_s": n / med / 1e6,
                              "speedup_vs_1": t1 / med, "root_equals_one_device": cap == ref}), flush=True)
            del leaves, out
            for r in range(g):
                with torch.cuda.device(r):
                    torch.cuda.empty_cache()
            g *= 2


def host_mode(ctxs):
    ndev = len(ctxs)
    lib = ctxs[0].lib
    for lg, w, h in ((24, 4, 0), (20, 135, 4)):
        n, ncap = 1 << lg, 1 << h
        h_leaves = torch.empty((n, w), dtype=torch.int64).pin_memory()
        h_leaves.copy_(bench.splitmix_torch(0, n * w, torch.device("cuda", 0)).view(n, w))
        h_dig = torch.empty((2 * (n - ncap), 4), dtype=torch.int64).pin_memory()
        h_cap = torch.empty((ncap, 4), dtype=torch.int64).pin_memory()
        h_dig.zero_()
        ref, t1, g = None, None, 1
        while g <= ndev:
            handles = (C.c_void_p * g)(*[c.h for c in ctxs[:g]])

            def run():
                rc = lib.pmt_merkle_tree_build_multi(handles, g, C.cast(h_leaves.data_ptr(), u64p), n, w, h, C.cast(h_dig.data_ptr(), u64p),
                                                     C.cast(h_cap.data_ptr(), u64p))
                ctxs[0].check(rc)

            best, med = timeit(run, reps=5, warm=2)
            dig = h_dig.numpy().view(np.uint64)
            cs = [int(np.bitwise_xor.reduce(dig.reshape(-1)[k::997])) for k in range(3)] + h_cap.numpy().view(np.uint64).reshape(-1)[:4].tolist()
            ref = cs if ref is None else ref
            t1 = med if t1 is None else t1
            print(json.dumps({"multi_ctx_host_devices": g, "log2_leaves": lg, "width": w, "cap_height": h, "ms_best": best, "ms_median": med,
                              "M_leaves_per_s": n / med / 1e3, "speedup_vs_1": t1 / med, "equals_one_device_build": cs == ref}), flush=True)
            g *= 2
        del h_leaves, h_dig


def mmr_mode(ctxs):
    ndev = len(ctxs)
    lib = ctxs[0].lib
    m = 1 << 24
    size = 2 * m - 1
    h_leaves = torch.empty(m, dtype=torch.int64).pin_memory()
    h_leaves.copy_(bench.splitmix_torch(0, m, torch.device("cuda", 0)))
    h_el = torch.empty((size, 4), dtype=torch.int64).pin_memory()
    h_el.zero_()
    ref, t1, g = None, None, 1
    while g <= ndev:
        handles = (C.c_void_p * g)(*[c.h for c in ctxs[:g]])

        def run():
            rc = lib.pmt_mmr_extend_multi(handles, g, C.cast(h_el.data_ptr(), u64p), 0, C.cast(h_leaves.data_ptr(), u64p), m)
            ctxs[0].check(rc)

        best, med = timeit(run, reps=5, warm=2)
        el = h_el.numpy().view(np.uint64)
        cs = [int(np.bitwise_xor.reduce(el.reshape(-1)[k::997])) for k in range(3)] + el[-1].tolist()
        ref = cs if ref is None else ref
        t1 = med if t1 is None else t1
        print(json.dumps({"mmr_multi_ctx_host_devices": g, "leaves": m, "ms_best": best, "ms_median": med, "M_leaves_per_s": m / med / 1e3,
                          "speedup_vs_1": t1 / med, "equals_one_device_append": cs == ref}), flush=True)
        g *= 2


def main():
    ndev = torch.cuda.device_count()
    ctxs = [_lib.Context(i) for i in range(ndev)]
    modes = sys.argv[1:] or ["dev", "host", "mmr"]
    if "dev" in modes:
        dev_mode(ctxs)
    if "host" in modes:
        host_mode(ctxs)
    if "mmr" in modes:
        mmr_mode(ctxs)


if __name__ == "__main__":
    main()

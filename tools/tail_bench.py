#!/usr/bin/env python3
"""Tuning aid: where the latency-bound tail of a tree goes.

For trees of 2^lg leaves x 4 felts (cap 0, upstream layout, device-resident) prints, per PMT_COOP_MAX_LOG2 setting, the
whole build time (CUDA events on the ctx stream, median of 20) and the per-kernel times of libpmt's own profile
(pmt_profile_read: events around every launch).  The tail of a big tree = the k_tree_coop line: all levels of at most
2^PMT_COOP_MAX_LOG2 nodes in one launch.  usage: tail_bench.py [lg ...]   (default 10 14 16 18 20 21 24)
One JSON line per (lg, threshold, fused)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(lg, coop, fused):
    import torch
    import bench
    from plonky2_merkle_trees_b200 import _lib
    from plonky2_merkle_trees_b200.device import dev_u64, dptr
    dev = torch.device("cuda", 0)
    os.environ["PMT_COOP_MAX_LOG2"] = str(coop)          # read by pmt_init
    os.environ["PMT_FUSE_SUBTREES"] = "1" if fused else "0"
    ctx = _lib.Context(0)
    n, w = 1 << lg, 4
    d_leaves = bench.splitmix_torch(0, n * w, dev).view(n, w)
    d_dig, d_cap = dev_u64((2 * n - 2, 4), dev), dev_u64((1, 4), dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def build():
        ctx.call("pmt_merkle_tree_build_dev", dptr(d_leaves), n, w, 0, dptr(d_dig), dptr(d_cap))

    for _ in range(3):
        build()
    ctx.sync()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); build(); e1.record(stream); ctx.sync()
        ts.append(e0.elapsed_time(e1))
    ctx.profile(True)
    reps = 10
    for _ in range(reps):
        build()
    prof = ctx.profile_read()
    ctx.profile(False)
    ts.sort()
    print(json.dumps({"log2_leaves": lg, "coop_max_log2": coop, "fused": fused, "ms_median": ts[len(ts) // 2], "ms_best": ts[0],
                      "root0": "%016x" % (int(d_cap.cpu().view(-1)[0].item()) & (2**64 - 1)),
                      "kernels_us_per_build": {k: {"launches": v["launches"] // reps, "us": round(1e3 * v["ms"] / reps, 2)} for k, v in prof.items()}}),
          flush=True)
    ctx.close()


def main():
    lgs = [int(a) for a in sys.argv[1:]] or [10, 14, 16, 18, 20, 21, 24]
    for lg in lgs:
        for coop, fused in [(13, True), (13, False), (12, True), (14, True), (15, True), (16, True)]:
            if coop > lg - 1 and coop != 13:
                continue
            measure(lg, coop, fused)


if __name__ == "__main__":
    main()

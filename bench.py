#!/usr/bin/env python3
"""bench.py -- headline benchmark: Poseidon-Goldilocks Merkle leaves hashed/sec (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (libpmt on B200)
    python bench.py --impl reference --gpus N ...            # reference arm: the CPU path on the box's host cores

Workload (config.workload): plonky2 `MerkleTree::new` over 2^24 leaves x 4 Goldilocks felts per GPU, cap_height 0,
output in upstream's `digests` layout.  One step = one complete tree build.  With N > 1 the tree has N * 2^24 leaves,
subtree-sharded: every rank builds its 2^24-leaf subtree on its own GPU, the N roots are exchanged with one NCCL
all_gather and the log2(N) top levels are finished on every rank ("scaling": "weak": per-GPU work is fixed).

`value`      leaves/s, inputs resident in HBM, CUDA-event timing on the launching stream, max over ranks.
`e2e`        the same metric through the host-buffer C ABI call (pmt_merkle_tree_build): pinned host leaves -> H2D ->
             build -> D2H of every digest + cap, all inside the timed region.
`roofline`   dominant kernel = k_level (one two_to_one per thread).  The path is integer-pipe bound:
             achieved = permutations/s x 10 588 MAC32 / measured IMAD.WIDE.U32 issue peak; the HBM view (96 algorithmic
             bytes per permutation against MEASURED_PEAKS.json) is reported beside it as evidence that HBM is not the limit.
`cpu_baseline` the CPU oracle's restatement of MerkleTree::new (OpenMP fork-join like rayon's) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_LEAVES_PER_GPU = 24
WIDTH = 4
CAP_HEIGHT = 0
SEED = 0x706D745F62323030
P = 0xFFFFFFFF00000001
MAC32_PER_PERM = 10588            # spec-level 32x32->64 multiply-accumulates per permutation
ALG_BYTES_PER_PERM = 96           # read two 32 B children, write one 32 B parent
# measured on this pool's B200 by tools/perm_bench.cu (profiles/pipes_r1.jsonl): the multiply-accumulate issue peak of
# the SM's fma / fp64 pipes = 63.7 lanes/clk/SM (32-bit IMAD; DFMA reaches 58.3) x 148 SMs x 1965 MHz.  IMAD.WIDE.U32, the
# instruction a 32x32->64 MAC maps to literally, issues at HALF that rate (31.7 lanes/clk/SM = 9.2 T/s, with or without
# an addend); the kernel runs the MDS layers' MACs as DFMAs, which is why it can exceed the IMAD.WIDE-only figure.
IMAD_WIDE_PEAK_PER_S = 9.21e12
PEAK_MAC32_PER_S = 18.53e12


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------------------------
# synthetic leaves: SURVEY.md 8(d) stateless SplitMix64 generator, identical on CPU (numpy) and GPU (torch int64)
# ------------------------------------------------------------------------------------------------------------------
def splitmix_numpy(start, count):
    import numpy as np
    idx = np.arange(start + 1, start + count + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(SEED) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        z = np.where(z >= np.uint64(P), z - np.uint64(P), z)
    return z


def splitmix_torch(start, count, device):
    """same generator with int64 wrap-around arithmetic (bit patterns identical to the uint64 version)."""
    import torch

    def s64(x):
        x &= (1 << 64) - 1
        return x - (1 << 64) if x >> 63 else x

    def lsr(t, k):
        return (t >> k) & ((1 << (64 - k)) - 1)

    out = torch.empty(count, dtype=torch.int64, device=device)
    chunk = 1 << 24
    for off in range(0, count, chunk):
        m = min(chunk, count - off)
        idx = torch.arange(start + off + 1, start + off + m + 1, dtype=torch.int64, device=device)
        z = idx * s64(0x9E3779B97F4A7C15) + s64(SEED)
        z = (z ^ lsr(z, 30)) * s64(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * s64(0x94D049BB133111EB)
        z = z ^ lsr(z, 31)
        ge_p = (z < 0) & (z >= s64(P))           # unsigned z >= p
        out[off:off + m] = torch.where(ge_p, z + (2**32 - 1), z)   # z - p == z + 2^32 - 1 (mod 2^64)
    return out


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region.  NVML in a thread (one sample per ~5 ms: the timed region of 10
    steps is only ~140 ms, too short for `nvidia-smi -lms`); falls back to polling nvidia-smi if pynvml is unusable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.idx, self.sm, self.reasons, self.max_mhz = gpu_index, [], set(), None
        self.stop_flag, self.thread, self.how = threading.Event(), None, None

    def _nvml_loop(self, nv, h):
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def _smi_loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.sm.append(float(f[0])); self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else None
            h = nv.nvmlDeviceGetHandleByIndex(ids[self.idx] if ids else self.idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.how = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
        except Exception:
            self.how = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no clock samples"], "samples": 0, "how": self.how}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "how": self.how}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm (oracle = "port" of the reference algorithm; the Rust reference itself cannot be built here)
# ------------------------------------------------------------------------------------------------------------------
def cpu_tree_throughput(log2_sample, reps=1):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as orc
    orc.build()
    n = 1 << log2_sample
    rows = splitmix_numpy(0, n * WIDTH).reshape(n, WIDTH)
    threads = orc.max_threads()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.merkle_tree_new(rows, CAP_HEIGHT, threads=threads, fast=True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n / best, threads, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None, 0
    log2_sample = 20
    for _ in range(args.warmup):
        cpu_tree_throughput(log2_sample)
    times = []
    threads = 1
    for _ in range(args.steps):
        v, threads, dt = cpu_tree_throughput(log2_sample)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = (1 << log2_sample) / (ms * 1e-3)
    sample = "MerkleTree::new on 2^%d x %d leaves per step (bounded sample of the 2^%d workload; throughput is linear in n)" % (
        log2_sample, WIDTH, LOG2_LEAVES_PER_GPU)
    line = {
        "impl": "reference", "metric": "poseidon_goldilocks_merkle_leaves_per_sec", "value": value, "unit": "leaves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks field)", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "leaves/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference path (oracle/pmt_oracle.c, OpenMP fork-join like rayon, "
                                 "fast partial rounds): the Rust reference cannot be built in this image (no cargo; plonky2 un-vendored)"},
        "e2e": {"value": value, "unit": "leaves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line, 0


def workload_config(n_gpus):
    return {"workload": "plonky2 MerkleTree::new, 2^%d leaves x %d felts per GPU, cap_height %d, upstream digests layout; "
                        "N GPUs = one tree of N*2^%d leaves, subtree-sharded, NCCL all_gather of the roots" % (
                            LOG2_LEAVES_PER_GPU, WIDTH, CAP_HEIGHT, LOG2_LEAVES_PER_GPU),
            "leaves_per_gpu": 1 << LOG2_LEAVES_PER_GPU, "leaf_width": WIDTH, "cap_height": CAP_HEIGHT,
            "global_leaves": n_gpus << LOG2_LEAVES_PER_GPU, "parallelism": "subtree-shard x%d" % n_gpus,
            "l2_policy": "inputs (512 MiB leaves + 1 GiB digests per GPU) are larger than the 126 MB L2"}


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- libpmt has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from plonky2_merkle_trees_b200 import _lib, build as pmt_build, sharded
    from plonky2_merkle_trees_b200.device import dptr
    pmt_build.build()
    ctx = _lib.Context(local_rank)
    eng = sharded.CudaEngine(ctx)

    n_local = 1 << LOG2_LEAVES_PER_GPU
    n_total = world * n_local
    d_leaves = splitmix_torch(rank * n_local * WIDTH, n_local * WIDTH, dev).view(n_local, WIDTH)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return sharded.build_sharded_tree(d_leaves, n_total, CAP_HEIGHT, eng)

    tree = None
    for _ in range(max(args.warmup, 3)):
        tree = step()
    barrier()

    # ---- timed region: K steps, CUDA events on the stream the kernels are launched on ---------------------------------
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank)
    launches0 = ctx.launches
    ctx.profile(True)
    clocks.start()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        tree = step()
    e1.record(stream)
    barrier()
    clk = clocks.stop()
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launches - launches0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = n_total / (ms_step * 1e-3)

    # ---- e2e: host-buffer C ABI call, pinned host memory, H2D + D2H inside the timed region ------------------------------
    from plonky2_merkle_trees_b200._lib import u64p
    from plonky2_merkle_trees_b200.device import bind_to_gpu_numa_node
    import ctypes as C
    # host buffers on the GPU's own NUMA node (one rank per GPU: otherwise 8 ranks' pinned buffers land wherever each
    # process happened to run); the affinity is restored before the CPU baseline below uses the host cores
    numa_node, prev_affinity = (None, None) if os.environ.get("PMT_NO_NUMA_BIND") else bind_to_gpu_numa_node(local_rank)
    h_leaves = torch.empty((n_local, WIDTH), dtype=torch.int64).pin_memory()
    h_leaves.copy_(d_leaves)
    n_dig = 2 * (n_local - 1)
    h_digests = torch.empty((n_dig, 4), dtype=torch.int64).pin_memory()
    h_cap = torch.empty((1, 4), dtype=torch.int64).pin_memory()
    del tree
    torch.cuda.empty_cache()

    def e2e_step():
        ctx.call("pmt_merkle_tree_build", C.cast(h_leaves.data_ptr(), u64p), n_local, WIDTH, CAP_HEIGHT,
                 C.cast(h_digests.data_ptr(), u64p), C.cast(h_cap.data_ptr(), u64p))
        if world > 1:   # exchange the roots from host memory (32 B per rank) and finish the top on every rank
            roots = torch.empty((world, 4), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(roots, h_cap.to(dev))
            top = eng.top_levels(roots, CAP_HEIGHT)
            ctx.sync()
            return top[-1].cpu()
        return h_cap

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    e2e_value = n_total / (e2e_ms * 1e-3)
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)

    # parity spot check of what was just measured (size-independent property: leaf digests are the canonical no-op copy)
    hd = h_digests.numpy().view(np.uint64)
    hl = h_leaves.numpy().view(np.uint64)
    assert np.array_equal(hd[0], hl[0]) and np.array_equal(hd[1], hl[1]) and np.array_equal(hd[4], hl[2]), "leaf digests are not the no-op copy"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None, 0

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------
    peaks, peak_src = measured_peaks()
    lvl = prof.get("k_level", {"launches": 0, "ms": 0.0, "units": 0.0})
    per_launch_ms = lvl["ms"] / max(lvl["launches"], 1)
    perms_per_s = lvl["units"] / (lvl["ms"] * 1e-3) if lvl["ms"] else 0.0
    mac32 = perms_per_s * MAC32_PER_PERM
    hbm_gbs = perms_per_s * ALG_BYTES_PER_PERM / 1e9
    total_prof_ms = sum(v["ms"] for v in prof.values()) or 1.0
    traffic, hardware = None, None
    try:   # one `ncu --set full` capture of this kernel (profiles/), scaled to the average launch of this run
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_r1c.json")) as f:
            cap = json.load(f)
        traffic = cap.get("k_level_dram_bytes_per_permutation") * lvl["units"] / max(lvl["launches"], 1)
        hardware = dict(cap.get("hardware_view", {}), source="profiles/k_level_r1c_summary.txt (ncu --set full, not this run)")
    except Exception:
        pass
    roofline = {
        "kernel": "k_level<Plonky2> (one two_to_one per thread)", "bound": "int32-imad (fma-heavy pipe)",
        "achieved": mac32 / 1e12, "peak": PEAK_MAC32_PER_S / 1e12, "unit": "TMAC32/s", "frac": mac32 / PEAK_MAC32_PER_S,
        "peak_source": "measured multiply-accumulate issue peak of the fma/fp64 pipes (32-bit IMAD 63.7 lanes/clk/SM), "
                       "tools/perm_bench.cu (profiles/pipes_r1.jsonl)",
        "imad_wide_only_peak": IMAD_WIDE_PEAK_PER_S / 1e12,
        "algorithmic_work_per_unit": "%d MAC32 per permutation (SURVEY.md 8(d))" % MAC32_PER_PERM,
        "perms_per_s": perms_per_s, "launches_timed": lvl["launches"], "avg_launch_ms": per_launch_ms,
        "share_of_step": lvl["ms"] / total_prof_ms,
        "hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_gbs / peaks["hbm_gbs"],
                "peak_source": peak_src + " MEASURED_PEAKS.json", "algorithmic_bytes_per_unit": ALG_BYTES_PER_PERM},
        "traffic": traffic,
        # the kernel takes algebraic shortcuts (paired partial rounds, frequency-domain MDS), so the spec-count figure above
        # is the analogue of 2N^3 for a GEMM; this is what the SM actually executed
        "hardware_view": hardware,
    }

    # ---- CPU baseline on this box's host cores (bounded sample) ---------------------------------------------------------
    cpu_log2 = 22
    cpu_v, cpu_threads, cpu_dt = cpu_tree_throughput(cpu_log2)
    cpu = {"value": cpu_v, "unit": "leaves/s", "cores": cpu_threads, "kind": "port",
           "sample": "one MerkleTree::new over 2^%d x %d leaves (%.2f s wall on %d threads) with the oracle's C restatement "
                     "(OpenMP fork-join, fast partial rounds)" % (cpu_log2, WIDTH, cpu_dt, cpu_threads)}

    line = {
        "metric": "poseidon_goldilocks_merkle_leaves_per_sec", "value": value, "unit": "leaves/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks field, 32-bit IMAD limbs)", "data": "synthetic",
        "config": workload_config(world),
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "leaves/s", "h2d_bytes_per_step": n_local * WIDTH * 8,
                "d2h_bytes_per_step": (n_dig + 1) * 32, "ms_per_step": e2e_ms,
                "api": "pmt_merkle_tree_build (host buffers, pinned; every digest downloaded)",
                "host_numa_node": numa_node},
        "gpu_launches": launches,
        "kernels": prof,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if world > 1:
        dist.destroy_process_group()
    return line, 0


class StdoutToStderr:
    """Everything written to fd 1 while active goes to stderr (NCCL prints its version banner to stdout when NCCL_DEBUG is
    set on the box); the JSON line is printed after restore(), so stdout carries exactly one line."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def restore(self):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    guard = StdoutToStderr()
    line, rc = run_reference(args) if args.impl == "reference" else run_ours(args)
    guard.restore()
    if line is not None:
        print(json.dumps(line), flush=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())

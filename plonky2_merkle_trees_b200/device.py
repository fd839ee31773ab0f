"""Device-memory plumbing for the host mirror: torch tensors are only buffers here (allocation, H2D/D2H, streams)."""
import ctypes as C
import os

import numpy as np
import torch


def dev_u64(shape, device):
    """Uninitialised device buffer of u64 (stored as int64: torch has no arithmetic on uint64 and needs none here)."""
    return torch.empty(shape, dtype=torch.int64, device=device)


def to_device(arr, device, pinned=False):
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.uint64))
    t = torch.from_numpy(a.view(np.int64))
    if pinned:
        t = t.pin_memory()
    return t.to(device, non_blocking=pinned)


def to_host(t):
    return t.detach().cpu().numpy().view(np.uint64)


def dptr(t):
    return C.c_void_p(t.data_ptr())


def order_after_torch(ctx, device=None):
    """Make the ctx stream wait for the work already enqueued on torch's current stream (one event record + wait, no host
    sync).  The ctx stream is non-blocking, so a tensor produced by torch kernels (e.g. synthetic leaves) is otherwise not
    guaranteed to be complete when a libpmt kernel reads it."""
    dev = torch.device("cuda", ctx.device) if device is None else device
    cur = torch.cuda.current_stream(dev)
    if cur.cuda_stream != (ctx.stream or 0) and not cur.query():   # query(): nothing pending -> nothing to wait for
        torch.cuda.ExternalStream(ctx.stream, device=dev).wait_stream(cur)


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that host buffers allocated afterwards
    (first touch, cudaHostAlloc) are local to the PCIe root the DMA goes through.  With one rank per GPU and unbound
    processes the pinned buffers of 8 ranks land on whatever node each process happened to run on.  Returns
    (node, previous affinity) or (None, None) when the topology is not visible (containers, single-node hosts): then
    nothing is changed.  Undo with os.sched_setaffinity(0, previous)."""
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None, None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _cpulist(f.read())
        previous = os.sched_getaffinity(0)
        cpus &= previous
        if not cpus or cpus == previous:
            return None, None
        os.sched_setaffinity(0, cpus)
        return node, previous
    except Exception:   # no driver / no sysfs / restricted container: leave the process alone
        return None, None

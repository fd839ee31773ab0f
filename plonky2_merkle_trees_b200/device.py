"""Device-memory plumbing for the host mirror: torch tensors are only buffers here (allocation, H2D/D2H, streams)."""
import ctypes as C

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

"""ctypes binding of libpmt.so (the C ABI in include/pmt.h).  No CPU fallback: if the library or a CUDA device is
missing, importing works but creating a Context raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("PMT_SO", os.path.join(HERE, "libpmt.so"))  # PMT_SO: A/B testing of kernel builds

PMT_OK, PMT_E_INVALID_ARG, PMT_E_NOT_POW2, PMT_E_OOM, PMT_E_CUDA, PMT_E_RANGE, PMT_E_NCCL = 0, -1, -2, -3, -4, -5, -6

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
u32p = C.POINTER(C.c_uint32)
_SZ, _U32, _INT, _VP = C.c_size_t, C.c_uint32, C.c_int, C.c_void_p

# name -> (restype, argtypes); ctx is always the first argument unless stated
_SIGNATURES = {
    "pmt_init": (_INT, [C.POINTER(_VP), _INT]),
    "pmt_destroy": (None, [_VP]),
    "pmt_last_error": (C.c_char_p, [_VP]),
    "pmt_version": (C.c_char_p, []),
    "pmt_device_id": (_INT, [_VP]),
    "pmt_set_stream": (_INT, [_VP, _VP]),
    "pmt_get_stream": (_VP, [_VP]),
    "pmt_sync": (_INT, [_VP]),
    "pmt_kernel_launches": (C.c_uint64, [_VP]),
    "pmt_profile_enable": (_INT, [_VP, _INT]),
    "pmt_profile_read": (_INT, [_VP, C.c_char_p, _SZ]),
    "pmt_malloc": (_INT, [_VP, _SZ, C.POINTER(_VP)]),
    "pmt_free": (_INT, [_VP, _VP]),
    "pmt_memcpy_h2d": (_INT, [_VP, _VP, _VP, _SZ]),
    "pmt_memcpy_d2h": (_INT, [_VP, _VP, _VP, _SZ]),
    "pmt_host_register": (_INT, [_VP, _VP, _SZ]),
    "pmt_host_unregister": (_INT, [_VP, _VP]),
    "pmt_host_alloc": (_INT, [_VP, _SZ, C.POINTER(_VP)]),
    "pmt_host_free": (_INT, [_VP, _VP]),
    "pmt_permute": (_INT, [_VP, u64p, _SZ, u64p]),
    "pmt_hash_two_to_one": (_INT, [_VP, u64p, u64p, _SZ, u64p]),
    "pmt_hash_or_noop": (_INT, [_VP, u64p, _SZ, _SZ, u64p]),
    "pmt_hash_no_pad": (_INT, [_VP, u64p, _SZ, _SZ, u64p]),
    "pmt_permute_dev": (_INT, [_VP, _VP, _SZ, _VP]),
    "pmt_hash_two_to_one_dev": (_INT, [_VP, _VP, _VP, _SZ, _VP]),
    "pmt_hash_rows_dev": (_INT, [_VP, _VP, _SZ, _SZ, _INT, _VP]),
    "pmt_simple_tree_build": (_INT, [_VP, u64p, _SZ, u64p, u64p]),
    "pmt_simple_tree_build_dev": (_INT, [_VP, _VP, _SZ, _VP, _VP]),
    "pmt_simple_tree_prove_dev": (_INT, [_VP, _VP, _SZ, _VP, _SZ, _VP]),
    "pmt_simple_tree_verify_dev": (_INT, [_VP, _VP, _VP, _SZ, _VP, _VP, _SZ, _VP]),
    "pmt_merkle_tree_build": (_INT, [_VP, u64p, _SZ, _SZ, _U32, u64p, u64p]),
    "pmt_merkle_tree_build_dev": (_INT, [_VP, _VP, _SZ, _SZ, _U32, _VP, _VP]),
    "pmt_merkle_tree_build_multi": (_INT, [C.POINTER(_VP), _SZ, u64p, _SZ, _SZ, _U32, u64p, u64p]),  # first arg: ctx array
    "pmt_merkle_tree_build_from_columns_dev": (_INT, [_VP, _VP, _SZ, _SZ, _INT, _U32, _VP, _VP, _VP]),
    "pmt_merkle_prove_dev": (_INT, [_VP, _VP, _SZ, _U32, _VP, _SZ, _VP]),
    "pmt_merkle_verify_dev": (_INT, [_VP, _VP, _SZ, _VP, _SZ, _VP, _U32, _VP, _SZ, _VP]),
    "pmt_merkle_tree_build_multi_dev": (_INT, [C.POINTER(_VP), _SZ, C.POINTER(_VP), _SZ, _SZ, _U32, C.POINTER(_VP), _VP, _VP, _VP]),  # ctx array first
    "pmt_nccl_unique_id": (_INT, [_VP, _VP]),
    "pmt_comm_init": (_INT, [_VP, _VP, _INT, _INT]),
    "pmt_comm_destroy": (_INT, [_VP]),
    "pmt_merkle_tree_build_sharded_dev": (_INT, [_VP, _VP, _SZ, _SZ, _U32, _VP, _VP, _VP, _VP]),
    "pmt_comm_uses_peer_memory": (_INT, [_VP]),
    "pmt_mmr_shard_plan": (_INT, [_SZ, _SZ, u32p, C.POINTER(_SZ), C.POINTER(_SZ)]),  # no ctx: pure index math
    "pmt_mmr_build_sharded_dev": (_INT, [_VP, _VP, _SZ, _VP, _VP, _VP, _VP, _VP]),
    "pmt_top_levels_dev": (_INT, [_VP, _VP, _SZ, _U32, _VP]),
    "pmt_top_levels_batch_dev": (_INT, [_VP, _VP, _SZ, _SZ, _U32, _VP]),
    "pmt_mmr_size": (_SZ, [_SZ]),
    "pmt_mmr_index": (_SZ, [_SZ]),
    "pmt_mmr_extend": (_INT, [_VP, u64p, _SZ, u64p, _SZ]),
    "pmt_mmr_extend_dev": (_INT, [_VP, _VP, _SZ, _VP, _SZ]),
    "pmt_mmr_extend_multi": (_INT, [C.POINTER(_VP), _SZ, u64p, _SZ, u64p, _SZ]),  # first arg: ctx array
    "pmt_mmr_multi_plan": (_INT, [_SZ, _SZ, _SZ, u32p, C.POINTER(_SZ), C.POINTER(_SZ)]),  # no ctx: pure index math
    "pmt_mmr_peaks_dev": (_INT, [_VP, _VP, _SZ, _VP, u32p]),
    "pmt_mmr_bag_dev": (_INT, [_VP, _VP, _SZ, _VP]),
    "pmt_mmr_prove_dev": (_INT, [_VP, _VP, _SZ, _VP, _SZ, _VP, _VP, _VP]),
    "pmt_mmr_verify_dev": (_INT, [_VP, _VP, _SZ, _VP, _VP, _VP, _VP, _U32, _VP, _VP]),
    "pmt_mmr_bag": (_INT, [_VP, u64p, _SZ, u64p]),
    "pmt_mmr_peaks": (_INT, [_VP, u64p, _SZ, u64p, u32p]),
    # host-buffer proofs (gathers on the host) and verification (upload, fold on the GPU, download)
    "pmt_simple_tree_prove": (_INT, [_VP, u64p, _SZ, u64p, _SZ, u64p]),
    "pmt_merkle_prove": (_INT, [_VP, u64p, _SZ, _U32, u64p, _SZ, u64p]),
    "pmt_mmr_prove": (_INT, [_VP, u64p, _SZ, u64p, _SZ, u64p, u8p, u32p]),
    "pmt_simple_tree_verify": (_INT, [_VP, u64p, u64p, _SZ, u64p, u64p, _SZ, u8p]),
    "pmt_merkle_verify": (_INT, [_VP, u64p, _SZ, u64p, _SZ, u64p, _U32, u64p, _SZ, u8p]),
    "pmt_mmr_verify": (_INT, [_VP, u64p, _SZ, u64p, u8p, u32p, u64p, _U32, u64p, i8p]),
}

_lib = None


class PmtError(RuntimeError):
    """Non-zero status from libpmt; `.code` is the PMT_E_* value.  The Rust shim turns these into the reference's panics."""

    def __init__(self, code, msg):
        super().__init__("libpmt error %d: %s" % (code, msg))
        self.code = code


def load():
    """dlopen libpmt.so and set the signatures.  Raises if the library was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise ImportError("libpmt.so is not built: run `python -m plonky2_merkle_trees_b200.build` "
                              "(there is no CPU fallback)")
        lib = C.CDLL(SO)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def signatures():
    """name -> (restype, argtypes) as bound on the library (tests compare them with include/pmt.h)"""
    return dict(_SIGNATURES)


def as_u64(x, shape=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64))
    return a if shape is None else a.reshape(shape)


def ptr(a):
    return a.ctypes.data_as(u64p)


class Context:
    """One pmt_ctx: one CUDA device, one stream.  Not thread-safe (use one per host thread / rank)."""

    def __init__(self, device=0):
        self.lib = load()
        h = _VP()
        rc = self.lib.pmt_init(C.byref(h), int(device))
        if rc != PMT_OK:
            raise PmtError(rc, "pmt_init failed: no usable CUDA device %d (libpmt has no CPU fallback)" % device)
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.pmt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != PMT_OK:
            raise PmtError(rc, self.lib.pmt_last_error(self.h).decode())

    def call(self, name, *args):
        if name.endswith("_dev"):
            # device-pointer entry points read tensors that torch may still be producing on ITS current stream; the ctx
            # stream is non-blocking, so order it after torch's stream first (event wait, no host sync)
            from .device import order_after_torch
            order_after_torch(self)
        self.check(getattr(self.lib, name)(self.h, *args))

    def sync(self):
        self.call("pmt_sync")

    @property
    def launches(self):
        return int(self.lib.pmt_kernel_launches(self.h))

    @property
    def stream(self):
        return self.lib.pmt_get_stream(self.h)

    def profile(self, on):
        self.call("pmt_profile_enable", 1 if on else 0)

    def profile_read(self):
        """-> {kernel: {"launches": n, "ms": total, "units": permutations}} and resets the records."""
        buf = C.create_string_buffer(1 << 14)
        self.call("pmt_profile_read", buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms, units = line.split()
            out[name] = {"launches": int(n), "ms": float(ms), "units": float(units)}
        return out

    def set_stream(self, cuda_stream_ptr):
        self.call("pmt_set_stream", _VP(cuda_stream_ptr) if cuda_stream_ptr else None)


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]

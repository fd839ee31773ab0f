"""Build libpmt.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libpmt.so")
SOURCES = [os.path.join(HERE, "csrc", "pmt_api.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", f) for f in
                  ("goldilocks.cuh", "poseidon.cuh", "poseidon_coop.cuh", "poseidon_constants.cuh", "poseidon_freq.cuh",
                   "poseidon_freq_constants.cuh", "merkle_kernels.cuh")] + [
    os.path.join(HERE, "..", "include", "pmt.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-cudart", "static", "-ldl"]


def needs_build():
    return not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS)


def build(force=False, verbose=False):
    """Safe under concurrency (torchrun starts one process per GPU, each of which calls this): one process holds the lock
    and compiles into a temporary file that is renamed over libpmt.so atomically; the others wait and find it up to date."""
    if not force and not needs_build():
        return SO
    import fcntl
    with open(SO + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or needs_build():
                nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
                tmp = "%s.tmp.%d" % (SO, os.getpid())
                cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
                try:
                    subprocess.check_call(cmd)
                    os.replace(tmp, SO)
                finally:
                    if os.path.exists(tmp):
                        os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))

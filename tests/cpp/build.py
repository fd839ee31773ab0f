"""Compile tests/cpp/test_host_mirror.cpp (the reference's tests against include/pmt.hpp) into tests/cpp/_build/.
The binary finds libpmt.so and the oracle through $ORIGIN-relative rpaths, so it travels to the GPU box with the tree."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "test_host_mirror.cpp")
BIN = os.path.join(HERE, "_build", "test_host_mirror")
DEPS = [SRC, os.path.join(ROOT, "include", "pmt.hpp"), os.path.join(ROOT, "include", "pmt.h"),
        os.path.join(ROOT, "oracle", "pmt_oracle.h")]


def build(force=False):
    libs = [os.path.join(ROOT, "plonky2_merkle_trees_b200", "libpmt.so"), os.path.join(ROOT, "oracle", "libpmt_oracle.so")]
    for lib in libs:
        if not os.path.exists(lib):
            raise RuntimeError("%s is missing: run __graft_entry__.build() first" % lib)
    if not force and os.path.exists(BIN) and all(os.path.getmtime(d) <= os.path.getmtime(BIN) for d in DEPS + libs):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-pthread", "-o", BIN, SRC,
                           "-L" + os.path.dirname(libs[0]), "-lpmt", "-L" + os.path.dirname(libs[1]), "-lpmt_oracle",
                           "-Wl,-rpath,$ORIGIN/../../../plonky2_merkle_trees_b200:$ORIGIN/../../../oracle"])
    return BIN


if __name__ == "__main__":
    print(build(force=True))

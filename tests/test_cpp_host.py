"""The C++ host mirror (include/pmt.hpp): the reference's own Rust tests restated in C++ over the C ABI
(tests/cpp/test_host_mirror.cpp).  CPU: it compiles against the header, the index-math tests pass and the engine refuses
to start without a device (no CPU fallback).  GPU: every test passes, checked against the known answers of
simple_merkle_tree.rs:117-310 / merkle_mountain_ranges.rs:278-374 and the oracle."""
import importlib.util
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _binary():
    import __graft_entry__ as entry
    entry.build()
    spec = importlib.util.spec_from_file_location("cpp_build", os.path.join(HERE, "cpp", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_cpp_mirror_compiles_and_refuses_without_device():
    import torch
    exe = _binary()
    if torch.cuda.is_available():
        pytest.skip("a device is present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "no CPU fallback" in r.stdout and "index-math tests passed" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_tests_on_gpu():
    exe = _binary()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout, r.stdout
    for name in ("test_build_merkle_tree_4_leaves", "test_verify_merkle_proof_16", "test_get_proof", "test_plonky2_merkle_tree"):
        assert name in r.stdout

"""Runs the C++ drop-in binaries on the GPU: the reference's public API
(#include <spblas/spblas.hpp>, -DSPBLAS_ENABLE_B200) compiled against the reference's own
headers + integration/enable_b200.patch + include/spblas/vendor/b200/.  The binaries are
built where /root/reference exists (tests/cpp/Makefile, by __graft_entry__.build()) and
travel to the GPU box under build/cpp_tests/."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "build", "cpp_tests")


def _run(name):
    path = os.path.join(BIN, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs /root/reference at build time)")
    r = subprocess.run([path], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    return r


def test_dropin_harness(cuda):
    r = _run("dropin_test")
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 failures" in r.stdout


def test_reference_own_device_spmv_test_unmodified(cuda):
    """test/gtest/device/spmv_test.cpp of the reference, byte-for-byte, on this backend."""
    r = _run("ref_device_spmv_test")
    assert r.returncode == 0, r.stdout[-2000:]
    for name in ("thrust_CsrView.SpMV", "thrust_CsrView.SpMV_Ascaled", "thrust_CsrView.SpMV_BScaled"):
        assert f"[  OK  ] {name}" in r.stdout


def test_reference_own_device_example_unmodified(cuda):
    r = _run("ref_example_device_spmv")
    assert r.returncode == 0
    assert "Example is completed!" in r.stdout

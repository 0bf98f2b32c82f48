"""The C-ABI library loads without a GPU and exports every symbol include/spblas_b200.h
declares (no compute calls here)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "spblas_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"SPBLAS_B200_API\s+[\w\s\*]+?\b(spblas_b200_\w+)\s*\(", text)
    return sorted(set(names))


def test_header_and_binding_list_agree():
    from spblas_reference_b200 import _cabi
    assert _declared() == sorted(_cabi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from spblas_reference_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        _cabi.build()
    L = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), f"{name} not exported"
    assert L.spblas_b200_version() == 100
    L.spblas_b200_status_string.restype = ctypes.c_char_p
    assert L.spblas_b200_status_string(2) == b"shape mismatch"


def test_library_is_sm100a_and_has_no_cusparse():
    """The product is hand-written CUDA for sm_100a: the binary carries sm_100a code and
    does not link cuSPARSE."""
    import subprocess
    from spblas_reference_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        _cabi.build()
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", _cabi.LIB_PATH],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "cusparse" not in ldd.lower()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the package or the C++ headers may
    reference it."""
    bad = []
    for base in ("spblas_reference_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".hpp", ".h", ".cu", ".cuh", "Makefile")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"(import|from)\s+oracle|liboracle|libspblas_ref|oracle/", text):
                        # comments that merely mention the oracle's path are fine in .cu docs
                        if f.endswith((".cu", ".cuh", ".hpp", ".h")) and "#include" not in "".join(
                                l for l in text.splitlines() if "oracle" in l):
                            continue
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad

"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/xaac_b200.h declares, the ctypes binding covers exactly that set, and the product package never
reaches into oracle/."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "xaac_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xaac_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from libxaac_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the product library first (make lib / __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 9
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/xaac_b200.h but not exported"


def test_binding_matches_header():
    from libxaac_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import libxaac_b200
    with pytest.raises(libxaac_b200.XaacB200Error):
        libxaac_b200.Context(0)


def test_product_never_imports_oracle():
    """No import / include / dlopen of anything under oracle/ from the product package."""
    pkg = os.path.join(ROOT, "libxaac_b200")
    bad = re.compile(r"(^\s*(import|from)\s+\S*oracle)|(#\s*include[^\n]*oracle)|(liboracle\.so)|(libxaac_ref)|(CDLL\([^)]*oracle)|(dlopen\([^)]*oracle)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")):
                for line in open(os.path.join(dirpath, f), errors="replace").read().splitlines():
                    assert not bad.search(line), f"{f}: product code reaches into oracle/: {line.strip()}"

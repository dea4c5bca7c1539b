"""CPU-side checks of the C-ABI boundary: the library loads without a GPU and exports exactly the
symbols include/cgpt.h declares; the ctypes table covers all of them."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from conftest import PKG, ROOT


@pytest.fixture(scope="module")
def lib_path():
    path = os.path.join(PKG, "codonlm_b200", "libcgpt_b200.so")
    if not os.path.exists(path):
        subprocess.check_call([sys.executable, os.path.join(PKG, "build.py")])
    return path


def _declared():
    text = open(os.path.join(ROOT, "include", "cgpt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cgpt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cgpt.h but not exported"


def test_ctypes_table_matches_header(lib_path):
    from codonlm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.cgpt_version() == 100


def test_no_cpu_fallback():
    import torch
    from codonlm_b200 import _lib, ops
    with pytest.raises(_lib.CgptError, match="no CPU"):
        ops.segment_ids(torch.zeros((1, 4), dtype=torch.int64), 3)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").lower() or f == "README.md", (dirpath, f)

"""The C-ABI library loads and exports every symbol include/gmix_b200.h declares; without a GPU the
product fails loudly instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _declared():
    txt = open(os.path.join(ROOT, "include", "gmix_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gmx_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import gmix_b200
    lib = ctypes.CDLL(gmix_b200.library_path())
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_bound():
    import gmix_b200
    lib = gmix_b200.load_library()
    assert b"sm_100a" in lib.gmx_version()
    assert gmix_b200.compress_bound(0) >= 6 and gmix_b200.compress_bound(65536) > 65536


def test_no_cpu_fallback_without_gpu():
    import torch
    import gmix_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gmix_b200.GmixError):
        gmix_b200.Context(0)


def test_product_never_references_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "gmix_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle_lib|gmix_oracle|oracle/", txt):
                    bad.append(f)
    assert not bad, bad

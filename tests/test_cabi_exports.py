"""CPU: the C-ABI shared library loads and exports every symbol include/isonform_b200.h declares
(no compute calls -- there is no GPU here), and the product refuses to run without a device."""
import os
import re

import pytest

from isonform_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "isonform_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(isof_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    L = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.EXPORTS) == names


def test_version():
    assert _lib.load().isof_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.IsofError):
        _lib.Context(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under isonform_b200/ may reference it."""
    pkg = os.path.join(ROOT, "isonform_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libisof_oracle" not in src, f

"""The C-ABI shared library loads and exports every symbol declared in include/matten_b200.h;
without an sm_100 device every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import matten_b200

from tests.helpers import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "matten_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from matten_b200 import _lib

    lib = _lib.load()
    names = _declared_symbols()
    assert "mt_conv_fwd" in names and "mt_csr_by_key" in names and len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/matten_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.mt_abi_version() == matten_b200.ABI_VERSION == 2


def test_no_cpu_fallback():
    from matten_b200 import ops

    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.edge_sh(x, 2)
    if not torch.cuda.is_available():
        from matten_b200 import _lib

        lib = _lib.load()
        rc = lib.mt_edge_sh(0, None, 4, 2, 1, None, None)
        assert rc != 0
        assert b"CUDA" in lib.mt_last_error() or b"device" in lib.mt_last_error()


def test_product_does_not_import_oracle():
    import subprocess
    import sys

    code = ("import sys; import matten_b200.model_factory, matten_b200.nn, matten_b200.ops; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for root, _, files in os.walk(os.path.join(ROOT, "matten_b200")):
        for f in files:
            if f.endswith(".py"):
                s = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M), f

"""CPU-only: the C-ABI library loads and exports every symbol include/jrc_cuda.h declares;
without a GPU the product refuses to run (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "jrc_cuda.h")).read()
    return sorted(set(re.findall(r"JRC_API\s+[\w\s\*]+?\b(jrc_\w+)\s*\(", src)))


def test_header_declares_what_the_binding_lists(jrc):
    assert declared_symbols() == sorted(jrc.EXPORTS)


def test_library_exports_every_declared_symbol(jrc):
    assert os.path.exists(jrc.LIB_PATH), "libjrc_cuda.so not built (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(jrc.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.jrc_abi_version.restype = ctypes.c_int32
    assert lib.jrc_abi_version() == 2


def test_no_cpu_fallback_without_a_device(jrc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(jrc.JrcError) as e:
        jrc.Chain(64, 4, 2, 4, 0, 8, 16)
    assert e.value.status == 3          # JRC_ERR_NO_DEVICE
    with pytest.raises(jrc.JrcError):
        jrc.mimo_ofdm_radar(64, 4, 2, 4, 5, False, False, 8, 8, False, "/tmp/x.csv")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "jrc_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f

"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/b200vit.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200vit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200vit_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    import uncertainty_vit_b200 as pkg
    lib = pkg._lib.lib()
    names = _declared()
    assert len(names) >= 40
    raw = ctypes.CDLL(pkg._lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/b200vit.h but not exported"
    assert sorted(pkg._lib.exported_symbols()) == names, "python prototypes and header disagree"
    assert lib.b200vit_abi_version() == 8


def test_header_cites_reference_lines():
    src = open(os.path.join(ROOT, "include", "b200vit.h")).read()
    assert len(re.findall(r"\w+\.py:\d+", src)) >= 15


def test_product_package_never_imports_oracle():
    pk = os.path.join(ROOT, "uncertainty-vit_b200")
    for dp, _, files in os.walk(pk):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions the oracle: product code must not use it"


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    import uncertainty_vit_b200 as pkg
    with pytest.raises(pkg._lib.B200VitError):
        pkg.ops.cast_bf16(torch.zeros(8))

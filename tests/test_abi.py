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


def test_ctypes_prototypes_match_the_header_arity_and_kinds():
    """Every prototype in _lib._PROTOS has as many arguments as the declaration in include/b200vit.h, and pointer / integer / float kinds agree
    (a drifted ctypes signature would pass garbage across the boundary without any error)."""
    import uncertainty_vit_b200 as pkg
    C = ctypes
    src = open(os.path.join(ROOT, "include", "b200vit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"#.*", "", src)
    decls = dict()
    for m in re.finditer(r"\b[\w\s\*]+?\b(b200vit_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        decls[name] = [] if args in ("", "void") else [a.strip() for a in args.split(",")]

    def kind_of_decl(a):
        if "*" in a:
            return "ptr"
        t = a.rsplit(None, 1)[0] if len(a.split()) > 1 else a
        return "float" if re.search(r"\b(float|double)\b", t) else "int"

    def kind_of_ctype(t):
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return "float" if t in (C.c_float, C.c_double) else "int"

    protos = pkg._lib._PROTOS
    assert set(protos) == set(decls)
    for name, (_, argtypes) in protos.items():
        want = decls[name]
        assert len(argtypes) == len(want), f"{name}: {len(argtypes)} ctypes arguments, header declares {len(want)}"
        for i, (t, a) in enumerate(zip(argtypes, want)):
            assert kind_of_ctype(t) == kind_of_decl(a), f"{name}: argument {i} ({a!r}) is {kind_of_decl(a)} in the header, {t} in _PROTOS"


def test_ctypes_structs_match_the_header_field_for_field():
    """b200vit_gemm_desc / b200vit_d2v_desc: same field names, order and kinds as the ctypes mirrors (a dropped field shifts every later one)."""
    import uncertainty_vit_b200 as pkg
    C = ctypes
    src = open(os.path.join(ROOT, "include", "b200vit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for cname, mirror in (("b200vit_gemm_desc", pkg._lib.GemmDesc), ("b200vit_d2v_desc", pkg._lib.D2VDesc)):
        body = re.search(r"typedef struct " + cname + r"\s*\{(.*?)\}\s*" + cname + r"\s*;", src, flags=re.S).group(1)
        fields = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            head, *rest = [p.strip() for p in stmt.split(",")]
            base = head.rsplit(None, 1)[0]
            names = [head.rsplit(None, 1)[1]] + rest
            for nm in names:
                ptr = "*" in nm or "*" in base
                is_float = bool(re.search(r"\b(float|double)\b", base)) and not ptr
                fields.append((nm.replace("*", "").strip(), "ptr" if ptr else ("float" if is_float else "int")))
        got = []
        for nm, t in mirror._fields_:
            k = "ptr" if (t in (C.c_void_p,) or (isinstance(t, type) and issubclass(t, C._Pointer))) else ("float" if t in (C.c_float, C.c_double) else "int")
            got.append((nm, k))
        assert got == fields, f"{cname}: ctypes mirror {got} != header {fields}"

"""The C-ABI library loads and exports every symbol include/b200sparse.h declares
(no compute call is made — runs without a GPU)."""
import ctypes
import os
import re

from legate_sparse import _native as N

ROOT = os.path.normpath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200sparse.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(N.lib_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b200sparse.h but not exported"


def test_python_binding_table_matches_header():
    names = set(_declared_symbols())
    assert names == set(N.SIGNATURES), names ^ set(N.SIGNATURES)
    # argument counts agree with the header prototypes
    src = open(os.path.join(ROOT, "include", "b200sparse.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, (_, args) in N.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(args), (name, n, len(args))


def test_version_and_error_string_without_gpu():
    lib = N.load()
    assert lib.b2s_version() == 100
    assert isinstance(N.last_error(), str)
    assert lib.b2s_reduce_workspace_bytes() > 0
    assert lib.b2s_spmv_plan_workspace_bytes(10, 100) > 0
    assert lib.b2s_spgemm_workspace_bytes(10, 100, 10) > 0

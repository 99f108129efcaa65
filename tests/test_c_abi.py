"""The C-ABI shared library loads and exports every symbol that include/asvgp_b200.h declares, and the ctypes
signature table matches the header's arity (no compute calls: this runs without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def declared():
    text = open(os.path.join(ROOT, "include", "asvgp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"ASVGP_API\s+([\w\s\*]+?)\s*\b(asvgp_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(3).strip()
        out[m.group(2)] = 0 if args in ("", "void") else args.count(",") + 1
    assert len(out) >= 9
    return out


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__

    __graft_entry__.build()
    from asvgp_b200 import _lib

    return _lib


def test_every_declared_symbol_is_exported(declared, lib):
    handle = lib.load()
    for name in declared:
        assert hasattr(handle, name), "libasvgp_sm100a.so does not export %s" % name
    assert handle.asvgp_abi_version() == 1
    assert handle.asvgp_last_error() == b""


def test_ctypes_table_matches_header(declared, lib):
    bound = dict(lib.SIGNATURES)
    bound.update({k: v[1] for k, v in lib.VALUE_FUNCTIONS.items()})
    for name, argtypes in bound.items():
        assert name in declared, "%s is bound in _lib.py but not declared in the header" % name
        assert len(argtypes) == declared[name], "%s: %d ctypes args vs %d in the header" % (name, len(argtypes), declared[name])
    missing = set(declared) - set(bound) - {"asvgp_abi_version", "asvgp_last_error"}
    assert not missing, "declared but not bound: %s" % sorted(missing)


def test_bad_arguments_fail_loudly_without_a_gpu(lib):
    """Argument validation happens before any CUDA call, so it can be exercised on the CPU box."""
    handle = lib.load()
    assert handle.asvgp_workspace_bytes_1d(0, 3, 0) == -1
    assert handle.asvgp_workspace_bytes_1d(10_000, 3, 0) > 0
    with pytest.raises(lib.AsvgpNativeError, match="order"):
        lib.call("asvgp_accum_1d", None, None, 10, None, 5, 9, None, None)


def test_product_has_no_cpu_fallback():
    """Nothing under asvgp_b200/ may import the oracle, and ops must refuse to run without CUDA."""
    import torch

    pkg = os.path.join(ROOT, "asvgp_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), "%s mentions the oracle" % fn
    if not torch.cuda.is_available():
        from asvgp_b200 import _lib, ops

        with pytest.raises(_lib.AsvgpNativeError):
            ops.device()

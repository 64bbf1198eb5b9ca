"""CPU: the C-ABI library builds, loads and exports every symbol include/sdb200.h declares; the
product fails loudly (no fallback) when handed CPU tensors or when the library is missing."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sdb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from sdb200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    from sdb200 import _lib
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in sdb200.h is not exported" % n
        assert n in _lib.SIGNATURES, "symbol %s has no ctypes signature" % n
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_error_string(lib):
    assert lib.sdb_version() >= 100
    assert isinstance(lib.sdb_last_error_string(), bytes)


def test_no_cpu_fallback(lib):
    from sdb200 import _lib, ops
    x = torch.zeros(1, 2, 2, 32)
    with pytest.raises(_lib.SdbError):
        ops.groupnorm(x, torch.ones(32), torch.zeros(32), 1e-5)
    with pytest.raises(_lib.SdbError):
        ops.ddim_step(torch.zeros(4), torch.zeros(4), 0.5, 0.6, 0.3, 0.0, 0.7)


def test_missing_library_is_loud(tmp_path):
    from sdb200 import _lib
    with pytest.raises(_lib.SdbError):
        _lib.load(str(tmp_path / "nope.so"))


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected on the host before any launch (safe without a GPU)."""
    import ctypes as C
    assert lib.sdb_groupnorm_ws_bytes(1, 16, 30, 32) == -1          # C % 4 != 0
    assert lib.sdb_groupnorm_ws_bytes(2, 4096, 320, 32) > 0
    rc = lib.sdb_layernorm(None, 4, 320, 1e-5, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.sdb_last_error_string()
    rc = lib.sdb_ddim_step(None, None, None, 1.0, None, 0.5, 0.6, 0.3, 0.0, 0.7, 1.0, None, None, 0, None)
    assert rc == -1

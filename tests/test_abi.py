"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/lancedb_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from lancedb_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "lancedb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    _native.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lancedb_b200.h but not exported"
    assert sorted(_native.EXPORTS) == names, "ctypes binding and header disagree on the symbol list"


def test_abi_version_and_struct_sizes():
    lib = _native.load()
    assert lib.lgpu_abi_version() == 2
    # must match the C layout: 8 x 4-byte fields, u64, 6 pointers / 10 x 4-byte fields
    assert ctypes.sizeof(_native.IndexDesc) == 32 + 8 + 6 * 8
    assert ctypes.sizeof(_native.SearchParams) == 40


def test_no_cpu_fallback_without_gpu():
    if _native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.GpuFlat(np.zeros((4, 4), np.float32))
    from tests.util import random_index
    ix = random_index(np.random.default_rng(0), dim=16, nlist=2, m=2, n=20)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.GpuIvfPq(ix)


def test_invalid_input_is_reported_before_touching_the_device():
    from tests.util import random_index
    bad = random_index(np.random.default_rng(0), dim=30, nlist=2, m=2, n=10)     # dsub = 15
    with pytest.raises(ValueError, match="sub-vector length"):
        _native.GpuIvfPq(bad)
    lib = _native.load()
    assert lib.lgpu_index_open(None, None) == _native.LGPU_INVALID_INPUT
    assert b"null" in lib.lgpu_last_error()


def test_product_path_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under lancedb_b200/ may reference it."""
    for dp, _, files in os.walk(os.path.join(ROOT, "lancedb_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f

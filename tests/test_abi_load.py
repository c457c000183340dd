"""CPU-only checks of the drop-in boundary: the library builds/loads, exports
every symbol include/b2webp.h declares, and the host logic that needs no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from go_webp_b200 import _abi, partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "b2webp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    names = _header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    # and the Python declarations cover the header exactly
    assert sorted(_abi.SYMBOLS) == names


def test_abi_version_and_strerror():
    lib = _abi.load()
    assert lib.b2_abi_version() == 0x0104
    assert lib.b2_strerror(0) == b"ok"
    # messages are the reference's sentinels (pkg/picture/errors.go:8-21)
    assert lib.b2_strerror(_abi.B2_ERR_BAD_DIM) == b"picture has invalid width/height"
    assert lib.b2_strerror(_abi.B2_ERR_NULL) == b"a pointer parameter is nil"


def test_struct_layouts_match_header():
    assert C.sizeof(_abi.ImportJob) == 5 * 8 + 6 * 4
    assert C.sizeof(_abi.UpsampleJob) == 5 * 8 + 8 * 4
    assert C.sizeof(_abi.ArgbJob) == 3 * 8 + 4 * 4
    assert C.sizeof(_abi.PlainJob) == 5 * 8 + 8 * 4
    assert C.sizeof(_abi.AlphaIndexJob) == 3 * 8 + 4 * 4
    assert C.sizeof(_abi.ExpandJob) == 2 * 8 + 2 * 4
    assert C.sizeof(_abi.ConvertJob) == 4 * 8 + 6 * 4


def test_null_context_is_rejected_without_a_gpu():
    lib = _abi.load()
    assert lib.b2_sync(None, 0) == _abi.B2_ERR_NULL
    assert lib.b2_rgba_to_yuv420a(None, 0, None, 0, 0, 0, None) == _abi.B2_ERR_NULL
    assert lib.b2_ctx_create(0, 0, C.byref(C.c_void_p())) == _abi.B2_ERR_INVALID_CONFIG


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "go_webp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".go")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "liboracle" not in txt, f


def test_partition_contiguous_and_lpt():
    for n, ws in [(256, 8), (1024, 8), (10, 4), (3, 8), (0, 2)]:
        seen = []
        for r in range(ws):
            seen += list(partition.contiguous_shard(n, ws, r))
        assert seen == list(range(n))
        sizes = [len(partition.contiguous_shard(n, ws, r)) for r in range(ws)]
        assert max(sizes) - min(sizes) <= 1
    rng = np.random.default_rng(0)
    dims = rng.integers(64, 8193, (256, 2))
    costs = [partition.import_cost_bytes(int(w), int(h)) for w, h in dims]
    parts = partition.lpt_partition(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(256))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) <= 1.05 * (sum(costs) / 8)  # LPT bound is far tighter here

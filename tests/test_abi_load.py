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
    assert lib.b2_abi_version() == 0x0105
    assert lib.b2_strerror(0) == b"ok"
    # messages are the reference's sentinels (pkg/picture/errors.go:8-21)
    assert lib.b2_strerror(_abi.B2_ERR_BAD_DIM) == b"picture has invalid width/height"
    assert lib.b2_strerror(_abi.B2_ERR_NULL) == b"a pointer parameter is nil"


_STRUCTS = {
    "b2_import_job": _abi.ImportJob, "b2_upsample_job": _abi.UpsampleJob, "b2_argb_job": _abi.ArgbJob,
    "b2_plain_job": _abi.PlainJob, "b2_alpha_index_job": _abi.AlphaIndexJob, "b2_expand_job": _abi.ExpandJob,
    "b2_convert_job": _abi.ConvertJob, "b2_picture_job": _abi.PictureJob, "b2_anim_blend_job": _abi.AnimBlendJob,
    "b2_anim_compare_job": _abi.AnimCompareJob, "b2_alpha_filter_job": _abi.AlphaFilterJob,
    "b2_alpha_dequant_job": _abi.AlphaDequantJob, "b2_rescale_job": _abi.RescaleJob,
    "b2_palette_job": _abi.PaletteJob, "b2_vp8l_chain_job": _abi.ChainJob, "b2_vp8l_chain_host": _abi.ChainHost, "b2_import_host": _abi.ImportHost, "b2_upsample_host": _abi.UpsampleHost, "b2_argb_host": _abi.ArgbHost,
}


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of EVERY struct of include/b2webp.h, from the header itself compiled
    as C99 by gcc, against the ctypes declarations the tests and the host mirror call through."""
    import subprocess
    txt = open(os.path.join(ROOT, "include", "b2webp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    structs = {}
    for m in re.finditer(r"typedef struct (b2_\w+) \{(.*?)\} (b2_\w+);", txt, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                name = re.sub(r"\[.*?\]", "", part.strip().split()[-1]).lstrip("*")
                fields.append(name)
        structs[m.group(1)] = fields
    assert sorted(structs) == sorted(_STRUCTS), "every struct of the header has a ctypes twin (and vice versa)"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b2webp.h"', 'int main(void) {']
    for name, fields in structs.items():
        lines.append(f'  printf("{name} %zu", sizeof({name}));')
        for f in fields:
            lines.append(f'  printf(" {f}=%zu", offsetof({name}, {f}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    seen = 0
    for line in out.strip().splitlines():
        parts = line.split()
        cls = _STRUCTS[parts[0]]
        assert C.sizeof(cls) == int(parts[1]), parts[0]
        c_fields = [p.split("=") for p in parts[2:]]
        assert [f for f, _ in c_fields] == [f[0] for f in cls._fields_], parts[0]
        for f, off in c_fields:
            assert getattr(cls, f).offset == int(off), (parts[0], f)
        seen += 1
    assert seen == len(_STRUCTS)


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

"""CPU-only: the C-ABI library loads and exports every symbol include/*.h declares,
and the ctypes table in morefusion_b200/_lib.py covers exactly that set."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    syms = set()
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            syms |= set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", text))
    return syms


def test_library_exports_header_symbols():
    from morefusion_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from morefusion_b200 import build
        build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    missing = [s for s in sorted(syms) if not hasattr(L, s)]
    assert not missing, missing
    assert syms == set(_lib.SIGNATURES), (syms ^ set(_lib.SIGNATURES))
    assert L.mf_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    import morefusion_b200 as m
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.functions.average_voxelization_3d(
            torch.zeros(4, 2), torch.zeros(4, 3), torch.zeros(4, dtype=torch.int32),
            batch_size=1, origin=(0, 0, 0), pitch=1.0, dimensions=(2, 2, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.functions.truncated_distance_function(
            torch.zeros(4, 3), pitch=1.0, origin=(0, 0, 0), dims=(2, 2, 2), truncation=1.0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "morefusion_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "from .. import oracle" not in text


def test_argument_validation_matches_reference():
    import torch
    import morefusion_b200 as m
    f = m.functions
    v, p, b = torch.zeros(4, 2), torch.zeros(4, 3), torch.zeros(4, dtype=torch.int32)
    with pytest.raises(ValueError, match="dimensions must be a tuple"):   # voxelization_3d.py:11-16
        f.average_voxelization_3d(v, p, b, batch_size=1, origin=(0, 0, 0), pitch=1.0,
                                  dimensions=[2, 2, 2])
    with pytest.raises(m.InvalidType):                                     # :20-32
        f.average_voxelization_3d(v.double(), p, b, batch_size=1, origin=(0, 0, 0), pitch=1.0,
                                  dimensions=(2, 2, 2))
    with pytest.raises(m.InvalidType):
        f.average_voxelization_3d(v, p, b.long(), batch_size=1, origin=(0, 0, 0), pitch=1.0,
                                  dimensions=(2, 2, 2))
    with pytest.raises(m.InvalidType):
        f.interpolate_voxel_grid(torch.zeros(1, 2, 4, 4), p, b)
    with pytest.raises(AssertionError):                                    # transform_points.py:8
        f.transform_points(torch.zeros(4, 2), torch.eye(4))
    with pytest.raises(TypeError):                                         # keyword-only args
        f.average_voxelization_3d(v, p, b, 1, (0, 0, 0), 1.0, (2, 2, 2))
    # transformation_matrix.py:6-17: shape contracts are asserts, for both call forms
    for q, t in ((torch.zeros(3, 4), torch.zeros(2, 3)), (torch.zeros(3, 5), torch.zeros(3, 3)),
                 (torch.zeros(4), torch.zeros(1, 3)), (torch.zeros(4), torch.zeros(4)),
                 (torch.zeros(2, 3, 4), torch.zeros(2, 3))):
        with pytest.raises(AssertionError):
            f.transformation_matrix(q, t)


def test_smoke_checker_imports_resolve():
    """__graft_entry__.smoke() imports tests/smoke_extra.py by path; every import statement inside
    its run() must resolve without a GPU (catches relative-import mistakes before the GPU box)."""
    import importlib
    import inspect
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    mod = importlib.import_module("smoke_extra")
    for line in inspect.getsource(mod.run).splitlines():
        stmt = line.strip()
        if stmt.startswith(("from ", "import ")):
            exec(stmt, {})


def test_header_is_plain_c_and_struct_layout_matches_ctypes(tmp_path):
    """include/morefusion_b200.h compiles as C (no C++/torch types in the boundary) and the
    GemmParams layout the Python side builds with ctypes is the one a C caller sees."""
    import ctypes
    import os
    import subprocess
    from morefusion_b200.contrib.singleview_3d.models.model import GemmParams
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fields = [f[0] for f in GemmParams._fields_]
    src = tmp_path / "layout.c"
    prints = "\n".join(f'  printf("{f} %zu\\n", offsetof(GemmParams, {f}));' for f in fields)
    src.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include "morefusion_b200.h"\n'
        "int main(void) {\n" + prints + '\n  printf("sizeof %zu\\n", sizeof(GemmParams));\n'
        "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = dict(zip(out[0::2], map(int, out[1::2])))
    for f in fields:
        assert got[f] == getattr(GemmParams, f).offset, f
    assert got["sizeof"] == ctypes.sizeof(GemmParams)

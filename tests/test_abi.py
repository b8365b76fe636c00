"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares
(no compute calls here); compute entry points fail loudly without a device."""
import ctypes as C
import glob
import os
import re

import pytest

from bevy_terrain_b200 import _ffi as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(bt_[a-z0-9_]+)\s*\(", text))
    return names


def test_every_declared_symbol_is_exported():
    lib = F.load_library()
    names = declared_symbols()
    assert len(names) > 50
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_abi_version_and_struct_sizes():
    lib = F.load_library()
    assert lib.bt_abi_version() == 1
    # encase std140/std430 sizes of the reference structs
    assert C.sizeof(F.Tile) == 16 and C.sizeof(F.Indirect) == 16
    assert C.sizeof(F.ViewConfigUniform) == 52
    assert C.sizeof(F.SideParameter) == 112 and C.sizeof(F.ModelApproximation) == 688
    assert C.sizeof(F.CullingUniform) == 160
    assert C.sizeof(F.Vertex) == 40
    assert F.ModelApproximation.sides.offset == 16 and F.SideParameter.c.offset == 16


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the context cannot be created: the product never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = F.load_library()
    h = C.c_void_p()
    assert lib.bt_context_create(0, C.byref(h)) == F.BT_ERR_CUDA
    assert not h.value


def test_product_does_not_reference_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may touch oracle/"""
    pkg = os.path.join(ROOT, "bevy_terrain_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"bt_oracle|libbt_oracle|from tests|import tests|oracle/", text) and f != "_ffi.py":
                    # comments that merely mention "the oracle" are fine; linking/importing is not
                    if re.search(r"#include\s+\".*oracle|import .*oracle|CDLL\(.*oracle|-lbt_oracle", text):
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders

"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares
(no compute calls here); compute entry points fail loudly without a device."""
import ctypes as C
import glob
import os
import re
import sys

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


# ------------------------------------------------------------------------------------------------
# Rust shim (rust/bevy_terrain_b200): source-only here (no cargo in the image), so what CAN be checked is checked as text:
# the extern block is the generator's output for the current header, the wrappers only name things that exist, and the
# #[repr(C)] structs have the byte size of the C structs (as laid out by the ctypes mirror, which the GPU tests exercise).
# ------------------------------------------------------------------------------------------------
RUST = os.path.join(ROOT, "rust", "bevy_terrain_b200", "src")


def test_rust_extern_block_is_current():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_rust_ffi
    with open(os.path.join(RUST, "ffi_extern.rs")) as f:
        committed = f.read()
    assert committed == gen_rust_ffi.generate(), "run `python tools/gen_rust_ffi.py --write` after changing the header"
    names = set(re.findall(r"pub fn (bt_\w+)\(", committed))
    assert names == set(declared_symbols())


def test_rust_wrappers_name_only_what_the_ffi_module_declares():
    with open(os.path.join(RUST, "ffi.rs")) as f:
        ffi = f.read()
    with open(os.path.join(RUST, "ffi_extern.rs")) as f:
        ffi += f.read()
    declared = set(re.findall(r"pub (?:fn|struct|const|type) (\w+)", ffi))
    declared |= set(n.strip() for m in re.findall(r"opaque!\(([^)]*)\)", ffi) for n in m.split(","))
    with open(os.path.join(RUST, "lib.rs")) as f:
        used = set(re.findall(r"ffi::(\w+)", f.read()))
    assert used - declared == set()


_RUST_SIZES = {"u8": 1, "u16": 2, "u32": 4, "i32": 4, "f32": 4, "u64": 8, "f64": 8, "usize": 8, "bt_status": 4}


def _rust_struct_sizes():
    """size and alignment of every plain #[repr(C)] struct of ffi.rs, computed from its field list with C layout rules"""
    with open(os.path.join(RUST, "ffi.rs")) as f:
        text = f.read()
    structs = {}
    for name, body in re.findall(r"pub struct (\w+) \{([^}]*)\}", text):
        fields = [f.strip() for f in re.split(r",\s*pub ", re.sub(r"/[/*][^\n]*", "", body))]
        structs[name] = [re.sub(r"^pub ", "", f).split(":", 1)[1].strip().rstrip(",").strip() for f in fields if ":" in f]
    done = {}

    def layout(ty):
        if ty.startswith("*"):
            return 8, 8
        m = re.match(r"\[(.+); (\d+)\]$", ty)
        if m:
            s, a = layout(m.group(1).strip())
            return s * int(m.group(2)), a
        if ty in _RUST_SIZES:
            return _RUST_SIZES[ty], _RUST_SIZES[ty]
        if ty not in done:
            off, align = 0, 1
            for f in structs[ty]:
                s, a = layout(f)
                off = (off + a - 1) // a * a + s
                align = max(align, a)
            done[ty] = ((off + align - 1) // align * align, align)
        return done[ty]

    return {n: layout(n)[0] for n in structs if n != "_p"}


def test_rust_struct_sizes_match_the_c_structs():
    import ctypes
    from bevy_terrain_b200 import _ffi
    sizes = _rust_struct_sizes()
    checked = 0
    for name, size in sizes.items():
        mirror = getattr(_ffi, "".join(w.capitalize() for w in name.split("_")[1:]), None)   # bt_view_frame -> ViewFrame
        if mirror is None:
            continue
        assert ctypes.sizeof(mirror) == size, name
        checked += 1
    assert checked >= 20

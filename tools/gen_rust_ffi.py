"""Generates the `extern "C"` block of rust/bevy_terrain_b200/src/ffi.rs from include/bevy_terrain_b200.h so that the Rust
shim lists every entry point of the C ABI.  (The shim is source-only here: there is no Rust toolchain in the image.)

    python tools/gen_rust_ffi.py > /tmp/extern_block.rs
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALARS = {"uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "int": "i32", "size_t": "usize", "float": "f32", "double": "f64",
           "bt_status": "bt_status", "char": "c_char", "void": "c_void", "uint16_t": "u16", "uint8_t": "u8"}


def rust_type(ctype: str, array: bool) -> str:
    ctype = re.sub(r"\*\s*const\b", "*", ctype.strip())     # `T* const*`: the inner const has no Rust counterpart in a raw pointer chain
    const = ctype.startswith("const ")
    base = ctype.replace("const ", "").strip()
    stars = base.count("*")
    base = base.replace("*", "").strip()
    r = SCALARS.get(base, base)
    if array:
        stars += 1
    for i in range(stars):
        r = ("*const " if (const and i == 0) else "*mut ") + r
    return r


def generate() -> str:
    import io
    buf = io.StringIO()
    old = sys.stdout
    sys.stdout = buf
    try:
        emit()
    finally:
        sys.stdout = old
    return buf.getvalue()


def main():
    text = generate()
    if "--write" in sys.argv:
        open(os.path.join(ROOT, "rust", "bevy_terrain_b200", "src", "ffi_extern.rs"), "w").write(text)
    else:
        sys.stdout.write(text)


def emit():
    src = open(os.path.join(ROOT, "include", "bevy_terrain_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"^\s*((?:const\s+)?[A-Za-z_][\w]*\s*\**)\s*(bt_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.M)
    print("// GENERATED from include/bevy_terrain_b200.h by tools/gen_rust_ffi.py — do not edit")
    print('#[link(name = "bevy_terrain_b200")]\nextern "C" {')
    for ret, name, params in protos:
        args = []
        params = " ".join(params.split())
        if params and params != "void":
            for p in params.split(","):
                p = p.strip()
                m = re.match(r"(.*?)([A-Za-z_]\w*)\s*((?:\[[^\]]*\])*)$", p)
                ctype, pname, arr = m.group(1), m.group(2), m.group(3)
                pname = {"type": "type_", "ref": "ref_", "box": "box_", "in": "in_"}.get(pname, pname)
                args.append(f"{pname}: {rust_type(ctype, bool(arr))}")
        ret = ret.strip()
        rret = "" if ret == "void" else f" -> {rust_type(ret, False)}"
        print(f"    pub fn {name}({', '.join(args)}){rret};")
    print("}")
    print(f"// {len(protos)} entry points")


if __name__ == "__main__":
    main()

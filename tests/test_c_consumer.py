"""include/bevy_terrain_b200.h must be valid C (the reference-side binding is cgo/bindgen-style FFI, not C++): a plain-C program
(examples/c/host_only.c) is compiled with -std=c11 -Wall -Wextra -Werror against the header, linked with the library and run —
it exercises the host-only entry points (model, clipmap, Taylor approximation, sharded plan, image decoder) without a GPU."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_plain_c_program_compiles_links_and_runs(tmp_path):
    from tests.test_image_io import write_tiff
    lib_dir = os.path.join(ROOT, "bevy_terrain_b200", "csrc")
    exe = str(tmp_path / "host_only")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    build = subprocess.run([cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                            os.path.join(ROOT, "examples", "c", "host_only.c"), "-L", lib_dir, "-lbevy_terrain_b200", "-o", exe],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    img = (np.arange(48 * 40, dtype=np.uint32).reshape(48, 40) * 7 % 65536).astype(np.uint16)
    write_tiff(tmp_path / "height.tif", img, compression=5, predictor=2, rows_per_strip=16)
    env = dict(os.environ, LD_LIBRARY_PATH=lib_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    run = subprocess.run([exe, str(tmp_path / "height.tif")], capture_output=True, text=True, env=env, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    out = run.stdout
    assert "ABI version 1" in out and "clipmap:" in out and "sharded cube plan (rank 1 of 4)" in out
    assert "height.tif: 40 x 48, R16" in out
    assert "bt_context_create ->" in out

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


def _sum32(a: np.ndarray) -> int:
    w = np.frombuffer(a.tobytes(), dtype=np.uint32).astype(np.uint64)
    k = (np.arange(len(w), dtype=np.uint64) % np.uint64(251)) + np.uint64(1)
    return int((w * k).sum(dtype=np.uint64))


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_program_runs_the_whole_path_and_matches_the_python_mirror(ctx, tmp_path):
    """examples/c/frame.c drives source -> pyramid -> clipmap (device or host) -> atlas -> bt_view_run_frame through the C ABI from plain C;
    the same frame through the ctypes mirror (itself held to the oracle by the parity tests) must give the same tile list and vertices."""
    from bevy_terrain_b200 import _ffi as F, api
    lib_dir = os.path.join(ROOT, "bevy_terrain_b200", "csrc")
    exe = str(tmp_path / "frame")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    build = subprocess.run([cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                            os.path.join(ROOT, "examples", "c", "frame.c"), "-L", lib_dir, "-lbevy_terrain_b200", "-lm", "-o", exe],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    env = dict(os.environ, LD_LIBRARY_PATH=lib_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    outs = []
    for device_clipmap in ("1", "0"):
        run = subprocess.run([exe, device_clipmap], capture_output=True, text=True, env=env, timeout=300)
        assert run.returncode == 0, run.stdout + run.stderr
        outs.append({l.split(":")[0] if ":" in l and not l.startswith(("tile", "vertex")) else " ".join(l.split()[:-1]): l for l in run.stdout.strip().splitlines()})
    # device clipmap and host clipmap: same loads, table, tile list, vertices
    assert outs[0]["clipmap (device)"].split("):")[1] == outs[1]["clipmap (host)"].split("):")[1]
    assert outs[0]["frame"] == outs[1]["frame"] and outs[0]["tile list sum"] == outs[1]["tile list sum"] and outs[0]["vertex sum"] == outs[1]["vertex sum"]
    assert outs[0]["pyramid"] == outs[1]["pyramid"]

    # the same frame through the Python mirror
    SRC, LODS, T, B, ATLAS, GRID = 512, 4, 64, 2, 256, 16
    y, x = np.mgrid[0:SRC, 0:SRC].astype(np.uint64)
    source = (((x * 37 + y * 91 + ((x * y) >> 3)) % 60000) + 1000).astype(np.uint16)
    model = api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
    cfg = api.TerrainConfig(lod_count=LODS, model=model, atlas_size=ATLAS)
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=T, border_size=B, format=api.AttachmentFormat.R16))
    pre = api.Preprocessor(ctx, cfg)
    pre.preprocess_tile(api.PreprocessDataset(pixels=source, lod_range=range(0, LODS))).run()
    vcfg = api.TerrainViewConfig(grid_size=GRID)
    vcfg.geometry_tile_count = 1 << 16
    atlas = api.TileAtlas(ctx, cfg)
    tree = api.TileTree(atlas, vcfg, max_vertices=8192 * 2 * GRID * (GRID + 2))
    try:
        assert f"{pre.content_hash(0):016x}" in outs[0]["pyramid"]
        tree.add_existing_tiles(pre.tiles())
        pos = (100.0, 170.0, -50.0)
        api.load_resident_tiles(tree, pre, pos)
        tree.set_approximate_height(250.0)
        tiles, ind, _ = tree.frame_step(F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS)
        assert outs[0]["frame"] == f"frame: {len(tiles)} tiles, {ind.vertex_count} vertices"
        assert int(outs[0]["tile list sum"].split()[-1]) == _sum32(np.ascontiguousarray(tiles))
        verts = tree.download_vertices(0, ind.vertex_count)
        assert int(outs[0]["vertex sum"].split()[-1]) == _sum32(verts)
    finally:
        tree.close(); atlas.close(); pre.close()

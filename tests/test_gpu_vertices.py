"""GPU parity: bt_gen_vertices vs the oracle's restatement of vertex.wgsl (+ sample_normal).

The kernels are compiled without FMA contraction and use the oracle's deterministic log2, so the
bar here is BIT-EXACT equality of every vertex record (stronger than the 1e-5 relative tolerance
north_star allows); a tolerance check is kept as a second, looser assertion for diagnostics."""
import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests.scene import Scene, planar_view_positions, spherical_view_positions

pytestmark = pytest.mark.gpu

ALL = F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS


def _gpu_frame(ctx, scene, positions, flags, max_vertices=0):
    from bevy_terrain_b200 import api
    atlas = api.TileAtlas(ctx, scene.terrain_config())
    tree = api.TileTree(atlas, scene.cfg, max_vertices=max_vertices)
    out = []
    try:
        for pos in positions:
            loads = scene.move_view(pos)
            for tc, idx in loads:
                atlas.upload_tile(0, idx, scene.atlas[idx])
            tree.prepare(frame=scene.view_frame())
            tree.select_tiles()
            tiles, ind = tree.readback()
            tree.gen_vertices(flags)
            verts = tree.download_vertices(0, ind.vertex_count)
            out.append((tiles, verts))
    finally:
        tree.close()
        atlas.close()
    return out


def _assert_vertices_equal(got, exp):
    assert got.shape == exp.shape
    assert np.array_equal(got["tile_index"], exp["tile_index"])
    for f in ("position", "normal", "uv"):
        g, e = got[f], exp[f]
        same = (g.view(np.uint32) == e.view(np.uint32)) | (np.isnan(g) & np.isnan(e))
        if not same.all():
            bad = np.argwhere(~same)
            i = bad[0][0]
            rel = np.abs(g - e) / np.maximum(np.abs(e), 1e-30)
            raise AssertionError(f"{f}: {len(bad)} of {g.size} components differ bitwise (max rel {np.nanmax(rel):.3e}); "
                                 f"first at vertex {i}: got {g[i]} expected {e[i]}")


@pytest.mark.parametrize("flags", [0, F.BT_FLAG_NORMALS, F.BT_FLAG_MORPH | F.BT_FLAG_BLEND, ALL])
def test_vertices_planar_bit_exact(ctx, flags):
    scene = Scene("planar", lod_count=4, T=64, B=2, src=256, grid_size=16)
    positions = planar_view_positions(4)
    for pos, (tiles, verts) in zip(positions, _gpu_frame(ctx, scene, positions, flags)):
        scene.move_view(pos)
        exp = scene.oracle_vertices(tiles, flags)
        _assert_vertices_equal(verts, exp)


def test_vertices_planar_grid64_bit_exact(ctx):
    """BASELINE config 2 shape (grid 64, bilinear, normals on) at a reduced texture size"""
    scene = Scene("planar", lod_count=5, T=64, B=2, src=512, grid_size=64)
    pos = (100.0, 170.0, -50.0)
    (tiles, verts), = _gpu_frame(ctx, scene, [pos], F.BT_FLAG_NORMALS)
    scene.move_view(pos)
    n = len(tiles) * scene.vc.vertices_per_tile
    assert len(verts) == n
    # compare a bounded sample of whole tiles (the oracle is scalar C)
    rng = np.random.default_rng(3)
    pick = np.sort(rng.choice(len(tiles), size=min(64, len(tiles)), replace=False))
    vpt = scene.vc.vertices_per_tile
    vi = scene.oracle_vertex_inputs(tiles, F.BT_FLAG_NORMALS)
    from tests import oracle as O
    for t in pick:
        exp = O.gen_vertices(vi, int(t) * vpt, vpt)
        _assert_vertices_equal(verts[t * vpt:(t + 1) * vpt], exp)


@pytest.mark.parametrize("flags", [ALL, ALL | F.BT_FLAG_HIGH_PRECISION, 0, F.BT_FLAG_MORPH | F.BT_FLAG_BLEND])   # without NORMALS the record carries the MODEL normal (signed zeros of the cube-face coordinates included)
def test_vertices_spherical_bit_exact(ctx, flags):
    scene = Scene("spherical", lod_count=3, T=32, B=2, src=64, grid_size=16)
    positions = spherical_view_positions()[:4]
    for pos, (tiles, verts) in zip(positions, _gpu_frame(ctx, scene, positions, flags)):
        scene.move_view(pos)
        exp = scene.oracle_vertices(tiles, flags)
        _assert_vertices_equal(verts, exp)


def test_vertices_ellipsoid_bit_exact(ctx):
    """TerrainModel::ellipsoid: a "spherical" terrain whose mesh transform has the scale (major, minor, major); the host-side model
    (ellipsoid projection of the view, Taylor coefficients, clipmap) comes from the product's C++ through the ABI"""
    from bevy_terrain_b200 import api
    scene = Scene("ellipsoid", lod_count=3, T=32, B=2, src=64, grid_size=16)
    a = 6378137.0
    d = np.array([0.3, 0.9, -0.4]); d /= np.linalg.norm(d)
    positions = [tuple(d * (a + 400000.0)), tuple(np.array([-1.0, 0.01, 0.02]) * (a + 30000.0))]
    flags = ALL | F.BT_FLAG_HIGH_PRECISION
    atlas = api.TileAtlas(ctx, scene.terrain_config())
    tree = api.TileTree(atlas, scene.cfg)
    try:
        tree.add_existing_tiles(np.array(sorted(scene.existing), dtype=np.uint32))
        for pos in positions:
            loads = scene.move_view(pos)
            got_loads = tree.compute_requests(pos)            # the product's clipmap (ellipsoid model) against the reference twin
            assert [(tuple(int(v) for v in t), i) for t, i in got_loads] == loads
            for tc, idx in loads:
                atlas.upload_tile(0, idx, scene.atlas[idx])
                tree.mark_loaded(tc)
            tree.adjust_to_tile_atlas()
            tree.set_approximate_height(scene.approximate_height)
            tree.prepare()                                     # product-side model approximation + table
            assert bytes(tree.frame.approximation) == bytes(scene.approx)
            tree.select_tiles()
            tiles, ind = tree.readback()
            exp_tiles, _, _ = scene.oracle_select()
            assert np.array_equal(tiles, exp_tiles)
            tree.gen_vertices(flags)
            _assert_vertices_equal(tree.download_vertices(0, ind.vertex_count), scene.oracle_vertices(tiles, flags))
    finally:
        tree.close(); atlas.close()


def test_vertices_high_precision_path_is_taken(ctx):
    """close to the surface the Taylor path (compute_relative_position) must engage and still match"""
    scene = Scene("spherical", lod_count=3, T=32, B=2, src=64, grid_size=16, precision_threshold_distance=0.01)
    pos = (-(6371000.0 - 1500.0 + 2000.0), 100.0, -300.0)
    flags = ALL | F.BT_FLAG_HIGH_PRECISION
    (tiles, verts), = _gpu_frame(ctx, scene, [pos], flags)
    scene.move_view(pos)
    exp = scene.oracle_vertices(tiles, flags)
    exp_lo = scene.oracle_vertices(tiles, ALL)
    assert not np.array_equal(exp["position"], exp_lo["position"]), "scenario does not exercise HIGH_PRECISION"
    _assert_vertices_equal(verts, exp)


def test_vertices_structure_properties(ctx):
    """size-independent properties: duplicated strip indices, shared edges, crack-free morph"""
    scene = Scene("planar", lod_count=4, T=64, B=2, src=256, grid_size=16)
    pos = (100.0, 170.0, -50.0)
    (tiles, verts), = _gpu_frame(ctx, scene, [pos], ALL)
    vpt, vpr, g = scene.vc.vertices_per_tile, scene.vc.vertices_per_row, 16
    v = verts.reshape(len(tiles), g, vpr)
    # first/last index of every column strip is duplicated (degenerate triangles)
    raw = np.ascontiguousarray(v).view(np.uint8).reshape(len(tiles), g, vpr, 40)
    assert np.array_equal(raw[:, :, 0], raw[:, :, 1])
    assert np.array_equal(raw[:, :, -1], raw[:, :, -2])
    # column c's right vertices are column c+1's left vertices
    right = v[:, :-1, 2::2][:, :, :g + 1]
    left = v[:, 1:, 1::2][:, :, :g + 1]
    assert np.array_equal(right["position"], left["position"])
    # tile_index identifies the tile of every vertex
    assert np.array_equal(verts["tile_index"], np.repeat(np.arange(len(tiles), dtype=np.uint32), vpt))


def test_vertex_buffer_overflow_is_reported(ctx):
    scene = Scene("planar", lod_count=4, T=64, B=2, src=256, grid_size=16)
    with pytest.raises(F.BtError) as e:
        _gpu_frame(ctx, scene, [(100.0, 170.0, -50.0)], 0, max_vertices=576 * 10)
    assert e.value.status in (F.BT_ERR_CAPACITY, F.BT_ERR_INVALID_ARG)


def test_run_frame_equals_separate_calls(ctx):
    """bt_view_run_frame (upload + select + gen + overlapped tile-list read-back) == the four separate entry points"""
    from bevy_terrain_b200 import api
    scene = Scene("planar", lod_count=4, T=64, B=2, src=256, grid_size=16)
    atlas = api.TileAtlas(ctx, scene.terrain_config())
    tree = api.TileTree(atlas, scene.cfg, max_vertices=9000 * 576)
    try:
        tree.add_existing_tiles(np.array(sorted(scene.existing), dtype=np.uint32))
        for pos in planar_view_positions(4):
            loads = scene.move_view(pos)
            for tc, idx in loads:
                atlas.upload_tile(0, idx, scene.atlas[idx])
            # drive the product's own clipmap so that build_frame() has its table
            for tile, atlas_index in tree.compute_requests(pos):
                tree.mark_loaded(tile)
            tree.adjust_to_tile_atlas()
            tiles_f, ind_f, vb = tree.frame_step(ALL, tile_capacity=9000)
            verts_f = tree.download_vertices(0, ind_f.vertex_count)
            tree.prepare()
            tree.select_tiles()
            tiles_s, ind_s = tree.readback()
            tree.gen_vertices(ALL)
            verts_s = tree.download_vertices(0, ind_s.vertex_count)
            assert np.array_equal(tiles_f, tiles_s) and ind_f.vertex_count == ind_s.vertex_count
            assert verts_f.tobytes() == verts_s.tobytes()
            exp_tiles, _, _ = scene.oracle_select()
            assert np.array_equal(tiles_f, exp_tiles)
    finally:
        tree.close(); atlas.close()


def test_vertices_before_any_tile_is_loaded(ctx):
    """The first frames of the streaming loop: no tile has been marked loaded yet, so every tile-tree entry is INVALID
    ({0xFFFFFFFF, 0xFFFFFFFF}, tile_tree.rs:51-65).  The reference keeps rendering (WGSL clamps the texture-array index to the
    last layer); the kernel must do the same — no fault, same bits as the oracle — and the context must stay usable."""
    from bevy_terrain_b200 import api
    scene = Scene("planar", lod_count=4, T=64, B=2, src=256, grid_size=16)
    pos = planar_view_positions(1)[0]
    scene.move_view(pos)
    atlas = api.TileAtlas(ctx, scene.terrain_config())
    tree = api.TileTree(atlas, scene.cfg)
    try:
        # upload real texels to the LAST layer so the clamped read is visible in the result, but leave the table INVALID
        last = scene.atlas_size - 1
        rng = np.random.default_rng(9)
        texels = rng.integers(1, 65535, size=(scene.T, scene.T), dtype=np.uint16)
        atlas.upload_tile(0, last, texels)
        frame = scene.view_frame()
        invalid = np.full((len(scene.entries), 2), 0xFFFFFFFF, dtype=np.uint32)
        import ctypes as C
        frame.tile_tree = invalid.ctypes.data_as(C.POINTER(F.TileTreeEntry))
        tree.prepare(frame=frame)
        tree.select_tiles()
        tiles, ind = tree.readback()
        for flags in (F.BT_FLAG_NORMALS, ALL):
            tree.gen_vertices(flags)
            verts = tree.download_vertices(0, ind.vertex_count)
            host = scene.atlas.copy()
            saved_atlas, saved_entries = scene.atlas, scene.entries
            try:
                host[:] = 0
                host[last] = texels
                scene.atlas, scene.entries = host, invalid
                exp = scene.oracle_vertices(tiles, flags)
            finally:
                scene.atlas, scene.entries = saved_atlas, saved_entries
            _assert_vertices_equal(verts, exp)
        ctx.synchronize()     # no sticky CUDA error
    finally:
        tree.close()
        atlas.close()

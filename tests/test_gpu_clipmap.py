"""Device clipmap (bt_clipmap) against the host TileTree (bt_tile_tree_*, itself pinned to the Python restatement of
tile_tree.rs / tile_atlas.rs by tests/test_host_math.py): along a camera path, with some tiles missing on disk and some loads
left unfinished, the load requests (tile AND atlas index, in order), the origins and the best-tile table must be identical;
then one frame selected and generated from the device-resident table must equal the frame from the uploaded table."""
import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O
from tests.scene import planar_view_positions, spherical_view_positions

pytestmark = pytest.mark.gpu


def _model(api, which):
    if which == 0:
        return api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0), planar_view_positions(8)
    if which == 1:
        return api.TerrainModel.sphere((0.0, 0.0, 0.0), 6371000.0, -12000.0, 9000.0), spherical_view_positions()
    return api.TerrainModel.ellipsoid((0.0, 0.0, 0.0), 6378137.0, 6356752.314245, -12000.0, 9000.0), spherical_view_positions()


@pytest.mark.parametrize("which", [0, 1, 2])
def test_device_clipmap_equals_host_tile_tree(ctx, which):
    from bevy_terrain_b200 import api
    model, positions = _model(api, which)
    lod_count, atlas_size = 5, 400
    sides = 6 if which else 1
    existing = np.array([(s, l, x, y) for s in range(sides) for l in range(lod_count) for x in range(1 << l) for y in range(1 << l)
                         if not (l == 4 and (x + y) % 7 == 0)], dtype=np.uint32)
    vcfg = O.default_view_config()
    vcfg.geometry_tile_count = 1 << 17
    cfg = api.TerrainConfig(lod_count=lod_count, model=model, atlas_size=atlas_size)
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=32, border_size=2, format=api.AttachmentFormat.R16))
    atlases = [api.TileAtlas(ctx, cfg), api.TileAtlas(ctx, cfg)]
    host = api.TileTree(atlases[0], vcfg)
    dev = api.TileTree(atlases[1], vcfg, device_clipmap=True)
    try:
        host.add_existing_tiles(existing)
        dev.add_existing_tiles(existing)
        total_loads = 0
        for i, pos in enumerate(positions + positions[::-1]):
            exp = host.compute_requests(pos)
            got = dev.compute_requests(pos)
            assert got == exp, f"frame {i}: load requests differ"
            total_loads += len(exp)
            for j, (tc, _) in enumerate(exp):
                if j % 5 != 4:   # some tiles stay "loading": the table must fall back to their parents
                    host.mark_loaded(tc)
                    dev.mark_loaded(tc)
            host.adjust_to_tile_atlas()
            dev.adjust_to_tile_atlas()
            he, ho = host.download_table()
            de, do = dev.download_table()
            assert np.array_equal(do, ho), f"frame {i}: origins differ"
            assert np.array_equal(de, he), f"frame {i}: best-tile tables differ"
            h2d, d2h = dev.clipmap_transfer_bytes()
            assert h2d in (0, 16 * 2048)   # the loaded-tile set goes up only when it changed (hash capacity 4 x atlas_size -> 2048 slots)
        assert total_loads > 50
        # a frame that reads the device-resident table (no table upload) equals a frame from the uploaded table
        for t in (host, dev):
            t.set_approximate_height(250.0 if which == 0 else -1500.0)
            t.prepare()
            t.select_tiles()
        th, ih = host.readback()
        td, idv = dev.readback()
        assert np.array_equal(th, td) and ih.vertex_count == idv.vertex_count and len(th) > 0
        host.gen_vertices(F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS)
        dev.gen_vertices(F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS)
        n = len(th) * int(vcfg.grid_size) * 2 * (int(vcfg.grid_size) + 2)
        assert host.download_vertices(0, n).tobytes() == dev.download_vertices(0, n).tobytes()
    finally:
        host.close()
        dev.close()
        for a in atlases:
            a.close()


def test_clipmap_rejects_a_tree_already_updated_on_the_host(ctx):
    from bevy_terrain_b200 import api
    import ctypes as C
    model, positions = _model(api, 0)
    cfg = api.TerrainConfig(lod_count=3, model=model, atlas_size=64)
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=32, border_size=2, format=api.AttachmentFormat.R16))
    atlas = api.TileAtlas(ctx, cfg)
    tree = api.TileTree(atlas, O.default_view_config())
    try:
        tree.add_existing_tiles(np.array([(0, 0, 0, 0)], dtype=np.uint32))
        tree.compute_requests(positions[0])
        clip = C.c_void_p()
        assert ctx.lib.bt_clipmap_create(tree._view, tree._tree, C.byref(clip)) == F.BT_ERR_INVALID_ARG
        # and without a clipmap the view still insists on a host table
        f = tree.build_frame()
        f.tile_tree = None
        assert ctx.lib.bt_view_update(tree._view, C.byref(f), None) == F.BT_ERR_INVALID_ARG
    finally:
        tree.close()
        atlas.close()

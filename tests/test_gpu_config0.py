"""BASELINE.json configs[0] at full size — the reference's own CPU-runnable case (examples/minimal.rs shape): planar 1024^2
heightmap, 4 LODs, texture 512, border 2.  Small enough for the scalar oracle to recompute EVERYTHING: every pyramid tile,
the whole tile list (in order) and every generated vertex are compared bit for bit."""
import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests.scene import Scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("grid_size", [16, 64])
def test_config0_everything_bit_exact(ctx, grid_size):
    from bevy_terrain_b200 import api
    scene = Scene("planar", lod_count=4, T=512, B=2, src=1024, grid_size=grid_size)
    cfg = scene.terrain_config()
    pre = api.Preprocessor(ctx, cfg)
    pre.preprocess_tile(api.PreprocessDataset(pixels=scene.sources[0], lod_range=range(0, 4))).run()
    tiles = pre.tiles()
    assert len(tiles) == 85 == len(scene.existing)
    for tc in tiles:
        assert np.array_equal(pre.get_tile(0, tc), scene.pyramid.tile(*[int(v) for v in tc])), f"pyramid tile {tc} differs"
    atlas = api.TileAtlas(ctx, cfg)
    tree = api.TileTree(atlas, scene.cfg)
    try:
        tree.add_existing_tiles(tiles)
        for pos, flags in (((100.0, 170.0, -50.0), F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS),
                           ((-700.0, 400.0, 650.0), F.BT_FLAG_NORMALS)):
            scene.move_view(pos)
            api.load_resident_tiles(tree, pre, pos)
            tree.prepare()
            tree.select_tiles()
            got_tiles, ind = tree.readback()
            exp_tiles, exp_ind, _ = scene.oracle_select()
            assert np.array_equal(got_tiles, exp_tiles)
            assert ind.vertex_count == exp_ind.vertex_count
            tree.gen_vertices(flags)
            verts = tree.download_vertices(0, ind.vertex_count)
            exp = scene.oracle_vertices(got_tiles, flags)
            assert verts.tobytes() == exp.tobytes()
    finally:
        tree.close(); atlas.close(); pre.close()

"""The CUDA path reproduces the committed golden fixtures through the C ABI (no oracle involved at run time)."""
import numpy as np
import pytest

from tests.golden_util import GoldenFrame, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,spherical", [("frame_planar.npz", False), ("frame_spherical.npz", True)])
def test_cuda_reproduces_frame_goldens(ctx, name, spherical):
    from bevy_terrain_b200 import api
    g = load(name)
    fr0 = GoldenFrame(g, 0, spherical)
    model = api.TerrainModel.sphere((0, 0, 0), 6371000.0, -12000.0, 9000.0) if spherical else api.TerrainModel.planar((0, -100, 0), 2000.0, 0.0, 500.0)
    cfg = api.TerrainConfig(lod_count=fr0.lod_count, model=model, atlas_size=len(fr0.atlas))
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=fr0.T, border_size=fr0.B, format=api.AttachmentFormat.R16))
    atlas = api.TileAtlas(ctx, cfg)
    tree = api.TileTree(atlas, api.TerrainViewConfig(grid_size=16))
    try:
        for i in range(int(g["n_views"])):
            fr = GoldenFrame(g, i, spherical)
            for idx in range(len(fr.atlas)):
                if fr.atlas[idx].any():
                    atlas.upload_tile(0, idx, fr.atlas[idx])
            tree.prepare(frame=fr.view_frame())
            tree.select_tiles()
            tiles, ind = tree.readback()
            assert np.array_equal(tiles, fr.tiles)
            assert [ind.vertex_count, ind.instance_count, ind.base_vertex, ind.base_instance] == fr.indirect.tolist()
            tree.gen_vertices(fr.flags)
            verts = tree.download_vertices(0, ind.vertex_count)
            assert verts[fr.vertex_index.astype(np.int64)].tobytes() == fr.vertices.tobytes()
    finally:
        tree.close(); atlas.close()


def test_cuda_reproduces_preprocess_goldens(ctx):
    from bevy_terrain_b200 import api
    g = load("preprocess.npz")

    def cfg_for(kind_sph, fmt, L):
        model = api.TerrainModel.sphere((0, 0, 0), 1.0, 0.0, 1.0) if kind_sph else api.TerrainModel.planar((0, 0, 0), 1.0, 0.0, 1.0)
        c = api.TerrainConfig(lod_count=L, model=model)
        c.add_attachment(api.AttachmentConfig(name="a", texture_size=16, border_size=2, format=fmt))
        return c

    for fmt, tag in ((api.AttachmentFormat.R16, "r16"), (api.AttachmentFormat.Rgba8, "rgba8")):
        pre = api.Preprocessor(ctx, cfg_for(False, fmt, 3))
        try:
            pre.preprocess_tile(api.PreprocessDataset(pixels=np.ascontiguousarray(g[f"{tag}_source"]), lod_range=range(0, 3))).run()
            for t, d in zip(g[f"{tag}_tiles"], g[f"{tag}_data"]):
                assert np.array_equal(pre.get_tile(0, t), d)
        finally:
            pre.close()
    pre = api.Preprocessor(ctx, cfg_for(True, api.AttachmentFormat.R16, 2))
    try:
        pre.preprocess_spherical(api.SphericalDataset(faces=[np.ascontiguousarray(f) for f in g["sph_sources"]], lod_range=range(0, 2))).run()
        for t, d in zip(g["sph_tiles"], g["sph_data"]):
            assert np.array_equal(pre.get_tile(0, t), d)
    finally:
        pre.close()

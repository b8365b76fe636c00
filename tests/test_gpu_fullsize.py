"""Size-independent properties at BASELINE.json's full config-2 size (4096^2 source, 8 LODs, texture 512, grid 64),
where the scalar oracle is too slow to recompute everything: stitched borders == neighbour centres, downsample
consistency, vertex-stream structure, and a sampled bit-exact comparison with the oracle on whole tiles."""
import ctypes as C

import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O

pytestmark = pytest.mark.gpu
T, B, L, S = 512, 2, 8, 4096
Cn = T - 2 * B


@pytest.fixture(scope="module")
def pyramid(ctx):
    from bevy_terrain_b200 import api
    model = api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
    cfg = api.TerrainConfig(lod_count=L, model=model, atlas_size=8 * 64)
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=T, border_size=B, mip_level_count=4, format=api.AttachmentFormat.R16))
    src = ctx.synth_source(F.BT_FORMAT_R16, S, S, seed=0x5EED0001)
    pre = api.Preprocessor(ctx, cfg)
    pre.preprocess_tile(api.PreprocessDataset(device_pixels=src, width=S, height=S, lod_range=range(0, L))).run()
    yield cfg, pre
    pre.close()
    ctx.free(src)


def test_full_pyramid_tile_count_and_borders(pyramid):
    cfg, pre = pyramid
    tiles = pre.tiles()
    assert len(tiles) == (4 ** L - 1) // 3 == 21845
    rng = np.random.default_rng(0)
    for lod in (3, 6, 7):
        n = 1 << lod
        for _ in range(6):
            x, y = int(rng.integers(0, n - 1)), int(rng.integers(0, n - 1))
            a = pre.get_tile(0, (0, lod, x, y))
            right, below, diag = pre.get_tile(0, (0, lod, x + 1, y)), pre.get_tile(0, (0, lod, x, y + 1)), pre.get_tile(0, (0, lod, x + 1, y + 1))
            assert np.array_equal(a[B:B + Cn, B + Cn:], right[B:B + Cn, B:2 * B])
            assert np.array_equal(a[B + Cn:, B:B + Cn], below[B:2 * B, B:B + Cn])
            assert np.array_equal(a[B + Cn:, B + Cn:], diag[B:2 * B, B:2 * B])
            assert (a != 0).all()          # the synthetic source has no holes
    edge = pre.get_tile(0, (0, 7, 0, 5))   # terrain edge: clamp-repeat of the own centre column
    assert np.array_equal(edge[B:B + Cn, :B], np.repeat(edge[B:B + Cn, B:B + 1], B, axis=1))


def test_full_pyramid_sampled_tiles_match_oracle(pyramid):
    """a restricted dataset (the source window of lod-5 tile (9,20) and its 20 descendants, texture 512) run through
    BOTH the CUDA path and the oracle: byte-exact at the full tile size"""
    cfg, pre = pyramid
    src = O.synth_source(F.BT_FORMAT_R16, S, S, seed=0x5EED0001)
    # a dataset restricted to tile (5; 9, 20) spans exactly that tile's descendants at lods 5..7
    n5 = 32
    tl, br = (9 / n5, 20 / n5), (10 / n5, 21 / n5)
    sub2 = O.Pyramid(F.BT_KIND_PLANAR, F.BT_FORMAT_R16, T, B, L)
    from bevy_terrain_b200 import api
    pre2 = api.Preprocessor(pre.ctx, cfg)
    crop = np.ascontiguousarray(src[20 * 128:21 * 128 + 1, 9 * 128:10 * 128 + 1])   # the source window of that tile (+1 texel)
    kw = dict(top_left=tl, bottom_right=br)
    pre2.preprocess_tile(api.PreprocessDataset(pixels=crop, lod_range=range(5, 8), **kw)).run()
    sub2.add_datasets([O.make_dataset(crop, top_left=tl, bottom_right=br, lod_range=(5, 8))])
    sub2.run()
    got = {tuple(int(v) for v in t) for t in pre2.tiles()}
    exp = {tuple(int(v) for v in t) for t in sub2.tiles()}
    assert got == exp and len(exp) == 21
    for tc in sorted(exp):
        assert np.array_equal(pre2.get_tile(0, tc), sub2.tile(*tc)), tc
    pre2.close()


def test_full_frame_vertices(ctx, pyramid):
    from bevy_terrain_b200 import api
    cfg, pre = pyramid
    vcfg = api.TerrainViewConfig(grid_size=64)
    atlas = api.TileAtlas(ctx, cfg)
    tree = api.TileTree(atlas, vcfg)
    try:
        tree.add_existing_tiles(pre.tiles())
        pos = (123.0, -100.0 + 250.0 + 20.0, -321.0)
        n_loads = api.load_resident_tiles(tree, pre, pos)
        assert 0 < n_loads <= 512
        tree.prepare()
        tree.select_tiles()
        tiles, ind = tree.readback()
        assert 4000 < len(tiles) < 5200 and ind.vertex_count == len(tiles) * 8448
        flags = F.BT_FLAG_NORMALS | F.BT_FLAG_MORPH | F.BT_FLAG_BLEND
        tree.gen_vertices(flags)
        # structure on a sample of tiles: duplicated strip ends, shared columns, unit normals, heights inside the range
        rng = np.random.default_rng(1)
        for t in rng.choice(len(tiles), 24, replace=False):
            v = tree.download_vertices(int(t) * 8448, 8448).reshape(64, 132)
            raw = np.ascontiguousarray(v).view(np.uint8).reshape(64, 132, 40)
            assert np.array_equal(raw[:, 0], raw[:, 1]) and np.array_equal(raw[:, -1], raw[:, -2])
            assert np.array_equal(v[:-1, 2::2]["position"][:, :65], v[1:, 1::2]["position"][:, :65])
            assert (v["tile_index"] == t).all()
            nrm = np.linalg.norm(v["normal"].astype(np.float64), axis=-1)
            # blended normals are mix(n_a, n_b, ratio) without renormalisation (fragment.wgsl:103-107): |n| <= 1
            assert nrm.max() <= 1.0 + 1e-5 and nrm.min() > 0.5
            y = v["position"][..., 1]
            assert y.min() >= -100.0 - 1e-3 and y.max() <= 400.0 + 1e-3
        # bit-exact against the oracle on whole tiles, with the product's own inputs (atlas + table downloaded)
        host_atlas = np.empty((cfg.atlas_size, T, T), dtype=np.uint16)
        ctx.d2h(host_atlas, atlas.device_ptr(0))
        vi = O.VertexInputs()
        vi.view_config, vi.approximation = tree.frame.view_config, tree.frame.approximation
        vi.terrain_config = F.TerrainConfigUniform(L, 0.0, 500.0, 1000.0)
        vi.attachments[0].size, vi.attachments[0].scale, vi.attachments[0].offset = float(Cn), float(np.float32(Cn) / np.float32(T)), float(np.float32(B) / np.float32(T))
        vi.mesh = cfg.model.mesh_transform()
        vi.view_world_position[:] = [float(np.float32(p)) for p in pos]
        vi.spherical, vi.texture_size, vi.atlas_size = 0, T, cfg.atlas_size
        entries = np.frombuffer(tree.entries, dtype=np.uint32).reshape(-1, 2).copy()
        vi.tile_tree = entries.ctypes.data_as(C.POINTER(F.TileTreeEntry))
        vi.height_atlas = host_atlas.ctypes.data
        vi.tiles = tiles.ctypes.data_as(C.POINTER(F.Tile))
        vi.tile_count, vi.flags = len(tiles), flags
        for t in rng.choice(len(tiles), 48, replace=False):
            exp = O.gen_vertices(vi, int(t) * 8448, 8448)
            got = tree.download_vertices(int(t) * 8448, 8448)
            assert got.tobytes() == exp.tobytes(), f"tile {t} differs"
    finally:
        tree.close(); atlas.close()

"""BASELINE.json configs[2] at full size: planar 16384^2 height (R16) + albedo (Rgba8), terrain with 12 LODs of which the
dataset resolves LODs 0..5 (16384 / 508 = 32 tiles per side), texture 512, grid 64, mesh morph + blend on.
(The reference samples with a linear sampler at mip 0 on this path — vertex.wgsl / attachments.wgsl use textureSampleLevel(.., 0.0);
"trilinear + aniso" in the config name refers to its fragment stage, which is out of scope.)"""
import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O
from tests.fullsize_util import assert_tiles_bit_exact, check_strip_structure, oracle_inputs

pytestmark = pytest.mark.gpu
T, B, L, S, DATA_LODS = 512, 2, 12, 16384, 6
Cn = T - 2 * B


@pytest.fixture(scope="module")
def terrain(ctx):
    from bevy_terrain_b200 import api
    model = api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
    cfg = api.TerrainConfig(lod_count=L, model=model, atlas_size=DATA_LODS * 64)
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=T, border_size=B, mip_level_count=4, format=api.AttachmentFormat.R16))
    cfg.add_attachment(api.AttachmentConfig(name="albedo", texture_size=T, border_size=B, mip_level_count=4, format=api.AttachmentFormat.Rgba8))
    srcs = [ctx.synth_source(F.BT_FORMAT_R16, S, S, seed=0x5EED0001), ctx.synth_source(F.BT_FORMAT_RGBA8, S, S, seed=0x5EED0002)]
    pre = api.Preprocessor(ctx, cfg)
    for a in (0, 1):
        pre.preprocess_tile(api.PreprocessDataset(attachment_index=a, device_pixels=srcs[a], width=S, height=S, lod_range=range(0, DATA_LODS)))
    pre.run()
    yield cfg, pre
    pre.close()
    for p in srcs:
        ctx.free(p)


def test_config2_pyramids(terrain):
    cfg, pre = terrain
    for a in (0, 1):
        assert len(pre.tiles(a)) == (4 ** DATA_LODS - 1) // 3 == 1365
    rng = np.random.default_rng(5)
    for a in (0, 1):
        for lod in (3, 5):
            n = 1 << lod
            x, y = int(rng.integers(0, n - 1)), int(rng.integers(0, n - 1))
            t, right, below = (pre.get_tile(a, (0, lod, x + dx, y + dy)) for dx, dy in ((0, 0), (1, 0), (0, 1)))
            assert np.array_equal(t[B:B + Cn, B + Cn:], right[B:B + Cn, B:2 * B])
            assert np.array_equal(t[B + Cn:, B:B + Cn], below[B:2 * B, B:B + Cn])
    # albedo: a restricted dataset (the source window of lod-4 tile (5, 11) and its descendants) through BOTH implementations
    src = O.synth_source(F.BT_FORMAT_RGBA8, S, S, seed=0x5EED0002)
    n4, w = 16, S // 16
    tl, br = (5 / n4, 11 / n4), (6 / n4, 12 / n4)
    crop = np.ascontiguousarray(src[11 * w:12 * w + 1, 5 * w:6 * w + 1])
    from bevy_terrain_b200 import api
    one = api.TerrainConfig(lod_count=L, model=cfg.model)
    one.add_attachment(api.AttachmentConfig(name="albedo", texture_size=T, border_size=B, format=api.AttachmentFormat.Rgba8))
    pre2 = api.Preprocessor(pre.ctx, one)
    pre2.preprocess_tile(api.PreprocessDataset(pixels=crop, lod_range=range(4, 6), top_left=tl, bottom_right=br)).run()
    ref = O.Pyramid(F.BT_KIND_PLANAR, F.BT_FORMAT_RGBA8, T, B, L)
    ref.add_datasets([O.make_dataset(crop, top_left=tl, bottom_right=br, lod_range=(4, 6))])
    ref.run()
    got = {tuple(int(v) for v in t) for t in pre2.tiles()}
    exp = {tuple(int(v) for v in t) for t in ref.tiles()}
    assert got == exp and len(exp) == 5
    for tc in sorted(exp):
        assert np.array_equal(pre2.get_tile(0, tc), ref.tile(*tc)), tc
    pre2.close()


def test_config2_frame_vertices(ctx, terrain):
    from bevy_terrain_b200 import api
    cfg, pre = terrain
    vcfg = api.TerrainViewConfig(grid_size=64)
    atlas = api.TileAtlas(ctx, cfg)
    tree = api.TileTree(atlas, vcfg)
    try:
        tree.add_existing_tiles(pre.tiles(0))
        pos = (-410.0, -100.0 + 250.0 + 5.0, 377.0)      # 5 above the surface: the 12-LOD tree refines to its deepest levels
        assert 0 < api.load_resident_tiles(tree, pre, pos, attachments=(0, 1)) <= cfg.atlas_size
        tree.prepare()
        tree.select_tiles()
        tiles, ind = tree.readback()
        assert len(tiles) > 2000 and int(tiles[:, 1].max()) >= 9      # deeper than the data's last LOD (5): best-available lookups
        flags = F.BT_FLAG_NORMALS | F.BT_FLAG_MORPH | F.BT_FLAG_BLEND
        tree.gen_vertices(flags)
        rng = np.random.default_rng(2)
        for t in rng.choice(len(tiles), 16, replace=False):
            v = check_strip_structure(tree, int(t))
            y = v["position"][..., 1]
            assert y.min() >= -100.0 - 1e-3 and y.max() <= 400.0 + 1e-3
        vi = oracle_inputs(ctx, cfg, atlas, tree, tiles, flags, pos, False, T, B, 0.0, 500.0)
        deep = np.flatnonzero(tiles[:, 1] >= 9)
        pick = np.concatenate([rng.choice(len(tiles), 24, replace=False), rng.choice(deep, min(8, len(deep)), replace=False)])
        assert_tiles_bit_exact(tree, vi, pick)
        # ---- the configs[2] sampler ("trilinear + aniso") and the per-vertex albedo at full size: CUDA == oracle on whole tiles
        fflags = flags | F.BT_FLAG_FILTER_TRILINEAR_ANISO | F.BT_FLAG_ALBEDO
        tree.gen_vertices(fflags)
        mips = cfg.attachments[0].mip_level_count
        chain = O.mip_chain_texels(T, mips)
        hm = np.empty((cfg.atlas_size, chain), dtype=np.uint16)
        ctx.d2h(hm, atlas.mips_device_ptr(0)[0])
        aa = np.empty((cfg.atlas_size, T, T, 4), dtype=np.uint8)
        ctx.d2h(aa, atlas.device_ptr(1))
        am = np.empty((cfg.atlas_size, chain, 4), dtype=np.uint8)
        ctx.d2h(am, atlas.mips_device_ptr(1)[0])
        for slot in rng.choice(cfg.atlas_size, 6, replace=False):       # the GPU mip chains are the oracle's
            assert np.array_equal(hm[slot], O.generate_mipmaps(vi._keep[0][slot], mips))
            assert np.array_equal(am[slot], O.generate_mipmaps(aa[slot], mips))
        vi.flags = fflags
        vi.attachments[1] = vi.attachments[0]
        vpt = 8448
        for t in pick[:12]:
            exp, exp_col = O.gen_vertices_filtered(vi, hm, aa, am, mips, int(t) * vpt, vpt)
            assert tree.download_vertices(int(t) * vpt, vpt).tobytes() == exp.tobytes(), f"filtered tile {t} differs"
            assert np.array_equal(tree.download_albedo(int(t) * vpt, vpt), exp_col), f"albedo of tile {t} differs"
    finally:
        tree.close(); atlas.close()

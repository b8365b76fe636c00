"""GPU parity: bt_preprocessor (split / downsample / stitch kernels) vs the oracle's restatement of
src/shaders/preprocess/*.wgsl and the task graph of preprocessor.rs.  Byte-exact."""
import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O

pytestmark = pytest.mark.gpu


def _config(api, kind, lod_count, T, B, fmt):
    model = api.TerrainModel.sphere((0, 0, 0), 1.0, 0.0, 1.0) if kind == F.BT_KIND_SPHERICAL else api.TerrainModel.planar((0, 0, 0), 1.0, 0.0, 1.0)
    cfg = api.TerrainConfig(lod_count=lod_count, model=model)
    cfg.add_attachment(api.AttachmentConfig(name="a", texture_size=T, border_size=B,
                                            format=api.AttachmentFormat.R16 if fmt == F.BT_FORMAT_R16 else api.AttachmentFormat.Rgba8))
    return cfg


def _compare(pre, pyr, attachment=0):
    got_tiles = {tuple(int(v) for v in t) for t in pre.tiles(attachment)}
    exp_tiles = {tuple(int(v) for v in t) for t in pyr.tiles()}
    assert got_tiles == exp_tiles
    for tc in sorted(exp_tiles):
        g, e = pre.get_tile(attachment, tc), pyr.tile(*tc)
        if not np.array_equal(g, e):
            bad = np.argwhere(g != e)
            raise AssertionError(f"tile {tc}: {len(bad)} texels differ, first at {bad[0]}: got {g[tuple(bad[0])]} expected {e[tuple(bad[0])]}")
    return len(exp_tiles)


@pytest.mark.parametrize("fmt", [F.BT_FORMAT_R16, F.BT_FORMAT_RGBA8])
@pytest.mark.parametrize("holes", [False, True])
def test_preprocess_planar_full_pyramid(ctx, fmt, holes):
    from bevy_terrain_b200 import api
    T, B, L = 64, 2, 4
    src = O.synth_source(fmt, 300, 280, seed=0x5EED0001 if fmt == F.BT_FORMAT_R16 else 0x5EED0002, holes=holes)
    pyr = O.Pyramid(F.BT_KIND_SPHERICAL, fmt, T, B, L)   # Q7: preprocess_planar.rs runs with the default sphere model
    pyr.add_datasets([O.make_dataset(src, lod_range=(0, L))])
    pyr.run()
    pre = api.Preprocessor(ctx, _config(api, F.BT_KIND_SPHERICAL, L, T, B, fmt))
    try:
        pre.preprocess_tile(api.PreprocessDataset(pixels=src, lod_range=range(0, L))).run()
        assert _compare(pre, pyr) == 85
    finally:
        pre.close()


def test_preprocess_planar_model_equals_sphere_model_for_one_face(ctx):
    """Q7: for side-0-only data the planar and the spherical neighbour rules give the same pyramid"""
    from bevy_terrain_b200 import api
    T, B, L = 32, 2, 3
    src = O.synth_source(F.BT_FORMAT_R16, 128, 128)
    a = api.Preprocessor(ctx, _config(api, F.BT_KIND_SPHERICAL, L, T, B, F.BT_FORMAT_R16))
    b = api.Preprocessor(ctx, _config(api, F.BT_KIND_PLANAR, L, T, B, F.BT_FORMAT_R16))
    try:
        a.preprocess_tile(api.PreprocessDataset(pixels=src, lod_range=range(0, L))).run()
        b.preprocess_tile(api.PreprocessDataset(pixels=src, lod_range=range(0, L))).run()
        for tc in a.tiles():
            assert np.array_equal(a.get_tile(0, tc), b.get_tile(0, tc))
    finally:
        a.close(); b.close()


def test_preprocess_partial_dataset_and_overlay(ctx):
    """a dataset covering part of the face (ragged tile ranges, missing children) overlaid by a second one"""
    from bevy_terrain_b200 import api
    T, B, L = 32, 2, 4
    src1 = O.synth_source(F.BT_FORMAT_R16, 200, 200, seed=11, holes=True)
    src2 = O.synth_source(F.BT_FORMAT_R16, 90, 70, seed=12)
    kw1 = dict(top_left=(0.1, 0.2), bottom_right=(0.8, 0.9), lod_range=(1, L))
    kw2 = dict(top_left=(0.3, 0.3), bottom_right=(0.55, 0.6), lod_range=(0, L))
    pyr = O.Pyramid(F.BT_KIND_PLANAR, F.BT_FORMAT_R16, T, B, L)
    pyr.add_datasets([O.make_dataset(src1, **kw1)])
    pyr.add_datasets([O.make_dataset(src2, **kw2)])
    pyr.run()
    pre = api.Preprocessor(ctx, _config(api, F.BT_KIND_PLANAR, L, T, B, F.BT_FORMAT_R16))
    try:
        pre.preprocess_tile(api.PreprocessDataset(pixels=src1, top_left=kw1["top_left"], bottom_right=kw1["bottom_right"], lod_range=range(*kw1["lod_range"])))
        pre.preprocess_tile(api.PreprocessDataset(pixels=src2, top_left=kw2["top_left"], bottom_right=kw2["bottom_right"], lod_range=range(*kw2["lod_range"])))
        pre.run()
        _compare(pre, pyr)
    finally:
        pre.close()


def test_preprocess_spherical_cross_face_stitch(ctx):
    from bevy_terrain_b200 import api
    T, B, L = 32, 2, 3
    faces = [O.synth_source(F.BT_FORMAT_R16, 96, 96, seed=100 + s, holes=(s == 2)) for s in range(6)]
    pyr = O.Pyramid(F.BT_KIND_SPHERICAL, F.BT_FORMAT_R16, T, B, L)
    pyr.add_datasets([O.make_dataset(faces[s], side=s, lod_range=(0, L)) for s in range(6)])
    pyr.run()
    pre = api.Preprocessor(ctx, _config(api, F.BT_KIND_SPHERICAL, L, T, B, F.BT_FORMAT_R16))
    try:
        pre.preprocess_spherical(api.SphericalDataset(faces=faces, lod_range=range(0, L))).run()
        assert _compare(pre, pyr) == 6 * 21
    finally:
        pre.close()


def test_preprocess_odd_sizes(ctx):
    """texture sizes that break the 16-byte row alignment take the scalar store path"""
    from bevy_terrain_b200 import api
    T, B, L = 36, 1, 3
    src = O.synth_source(F.BT_FORMAT_R16, 77, 131, seed=5)
    pyr = O.Pyramid(F.BT_KIND_PLANAR, F.BT_FORMAT_R16, T, B, L)
    pyr.add_datasets([O.make_dataset(src, lod_range=(0, L))])
    pyr.run()
    pre = api.Preprocessor(ctx, _config(api, F.BT_KIND_PLANAR, L, T, B, F.BT_FORMAT_R16))
    try:
        pre.preprocess_tile(api.PreprocessDataset(pixels=src, lod_range=range(0, L))).run()
        _compare(pre, pyr)
    finally:
        pre.close()


def test_synth_source_matches_oracle(ctx):
    for fmt, dt, shape in ((F.BT_FORMAT_R16, np.uint16, (97, 211)), (F.BT_FORMAT_RGBA8, np.uint8, (97, 211, 4))):
        for holes in (False, True):
            ptr = ctx.synth_source(fmt, 211, 97, x0=1000, y0=77, seed=42, holes=holes)
            got = np.empty(shape, dtype=dt)
            ctx.d2h(got, ptr)
            ctx.free(ptr)
            assert np.array_equal(got, O.synth_source(fmt, 211, 97, x0=1000, y0=77, seed=42, holes=holes))


def test_mipmaps_match_oracle(ctx):
    import ctypes as C
    for fmt, dt, px in ((F.BT_FORMAT_R16, np.uint16, 1), (F.BT_FORMAT_RGBA8, np.uint8, 4)):
        T, levels = 64, 4
        src = O.synth_source(fmt, T, T, seed=9, holes=True)
        n_out = sum((T >> l) ** 2 for l in range(1, levels)) * px
        exp = np.zeros(n_out, dtype=dt)
        O.lib().bto_generate_mipmaps(fmt, T, levels, src.ctypes.data, exp.ctypes.data)
        d_in, d_out = ctx.alloc(src.nbytes), ctx.alloc(exp.nbytes)
        ctx.h2d(d_in, src)
        ctx.check(ctx.lib.bt_generate_mipmaps(ctx._h, fmt, T, levels, C.c_void_p(d_in), C.c_void_p(d_out), None))
        got = np.empty_like(exp)
        ctx.d2h(got, d_out)
        ctx.free(d_in); ctx.free(d_out)
        assert np.array_equal(got, exp)


def _read_varint(buf, pos):
    b = buf[pos]
    if b < 251:
        return b, pos + 1
    n = {251: 2, 252: 4, 253: 8}[b]
    return int.from_bytes(buf[pos + 1:pos + 1 + n], "little"), pos + 1 + n


def test_save_writes_reference_tile_files_and_config_tc(ctx, tmp_path):
    """Save tasks + save_tile_config (tile_atlas.rs:77-116,605-613): raw mip-0 texels in <side>_<lod>_<x>_<y>.bin and a
    bincode-standard (varint) Vec<TileCoordinate> in config.tc (formats/mod.rs:14-24)"""
    from bevy_terrain_b200 import api
    T, B, L = 32, 2, 3
    src = O.synth_source(F.BT_FORMAT_R16, 600, 300, seed=21)   # > 250 forces multi-byte varints? (coordinates stay small; the count does not matter)
    pre = api.Preprocessor(ctx, _config(api, F.BT_KIND_PLANAR, L, T, B, F.BT_FORMAT_R16))
    try:
        pre.preprocess_tile(api.PreprocessDataset(pixels=src, lod_range=range(0, L))).run()
        pre.save(0, str(tmp_path))
        tiles = {tuple(int(v) for v in t) for t in pre.tiles()}
        for tc in tiles:
            raw = np.fromfile(tmp_path / ("%d_%d_%d_%d.bin" % tc), dtype=np.uint16).reshape(T, T)
            assert np.array_equal(raw, pre.get_tile(0, tc))
        buf = (tmp_path / "config.tc").read_bytes()
        n, pos = _read_varint(buf, 0)
        listed = set()
        for _ in range(n):
            vals = []
            for _ in range(4):
                v, pos = _read_varint(buf, pos)
                vals.append(v)
            listed.add(tuple(vals))
        assert pos == len(buf) and listed == tiles and n == 21
        # ... and back: config.tc through the C decoder, tiles streamed from the files into an atlas
        cfg_tiles = api.load_tile_config(str(tmp_path / "config.tc"))
        assert {tuple(int(v) for v in t) for t in cfg_tiles} == tiles
        atlas = api.TileAtlas(ctx, _config(api, F.BT_KIND_PLANAR, L, T, B, F.BT_FORMAT_R16))
        try:
            for i, tc in enumerate(sorted(tiles)):
                atlas.load_tile_file(0, i, str(tmp_path), tc)
            for i, tc in enumerate(sorted(tiles)):
                assert np.array_equal(atlas.download_tile(0, i), pre.get_tile(0, tc))
            with pytest.raises(F.BtError) as e:
                atlas.load_tile_file(0, 0, str(tmp_path), (0, 9, 9, 9))
            assert e.value.status == F.BT_ERR_IO
        finally:
            atlas.close()
    finally:
        pre.close()


def test_preprocess_errors_are_reported_not_fatal(ctx):
    """reference behaviour: panics / silently nothing (Rgb8, Rg16 have no shader path); here: status codes"""
    from bevy_terrain_b200 import api
    import ctypes as C
    att = (F.AttachmentConfig * 1)(F.AttachmentConfig(32, 2, 1, 5))       # Rgb8 id = 5 (terrain_data/mod.rs:52)
    h = C.c_void_p()
    assert ctx.lib.bt_preprocessor_create(ctx._h, F.BT_KIND_PLANAR, att, 1, 3, F.Shard(0, 1), C.byref(h)) == F.BT_ERR_INVALID_ARG
    pre = api.Preprocessor(ctx, _config(api, F.BT_KIND_PLANAR, 3, 32, 2, F.BT_FORMAT_R16))
    try:
        with pytest.raises(F.BtError):   # lod_range beyond lod_count
            pre.preprocess_tile(api.PreprocessDataset(pixels=np.ones((8, 8), np.uint16), lod_range=range(0, 9)))
        with pytest.raises(F.BtError):   # empty extent
            pre.preprocess_tile(api.PreprocessDataset(pixels=np.ones((8, 8), np.uint16), top_left=(0.5, 0.5), bottom_right=(0.5, 0.9), lod_range=range(0, 2)))
        assert ctx.lib.bt_preprocessor_get_tile(pre._h, 0, C.byref(F.Tile(0, 0, 0, 0)), C.c_void_p(1), 32 * 32 * 2) == F.BT_ERR_NOT_FOUND
    finally:
        pre.close()


@pytest.mark.parametrize("fmt,src_w,src_h,T,B,L", [
    (F.BT_FORMAT_R16, 256, 256, 64, 2, 3),      # ratio 1.07, 512-byte rows: the TMA-staged form of the split
    (F.BT_FORMAT_R16, 128, 136, 64, 2, 3),      # source coarser than the tile grid (ratio 0.53 / 0.57): windows of ~10 rows
    (F.BT_FORMAT_R16, 928, 904, 64, 2, 3),      # ratio 3.87 / 3.77: the largest windows the stage geometry accepts (limit 4)
    (F.BT_FORMAT_R16, 1008, 1000, 64, 2, 3),    # ratio 4.2: no staging, every CTA gathers from global memory
    (F.BT_FORMAT_RGBA8, 252, 260, 64, 2, 3),    # Rgba8, 1008-byte rows (16-byte aligned): staged, one stage, shared-texel decode reuse
    (F.BT_FORMAT_RGBA8, 250, 250, 64, 2, 3),    # Rgba8, 1000-byte rows: not 16-byte aligned -> global form
    (F.BT_FORMAT_R16, 1400, 1320, 640, 2, 2),   # tiles wider than 512 pixels: two column blocks per band group
    (F.BT_FORMAT_RGBA8, 1024, 1024, 520, 4, 2), # ratio exactly 1 (C = 512), two column blocks, Rgba8
])
def test_split_forms_and_stage_geometries(ctx, fmt, src_w, src_h, T, B, L):
    """Both instantiations of split_staged_kernel (TMA-staged windows / global gather), windows from half a source row per output
    row to four, one and two column blocks: every tile byte-exact against the oracle."""
    from bevy_terrain_b200 import api
    src = O.synth_source(fmt, src_w, src_h, seed=0x5EED0000 + src_w, holes=(src_w % 16 == 0))
    pyr = O.Pyramid(F.BT_KIND_PLANAR, fmt, T, B, L)
    pyr.add_datasets([O.make_dataset(src, lod_range=(0, L))])
    pyr.run()
    pre = api.Preprocessor(ctx, _config(api, F.BT_KIND_PLANAR, L, T, B, fmt))
    try:
        pre.preprocess_tile(api.PreprocessDataset(pixels=src, lod_range=range(0, L))).run()
        assert _compare(pre, pyr) == (4 ** L - 1) // 3
    finally:
        pre.close()

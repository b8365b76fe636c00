"""BASELINE configs[2] variant: BT_FLAG_FILTER_TRILINEAR_ANISO (the reference's linear / linear / linear + anisotropy-16 sampler
evaluated per vertex over the generate_mipmaps chain) and BT_FLAG_ALBEDO (per-vertex sRGB-decoded colour stream), CUDA through
the C ABI against the oracle, bit for bit — vertices AND colours — on planar and spherical scenes, plus the GPU mip chains."""
import ctypes as C

import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O
from tests.scene import Scene, planar_view_positions, spherical_view_positions

pytestmark = pytest.mark.gpu
MIPS = 3
ALL = F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS


def _albedo_pyramid(scene):
    srcs = [O.synth_source(F.BT_FORMAT_RGBA8, scene.src_size, scene.src_size, seed=0xA1BED0 + 7 * s) for s in range(scene.sides)]
    pyr = O.Pyramid(scene.kind, F.BT_FORMAT_RGBA8, scene.T, scene.B, scene.lod_count)
    pyr.add_datasets([O.make_dataset(srcs[s], side=s, lod_range=(0, scene.lod_count)) for s in range(scene.sides)])
    pyr.run()
    return pyr


def _setup(ctx, kind):
    from bevy_terrain_b200 import api
    # blend_distance 64 instead of 2: with the default view config a grid step is always SMALLER than a texel of the sampled LOD
    # (tile lod ~ blend lod + 3), so the sampler would never leave mip 0; a larger blend distance makes geometry tiles sample
    # data several LODs finer than themselves and exercises the coarse mips, the mip interpolation and the "up" lookups
    # (load_distance 64 keeps every data LOD resident around the view; the spherical scene uses a 4 x 4 grid so that a grid step spans
    #  several texels of its 32-texel tiles)
    scene = Scene(kind, lod_count=3 if kind == "spherical" else 4, T=32 if kind == "spherical" else 64, B=2, src=64 if kind == "spherical" else 256,
                  grid_size=4 if kind == "spherical" else 16, blend_distance=64.0, load_distance=64.0)
    apyr = _albedo_pyramid(scene)
    cfg = scene.terrain_config()
    cfg.attachments[0].mip_level_count = MIPS
    cfg.add_attachment(api.AttachmentConfig(name="albedo", texture_size=scene.T, border_size=scene.B, mip_level_count=MIPS, format=api.AttachmentFormat.Rgba8))
    atlas = api.TileAtlas(ctx, cfg)
    tree = api.TileTree(atlas, scene.cfg)
    albedo_host = np.zeros((scene.atlas_size, scene.T, scene.T, 4), dtype=np.uint8)
    return scene, apyr, cfg, atlas, tree, albedo_host


@pytest.mark.parametrize("kind", ["planar", "spherical"])
def test_filtered_sampler_and_albedo_bit_exact(ctx, kind):
    scene, apyr, cfg, atlas, tree, albedo_host = _setup(ctx, kind)
    positions = spherical_view_positions()[:3] if kind == "spherical" else [planar_view_positions(1)[0], (0.0, 900.0, 0.0), (350.0, 400.0, -120.0)]
    T = scene.T
    chain = O.mip_chain_texels(T, MIPS)
    try:
        for pos in positions:
            for tc, idx in scene.move_view(pos):
                albedo_host[idx] = apyr.tile(*tc)
                atlas.upload_tile(0, idx, scene.atlas[idx])
                atlas.upload_tile(1, idx, albedo_host[idx])
                atlas.generate_mips(0, idx)
                atlas.generate_mips(1, idx)
            # host mip chains of every slot (the oracle's generate_mipmaps) and the device's must agree byte for byte
            hm = np.stack([O.generate_mipmaps(scene.atlas[i], MIPS) for i in range(scene.atlas_size)])
            am = np.stack([O.generate_mipmaps(albedo_host[i], MIPS) for i in range(scene.atlas_size)])
            ptr, nbytes = atlas.mips_device_ptr(0)
            assert nbytes == chain * 2
            dev = np.empty_like(hm)
            ctx.d2h(dev, ptr)
            assert np.array_equal(dev, hm)
            ptr, nbytes = atlas.mips_device_ptr(1)
            dev = np.empty_like(am)
            ctx.d2h(dev, ptr)
            assert np.array_equal(dev, am)
            tree.prepare(frame=scene.view_frame())
            tree.select_tiles()
            tiles, ind = tree.readback()
            n = ind.vertex_count
            for flags in (ALL | F.BT_FLAG_FILTER_TRILINEAR_ANISO | F.BT_FLAG_ALBEDO, F.BT_FLAG_NORMALS | F.BT_FLAG_ALBEDO,
                          F.BT_FLAG_FILTER_TRILINEAR_ANISO | F.BT_FLAG_BLEND):
                tree.gen_vertices(flags)
                got = tree.download_vertices(0, n)
                vi = scene.oracle_vertex_inputs(tiles, flags)
                vi.attachments[1] = vi.attachments[0]
                exp, exp_albedo = O.gen_vertices_filtered(vi, hm, albedo_host, am, MIPS, 0, n)
                for f in ("position", "normal", "uv"):
                    g, e = got[f].view(np.uint32), exp[f].view(np.uint32)
                    assert np.array_equal(g, e), f"{kind} flags {flags:#x} {f}: {int((g != e).sum())} components differ"
                assert np.array_equal(got["tile_index"], exp["tile_index"])
                if flags & F.BT_FLAG_ALBEDO:
                    colours = tree.download_albedo(0, n)
                    assert np.array_equal(colours, exp_albedo), f"{kind} flags {flags:#x}: {int((colours != exp_albedo).sum())} colours differ"
                    assert (colours >> 24 == 255).all() and len(np.unique(colours)) > 100          # opaque, and a real image
                if flags == (F.BT_FLAG_NORMALS | F.BT_FLAG_ALBEDO):    # without the filter flag the vertices are exactly the plain path's
                    plain = scene.oracle_vertices(tiles, F.BT_FLAG_NORMALS)
                    assert got.tobytes() == plain.tobytes()
            # the filter really selects coarser mips: with it the heights differ from the mip-0 path somewhere
            tree.gen_vertices(F.BT_FLAG_FILTER_TRILINEAR_ANISO)
            a = tree.download_vertices(0, n)["position"]
            tree.gen_vertices(0)
            b = tree.download_vertices(0, n)["position"]
            assert not np.array_equal(a, b)
    finally:
        tree.close(); atlas.close(); apyr.close()


def test_albedo_needs_an_rgba8_attachment(ctx):
    from bevy_terrain_b200 import api
    scene = Scene("planar", lod_count=3, T=32, B=2, src=64, grid_size=16)
    atlas = api.TileAtlas(ctx, scene.terrain_config())
    tree = api.TileTree(atlas, scene.cfg)
    try:
        scene.move_view(planar_view_positions(1)[0])
        tree.prepare(frame=scene.view_frame())
        tree.select_tiles()
        with pytest.raises(F.BtError) as e:
            tree.gen_vertices(F.BT_FLAG_ALBEDO)
        assert e.value.status == F.BT_ERR_INVALID_ARG
    finally:
        tree.close(); atlas.close()

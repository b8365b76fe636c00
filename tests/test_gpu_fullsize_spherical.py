"""BASELINE.json configs[3] at full size: spherical terrain, six 8192^2 cube faces (GEBCO-shaped synthetic R16), terrain with
10 LODs of which the dataset resolves LODs 0..4 (8192 / 508 = 16 tiles per face side), texture 512, grid 64.
The scalar oracle cannot recompute all of it, so: size-independent properties of the pyramid and of the vertex stream, and a
bit-exact comparison with the oracle on whole sampled tiles using the product's own inputs (atlas + table downloaded)."""
import ctypes as C

import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O

pytestmark = pytest.mark.gpu
T, B, L, S, DATA_LODS = 512, 2, 10, 8192, 5
Cn = T - 2 * B
R, MIN_H, MAX_H = 6371000.0, -12000.0, 9000.0     # examples/spherical.rs:5-9


@pytest.fixture(scope="module")
def globe(ctx):
    from bevy_terrain_b200 import api
    model = api.TerrainModel.sphere((0.0, 0.0, 0.0), R, MIN_H, MAX_H)
    cfg = api.TerrainConfig(lod_count=L, model=model, atlas_size=6 * DATA_LODS * 64)
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=T, border_size=B, mip_level_count=4, format=api.AttachmentFormat.R16))
    faces = [ctx.synth_source(F.BT_FORMAT_R16, S, S, seed=0x5EED0001 + 17 * s) for s in range(6)]
    pre = api.Preprocessor(ctx, cfg)
    pre.preprocess_spherical(api.SphericalDataset(device_faces=faces, width=S, height=S, lod_range=range(0, DATA_LODS))).run()
    yield cfg, pre
    pre.close()
    for p in faces:
        ctx.free(p)


def test_globe_pyramid_tiles_and_same_face_borders(globe):
    cfg, pre = globe
    tiles = pre.tiles()
    assert len(tiles) == 6 * (4 ** DATA_LODS - 1) // 3 == 2046
    rng = np.random.default_rng(3)
    for side in (0, 3, 5):
        for lod in (2, 4):
            n = 1 << lod
            x, y = int(rng.integers(0, n - 1)), int(rng.integers(0, n - 1))
            a = pre.get_tile(0, (side, lod, x, y))
            right, below = pre.get_tile(0, (side, lod, x + 1, y)), pre.get_tile(0, (side, lod, x, y + 1))
            assert np.array_equal(a[B:B + Cn, B + Cn:], right[B:B + Cn, B:2 * B])     # stitched border == neighbour centre
            assert np.array_equal(a[B + Cn:, B:B + Cn], below[B:2 * B, B:B + Cn])
            assert (a != 0).all()
    # a face-edge tile: its outer border comes from ANOTHER face (cross-face stitch), so it is filled, not clamp-repeated
    edge = pre.get_tile(0, (1, 4, 0, 7))
    assert (edge[:, :B] != 0).all()


def test_globe_frame_vertices(ctx, globe):
    from bevy_terrain_b200 import api
    cfg, pre = globe
    vcfg = api.TerrainViewConfig(grid_size=64)
    atlas = api.TileAtlas(ctx, cfg)
    tree = api.TileTree(atlas, vcfg)
    try:
        tree.add_existing_tiles(pre.tiles())
        # 30 km above the surface near a cube-face corner region (three faces in view)
        d = np.array([0.62, 0.55, 0.56]); d /= np.linalg.norm(d)
        pos = tuple(float(v) for v in d * (R + 30000.0))
        n_loads = api.load_resident_tiles(tree, pre, pos)
        assert 0 < n_loads <= cfg.atlas_size
        tree.prepare()
        tree.select_tiles()
        tiles, ind = tree.readback()
        assert 500 < len(tiles) < 60000 and ind.vertex_count == len(tiles) * 8448
        assert len({int(t[0]) for t in tiles}) >= 3                     # several cube faces contribute tiles
        flags = F.BT_FLAG_NORMALS | F.BT_FLAG_MORPH | F.BT_FLAG_BLEND
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tree.gen_vertices(flags)
        e0.record(); tree.gen_vertices(flags); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"\nconfigs[3] spherical frame: {len(tiles)} tiles, {ind.vertex_count / 1e6:.1f} M vertices, gen_vertices {ms:.3f} ms "
              f"-> {ind.vertex_count / ms / 1e3:.0f} Mverts/s")
        rng = np.random.default_rng(1)
        for t in rng.choice(len(tiles), 16, replace=False):
            v = tree.download_vertices(int(t) * 8448, 8448).reshape(64, 132)
            raw = np.ascontiguousarray(v).view(np.uint8).reshape(64, 132, 40)
            assert np.array_equal(raw[:, 0], raw[:, 1]) and np.array_equal(raw[:, -1], raw[:, -2])
            assert np.array_equal(v[:-1, 2::2]["position"][:, :65], v[1:, 1::2]["position"][:, :65])
            assert (v["tile_index"] == t).all()
            radius = np.linalg.norm(v["position"][..., :3].astype(np.float64), axis=-1)
            assert radius.min() >= R + MIN_H - 2.0 and radius.max() <= R + MAX_H + 2.0
            nrm = np.linalg.norm(v["normal"].astype(np.float64), axis=-1)
            assert nrm.max() <= 1.0 + 1e-5 and nrm.min() > 0.5
        # bit-exact against the oracle on whole tiles, with the product's own inputs
        host_atlas = np.empty((cfg.atlas_size, T, T), dtype=np.uint16)
        ctx.d2h(host_atlas, atlas.device_ptr(0))
        vi = O.VertexInputs()
        vi.view_config, vi.approximation = tree.frame.view_config, tree.frame.approximation
        vi.terrain_config = F.TerrainConfigUniform(L, MIN_H, MAX_H, float(np.float32(cfg.model.scale())))
        vi.attachments[0].size, vi.attachments[0].scale, vi.attachments[0].offset = float(Cn), float(np.float32(Cn) / np.float32(T)), float(np.float32(B) / np.float32(T))
        vi.mesh = cfg.model.mesh_transform()
        vi.view_world_position[:] = [float(np.float32(p)) for p in pos]
        vi.spherical, vi.texture_size = 1, T
        entries = np.frombuffer(tree.entries, dtype=np.uint32).reshape(-1, 2).copy()
        vi.tile_tree = entries.ctypes.data_as(C.POINTER(F.TileTreeEntry))
        vi.height_atlas = host_atlas.ctypes.data
        vi.tiles = tiles.ctypes.data_as(C.POINTER(F.Tile))
        vi.tile_count, vi.flags = len(tiles), flags
        for t in rng.choice(len(tiles), 32, replace=False):
            exp = O.gen_vertices(vi, int(t) * 8448, 8448)
            got = tree.download_vertices(int(t) * 8448, 8448)
            assert got.tobytes() == exp.tobytes(), f"tile {t} differs"
    finally:
        tree.close(); atlas.close()

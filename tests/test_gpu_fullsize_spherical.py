"""BASELINE.json configs[3] at full size: spherical terrain, six 8192^2 cube faces (GEBCO-shaped synthetic R16), terrain with
10 LODs of which the dataset resolves LODs 0..4 (8192 / 508 = 16 tiles per face side), texture 512, grid 64.
The scalar oracle cannot recompute all of it, so: size-independent properties of the pyramid and of the vertex stream, and a
bit-exact comparison with the oracle on whole sampled tiles using the product's own inputs (atlas + table downloaded)."""

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


def _direction(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


ALL = F.BT_FLAG_NORMALS | F.BT_FLAG_MORPH | F.BT_FLAG_BLEND
VIEWS = [   # the example's camera altitudes (examples/spherical.rs:96-100) + corner / edge views
    ("corner_30km", _direction([0.62, 0.55, 0.56]) * (R + 30000.0), ALL),
    ("minus_x_3km_high_precision", _direction([-1.0, 0.003, 0.002]) * (R + 3000.0), ALL | F.BT_FLAG_HIGH_PRECISION),   # < 0.001 R from the surface
    ("face_edge_100km", _direction([1.0, 0.02, 1.0]) * (R + 100000.0), ALL),
    ("far_3R", _direction([-1.0, 0.2, 0.1]) * (3.0 * R), ALL),
]


@pytest.mark.parametrize("name,pos,flags", VIEWS, ids=[v[0] for v in VIEWS])
def test_globe_frame_vertices(ctx, globe, name, pos, flags):
    from bevy_terrain_b200 import api
    from tests.fullsize_util import assert_tiles_bit_exact, check_strip_structure, oracle_inputs
    cfg, pre = globe
    pos = tuple(float(v) for v in pos)
    vcfg = api.TerrainViewConfig(grid_size=64)
    atlas = api.TileAtlas(ctx, cfg)
    tree = api.TileTree(atlas, vcfg)
    try:
        tree.add_existing_tiles(pre.tiles())
        n_loads = api.load_resident_tiles(tree, pre, pos)
        assert 0 < n_loads <= cfg.atlas_size
        tree.prepare()
        tree.select_tiles()
        tiles, ind = tree.readback()
        assert 6 <= len(tiles) < 120000 and ind.vertex_count == len(tiles) * 8448
        if name != "minus_x_3km_high_precision":
            assert len({int(t[0]) for t in tiles}) >= 2                  # several cube faces contribute tiles
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tree.gen_vertices(flags)
        e0.record(); tree.gen_vertices(flags); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"\nconfigs[3] {name}: {len(tiles)} tiles, {ind.vertex_count / 1e6:.1f} M vertices, gen_vertices {ms:.3f} ms "
              f"-> {ind.vertex_count / ms / 1e3:.0f} Mverts/s")
        rng = np.random.default_rng(1)
        for t in rng.choice(len(tiles), min(16, len(tiles)), replace=False):
            v = check_strip_structure(tree, int(t))
            radius = np.linalg.norm(v["position"][..., :3].astype(np.float64), axis=-1)
            assert radius.min() >= R + MIN_H - 2.0 and radius.max() <= R + MAX_H + 2.0
            nrm = np.linalg.norm(v["normal"].astype(np.float64), axis=-1)
            assert nrm.max() <= 1.0 + 1e-5 and nrm.min() > 0.5
        # bit-exact against the oracle on whole tiles, with the product's own inputs
        vi = oracle_inputs(ctx, cfg, atlas, tree, tiles, flags, pos, True, T, B, MIN_H, MAX_H)
        assert_tiles_bit_exact(tree, vi, rng.choice(len(tiles), min(32, len(tiles)), replace=False))
        if flags & F.BT_FLAG_HIGH_PRECISION:
            # the deepest tiles lie within precision_threshold_distance of the camera: the Taylor path must be taken there
            deepest = np.flatnonzero(tiles[:, 1] >= tiles[:, 1].max() - 1)
            vi.flags = flags & ~F.BT_FLAG_HIGH_PRECISION
            taken = [int(t) for t in deepest if O.gen_vertices(vi, int(t) * 8448, 8448).tobytes() != tree.download_vertices(int(t) * 8448, 8448).tobytes()]
            assert taken, "the view does not exercise HIGH_PRECISION"
            vi.flags = flags
            assert_tiles_bit_exact(tree, vi, taken[:16])      # ... and where it is taken, the result is still the oracle's
            print(f"HIGH_PRECISION path taken in {len(taken)} of the {len(deepest)} deepest tiles")
    finally:
        tree.close(); atlas.close()

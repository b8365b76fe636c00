"""Self-consistency of the oracle's preprocess restatement."""
import numpy as np

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O


def _run(kind, fmt, T, B, L, datasets):
    p = O.Pyramid(kind, fmt, T, B, L)
    p.add_datasets(datasets)
    p.run()
    return p


def test_constant_field_stays_constant():
    T, B, L = 32, 2, 4
    src = np.full((100, 100), 12345, dtype=np.uint16)
    p = _run(F.BT_KIND_PLANAR, F.BT_FORMAT_R16, T, B, L, [O.make_dataset(src, lod_range=(0, L))])
    assert len(p.tiles()) == 85
    for tc in p.tiles():
        assert (p.tile(*tc) == 12345).all()


def test_tile_count_and_naming_order():
    """overlapping_tiles iterates x outer / y inner (iproduct, preprocessor.rs:58-66); full pyramid = (4^L-1)/3"""
    src = O.synth_source(F.BT_FORMAT_R16, 64, 64)
    p = _run(F.BT_KIND_PLANAR, F.BT_FORMAT_R16, 16, 2, 3, [O.make_dataset(src, lod_range=(0, 3))])
    tiles = p.tiles()
    assert len(tiles) == 21
    assert tiles[:4].tolist() == [[0, 0, 0, 0], [0, 1, 0, 0], [0, 1, 0, 1], [0, 1, 1, 0]]


def test_stitched_borders_equal_neighbour_centres():
    T, B, L = 32, 2, 3
    C_ = T - 2 * B
    src = O.synth_source(F.BT_FORMAT_R16, 200, 200, seed=3)
    p = _run(F.BT_KIND_PLANAR, F.BT_FORMAT_R16, T, B, L, [O.make_dataset(src, lod_range=(0, L))])
    a, right = p.tile(0, 2, 1, 1), p.tile(0, 2, 2, 1)
    below, diag = p.tile(0, 2, 1, 2), p.tile(0, 2, 2, 2)
    assert np.array_equal(a[B:B + C_, B + C_:], right[B:B + C_, B:2 * B])      # right border <- neighbour's first centre columns
    assert np.array_equal(a[B + C_:, B:B + C_], below[B:2 * B, B:B + C_])      # bottom border
    assert np.array_equal(a[B + C_:, B + C_:], diag[B:2 * B, B:2 * B])         # corner
    edge = p.tile(0, 2, 0, 0)                                                   # terrain edge: clamp-repeat
    assert np.array_equal(edge[:B, B:B + C_], np.repeat(edge[B:B + 1, B:B + C_], B, axis=0))
    assert (edge[:B, :B] == edge[B, B]).all()


def test_downsample_validity_and_q3():
    """zero = no-data: excluded from the mean; an all-zero 2x2 block stays zero (Q3, not NaN)"""
    T, B, L = 16, 2, 2
    src = np.full((48, 48), 40000, dtype=np.uint16)
    src[:24, :24] = 0
    p = _run(F.BT_KIND_PLANAR, F.BT_FORMAT_R16, T, B, L, [O.make_dataset(src, lod_range=(0, L))])
    root = p.tile(0, 0, 0, 0)
    c = root[B:-B, B:-B]
    assert set(np.unique(c)) <= {0, 40000}
    assert (c[:4, :4] == 0).all() and (c[-4:, -4:] == 40000).all()


def test_downsample_close_to_cpu_mip_twin():
    """the shader's float mean (downsample.wgsl:29-39) vs the CPU twin generate_mipmap_r16 (integer mean,
    terrain_data/mod.rs:168-198): same validity rule, values within 1 (round-half-up vs floor)"""
    T = 32
    rng = np.random.default_rng(5)
    tile = rng.integers(0, 65536, (T, T)).astype(np.uint16)
    tile[rng.random((T, T)) < 0.2] = 0
    out = np.zeros((T // 2) ** 2, dtype=np.uint16)
    O.lib().bto_generate_mipmaps(F.BT_FORMAT_R16, T, 2, tile.ctypes.data, out.ctypes.data)
    mip = out.reshape(T // 2, T // 2)
    blocks = tile.reshape(T // 2, 2, T // 2, 2).transpose(0, 2, 1, 3).reshape(T // 2, T // 2, 4).astype(np.int64)
    cnt = (blocks != 0).sum(-1)
    exp = np.where(cnt > 0, blocks.sum(-1) // np.maximum(cnt, 1), 0)
    assert np.array_equal(mip, exp)


def test_spherical_cross_face_borders_are_consistent():
    """the border of a tile on a face edge equals the rotated centre strip of the neighbouring face's tile"""
    T, B, L = 16, 2, 2
    faces = [O.synth_source(F.BT_FORMAT_R16, 48, 48, seed=50 + s) for s in range(6)]
    p = _run(F.BT_KIND_SPHERICAL, F.BT_FORMAT_R16, T, B, L, [O.make_dataset(faces[s], side=s, lod_range=(0, L)) for s in range(6)])
    assert len(p.tiles()) == 30
    C_ = T - 2 * B
    # side 0, +x edge -> side 1 (NEIGHBOURING_SIDES[0][3] = 1), index (6+1-0)%6 = 1: (PS, PT) = no rotation
    a, b = p.tile(0, 1, 1, 0), p.tile(1, 1, 0, 0)
    assert np.array_equal(a[B:B + C_, B + C_:], b[B:B + C_, B:2 * B])
    # every stitched border texel is non-zero where sources are non-zero
    for tc in p.tiles():
        assert (p.tile(*tc) != 0).all()


def test_neighbour_table_is_symmetric_on_the_cube():
    """TileCoordinate::neighbours (coordinate.rs:208-279): if B is the edge neighbour of A across a face edge then A is an edge neighbour of B"""
    L = O.lib()
    import ctypes as C
    lod = 2
    n = 1 << lod
    for side in range(6):
        for x in range(n):
            for y in range(n):
                out = (F.Tile * 8)()
                t = F.Tile(side, lod, x, y)
                L.bto_tile_neighbours(C.byref(t), 1, out)
                for k in range(4):  # edge neighbours
                    nb = out[k]
                    assert nb.side != 0xFFFFFFFF
                    back = (F.Tile * 8)()
                    L.bto_tile_neighbours(C.byref(nb), 1, back)
                    assert any((b.side, b.lod, b.x, b.y) == (side, lod, x, y) for b in back[:4]), (side, x, y, k)


def test_rgba8_alpha_and_channels():
    T, B, L = 16, 2, 2
    src = O.synth_source(F.BT_FORMAT_RGBA8, 40, 40, seed=8)
    p = _run(F.BT_KIND_PLANAR, F.BT_FORMAT_RGBA8, T, B, L, [O.make_dataset(src, lod_range=(0, L))])
    for tc in p.tiles():
        t = p.tile(*tc)
        assert (t[..., 3] == 255).all() and t.shape == (T, T, 4)

"""Frustum-cull EXTENSION (BT_FLAG_FRUSTUM_CULL, default off: the reference does not cull at this commit)."""
import ctypes as C

import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O
from tests.frustum import look_at_perspective, planes_numpy
from tests.scene import Scene


def _cull_uniform(scene, m):
    c = O.culling(scene.view_pos)
    pl = planes_numpy(m)
    for i in range(5):
        c.planes[i][:] = [float(v) for v in pl[i]]
    return c


def test_planes_helper_matches_numpy():
    lib = F.load_library()
    m = look_at_perspective((10.0, 170.0, -30.0), (300.0, 100.0, 200.0))
    colmajor = np.ascontiguousarray(m.T).reshape(-1)   # the C ABI takes column-major like glam's Mat4
    out = ((F.f32 * 4) * 5)()
    lib.bt_culling_planes(colmajor.ctypes.data_as(C.POINTER(F.f32)), out)
    assert np.array_equal(np.array([list(r) for r in out], np.float32), planes_numpy(m))


def test_oracle_cull_is_a_conservative_subset():
    scene = Scene("planar", lod_count=4, T=16, B=2, src=32, grid_size=16)
    pos = (10.0, 170.0, -30.0)
    scene.move_view(pos)
    m = look_at_perspective(pos, (300.0, 100.0, 200.0))
    cull = _cull_uniform(scene, m)
    full, _, _ = O.select_tiles(scene.vc, scene.approx, cull, scene.mesh, False)
    kept, ind, res = O.select_tiles(scene.vc, scene.approx, cull, scene.mesh, False, frustum_cull=True, min_height=0.0, max_height=500.0)
    full_set, kept_set = {tuple(t) for t in full.tolist()}, {tuple(t) for t in kept.tolist()}
    assert kept_set <= full_set and 0 < len(kept_set) < 0.6 * len(full_set)
    assert ind.vertex_count == len(kept) * scene.vc.vertices_per_tile
    # conservative: every culled tile has all of a dense sample of its surface points outside the frustum
    pl = planes_numpy(m).astype(np.float64)
    rng = np.random.default_rng(0)
    for s, l, x, y in rng.permutation(sorted(full_set - kept_set))[:300]:
        u, v = np.meshgrid(np.linspace(0, 1, 5), np.linspace(0, 1, 5))
        px = ((x + u) / (1 << l) - 0.5) * 2000.0
        pz = ((y + v) / (1 << l) - 0.5) * 2000.0
        for h in (0.0, 250.0, 500.0):
            pts = np.stack([px.ravel(), np.full(25, -100.0 + h), pz.ravel(), np.ones(25)], 1)
            inside = ((pts @ pl.T) >= 0).all(axis=1)
            assert not inside.any(), (s, l, x, y)
    # with planes that accept everything the extension is the identity
    open_cull = O.culling(pos)
    for i in range(5):
        open_cull.planes[i][:] = [0.0, 0.0, 0.0, 1.0]
    same, _, _ = O.select_tiles(scene.vc, scene.approx, open_cull, scene.mesh, False, frustum_cull=True, min_height=0.0, max_height=500.0)
    assert np.array_equal(same, full)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["planar", "spherical"])
def test_cuda_cull_matches_oracle_in_order(ctx, kind):
    from bevy_terrain_b200 import api
    scene = Scene(kind, lod_count=3, T=16, B=2, src=32, grid_size=16)
    if kind == "planar":
        views = [((10.0, 170.0, -30.0), (300.0, 100.0, 200.0)), ((-500.0, 400.0, 100.0), (0.0, 150.0, 0.0))]
        hmin, hmax = 0.0, 500.0
    else:
        R = 6371000.0
        views = [((-(R + 20000.0), 0.0, 0.0), (-R, 50000.0, 80000.0)), ((0.0, R + 8000.0, 1000.0), (90000.0, R - 2000.0, 40000.0))]
        hmin, hmax = -12000.0, 9000.0
    atlas = api.TileAtlas(ctx, scene.terrain_config())
    tree = api.TileTree(atlas, scene.cfg)
    try:
        for pos, target in views:
            scene.move_view(pos)
            m = look_at_perspective(pos, target, near=1.0, far=1e8)
            frame = scene.view_frame()
            frame.culling = _cull_uniform(scene, m)
            tree.prepare(frame=frame)
            tree.select_tiles(flags=F.BT_FLAG_FRUSTUM_CULL)
            tiles, ind = tree.readback()
            exp, exp_ind, res = O.select_tiles(scene.vc, scene.approx, frame.culling, scene.mesh, scene.spherical, frustum_cull=True,
                                               min_height=hmin, max_height=hmax)
            assert 0 < len(exp) and np.array_equal(tiles, exp)
            assert ind.vertex_count == exp_ind.vertex_count
            tree.select_tiles(flags=0)       # default = reference behaviour (no cull)
            tiles0, _ = tree.readback()
            exp0, _, _ = scene.oracle_select()
            assert np.array_equal(tiles0, exp0) and len(tiles0) > len(tiles)
    finally:
        tree.close(); atlas.close()


# ---- min / max-height culling (BT_FLAG_MINMAX_CULL) -----------------------------------------------------------------
def _minmax_table(scene):
    return np.stack([O.tile_minmax(scene.atlas[i], scene.B) for i in range(scene.atlas_size)])


def test_oracle_minmax_table_bounds_every_bilinear_sample():
    """the table really bounds the heights the vertex path can produce inside a region (conservative culling depends on it)"""
    scene = Scene("planar", lod_count=3, T=32, B=2, src=96, grid_size=16, holes=True)
    scene.move_view((50.0, 300.0, 20.0))
    tab = _minmax_table(scene)
    rng = np.random.default_rng(4)
    loaded = [i for i in range(scene.atlas_size) if scene.atlas[i].any()]
    Cn, T = scene.T - 2 * scene.B, scene.T
    for i in rng.choice(loaded, min(8, len(loaded)), replace=False):
        tile = scene.atlas[i]
        for e in [0] + [int(v) for v in rng.integers(1, 65, 6)]:
            u0, u1, v0, v1 = (0.0, 1.0, 0.0, 1.0) if e == 0 else (((e - 1) % 8) / 8, ((e - 1) % 8 + 1) / 8, ((e - 1) // 8) / 8, ((e - 1) // 8 + 1) / 8)
            us, vs = rng.uniform(u0, u1, 64), rng.uniform(v0, v1, 64)
            for u, v in zip(us, vs):      # attachment_uv + exact bilinear of non-zero texels
                s = O.lib().bto_sample_r16(tile.ctypes.data, T, np.float32(u * Cn / T + scene.B / T), np.float32(v * Cn / T + scene.B / T))
                px, py = u * Cn + scene.B - 0.5, v * Cn + scene.B - 0.5
                foot = tile[int(np.floor(py)):int(np.floor(py)) + 2, int(np.floor(px)):int(np.floor(px)) + 2]
                if (foot == 0).any():
                    continue              # a no-data texel in the footprint: the sample is not a height
                assert tab[i, e, 0] / 65535.0 - 1e-6 <= s <= tab[i, e, 1] / 65535.0 + 1e-6


def test_oracle_minmax_cull_is_tighter_and_still_conservative():
    scene = Scene("planar", lod_count=4, T=16, B=2, src=32, grid_size=16)
    pos = (10.0, 170.0, -30.0)
    scene.move_view(pos)
    m = look_at_perspective(pos, (300.0, 100.0, 200.0))
    cull = _cull_uniform(scene, m)
    tab = _minmax_table(scene)
    full, _, _ = O.select_tiles(scene.vc, scene.approx, cull, scene.mesh, False)
    wide, _, _ = O.select_tiles(scene.vc, scene.approx, cull, scene.mesh, False, frustum_cull=True, min_height=0.0, max_height=500.0)
    tight, ind, _ = O.select_tiles_minmax(scene.vc, scene.approx, cull, scene.mesh, False, 0.0, 500.0, scene.entries, tab, scene.lod_count)
    full_set, wide_set, tight_set = ({tuple(t) for t in a.tolist()} for a in (full, wide, tight))
    assert tight_set <= full_set and len(tight_set) < len(wide_set)                # tighter bounds cull strictly more here
    assert ind.vertex_count == len(tight) * scene.vc.vertices_per_tile
    # conservative: a culled tile has no surface point (height taken from the RESIDENT data through the vertex path's own sampler bounds) in the frustum
    pl = planes_numpy(m).astype(np.float64)
    src = scene.sources[0].astype(np.float64) / 65535.0 * 500.0                    # the terrain's true heights
    n = src.shape[0]
    for s, l, x, y in sorted(full_set - tight_set)[:400]:
        u, v = np.meshgrid(np.linspace(0, 1, 6), np.linspace(0, 1, 6))
        gx, gy = (x + u) / (1 << l), (y + v) / (1 << l)
        h = src[np.clip((gy * n).astype(int), 0, n - 1), np.clip((gx * n).astype(int), 0, n - 1)]
        pts = np.stack([(gx.ravel() - 0.5) * 2000.0, -100.0 + h.ravel(), (gy.ravel() - 0.5) * 2000.0, np.ones(36)], 1)
        assert not ((pts @ pl.T) >= 0).all(axis=1).any(), (s, l, x, y)
    # an empty table (nothing resident) and open planes both fall back to known behaviour
    none, _, _ = O.select_tiles_minmax(scene.vc, scene.approx, cull, scene.mesh, False, 0.0, 500.0, scene.entries, np.zeros_like(tab), scene.lod_count)
    assert np.array_equal(none, wide)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["planar", "spherical"])
def test_cuda_minmax_cull_matches_oracle_in_order(ctx, kind):
    from bevy_terrain_b200 import api
    scene = Scene(kind, lod_count=3, T=16, B=2, src=32, grid_size=16)
    if kind == "planar":
        views = [((10.0, 170.0, -30.0), (300.0, 100.0, 200.0)), ((-500.0, 400.0, 100.0), (0.0, 150.0, 0.0))]
        hmin, hmax = 0.0, 500.0
    else:
        R = 6371000.0
        views = [((-(R + 20000.0), 0.0, 0.0), (-R, 50000.0, 80000.0)), ((0.0, R + 8000.0, 1000.0), (90000.0, R - 2000.0, 40000.0))]
        hmin, hmax = -12000.0, 9000.0
    atlas = api.TileAtlas(ctx, scene.terrain_config())
    tree = api.TileTree(atlas, scene.cfg)
    try:
        for pos, target in views:
            for tc, idx in scene.move_view(pos):
                atlas.upload_tile(0, idx, scene.atlas[idx])
                atlas.update_minmax(idx)
            tab = _minmax_table(scene)
            got_tab = atlas.minmax_table()
            loaded = [i for i in range(scene.atlas_size) if scene.atlas[i].any()]
            assert np.array_equal(got_tab[loaded], tab[loaded])                 # the device table is the oracle's
            tab[[i for i in range(scene.atlas_size) if i not in loaded]] = 0     # (never-updated slots are zero on the device)
            m = look_at_perspective(pos, target, near=1.0, far=1e8)
            frame = scene.view_frame()
            frame.culling = _cull_uniform(scene, m)
            tree.prepare(frame=frame)
            tree.select_tiles(flags=F.BT_FLAG_MINMAX_CULL)
            tiles, ind = tree.readback()
            exp, exp_ind, _ = O.select_tiles_minmax(scene.vc, scene.approx, frame.culling, scene.mesh, scene.spherical, hmin, hmax, scene.entries, tab, scene.lod_count)
            assert 0 < len(exp) and np.array_equal(tiles, exp)
            assert ind.vertex_count == exp_ind.vertex_count
            tree.select_tiles(flags=F.BT_FLAG_FRUSTUM_CULL)
            wide, _ = tree.readback()
            assert len(tiles) <= len(wide)
    finally:
        tree.close(); atlas.close()

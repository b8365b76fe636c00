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

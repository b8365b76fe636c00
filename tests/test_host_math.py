"""The product's C++ f64 host math (host_math.cpp, reached through the C ABI; no GPU needed) against
the oracle's independent C restatement and the Python clipmap reference."""
import ctypes as C

import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O
from tests.clipmap_ref import ClipmapRef
from tests.scene import planar_view_positions, spherical_view_positions


def _models():
    from bevy_terrain_b200 import api
    return [(api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0), O.terrain_model(F.BT_KIND_PLANAR, (0.0, -100.0, 0.0), 2000.0, 0.0, 500.0), planar_view_positions(6)),
            (api.TerrainModel.sphere((0.0, 0.0, 0.0), 6371000.0, -12000.0, 9000.0), O.terrain_model(F.BT_KIND_SPHERICAL, (0.0, 0.0, 0.0), 6371000.0, -12000.0, 9000.0), spherical_view_positions()),
            # WGS84-shaped ellipsoid (examples/spherical.rs:5-9 has MAJOR_AXES / MINOR_AXES of this size)
            (api.TerrainModel.ellipsoid((0.0, 0.0, 0.0), 6378137.0, 6356752.314245, -12000.0, 9000.0),
             O.terrain_model(F.BT_KIND_ELLIPSOIDAL, (0.0, 0.0, 0.0), 6378137.0, -12000.0, 9000.0, minor_axis=6356752.314245), spherical_view_positions())]


def test_constructors_and_uniforms_match():
    lib = F.load_library()
    for prod, orc, _ in _models():
        assert bytes(prod.c) == bytes(orc)
        assert prod.scale() == O.model_scale(orc)
        assert bytes(prod.mesh_transform()) == bytes(O.mesh_transform(orc))
        for g in (16, 64):
            cfg = O.default_view_config(grid_size=g)
            u = F.ViewConfigUniform()
            lib.bt_view_config_uniform_from(C.byref(cfg), C.byref(prod.c), C.byref(u))
            assert bytes(u) == bytes(O.view_config_uniform(cfg, orc))
    d = F.TerrainViewConfig()
    lib.bt_terrain_view_config_default(C.byref(d))
    assert bytes(d) == bytes(O.default_view_config())


def test_model_approximation_matches_oracle_bitwise():
    lib = F.load_library()
    for prod, orc, positions in _models():
        for pos in positions:
            a = F.ModelApproximation()
            assert lib.bt_model_approximation_compute(C.byref(prod.c), (F.f64 * 3)(*pos), 10, 123.5, C.byref(a)) == F.BT_OK
            assert bytes(a) == bytes(O.model_approximation(orc, pos, 10, 123.5))


def test_taylor_approximation_reproduces_f64_positions():
    """the 2nd-order series (terrain_model.rs:263-360) evaluated like compute_relative_position must agree with
    the exact f64 position near the view: validates the derivative algebra, not just its transcription"""
    L = O.lib()
    orc = O.terrain_model(F.BT_KIND_SPHERICAL, (0.0, 0.0, 0.0), 6371000.0, -12000.0, 9000.0)
    pos = (-(6371000.0 + 3000.0), 12345.0, -54321.0)
    ap = O.model_approximation(orc, pos, 10, 0.0)
    side = C.c_uint32(); uv = (C.c_double * 2)()
    L.bto_coordinate_from_world_position(C.byref(orc), (C.c_double * 3)(*pos), C.byref(side), uv)
    sp = ap.sides[side.value]
    for ds, dt in ((1e-4, 0.0), (0.0, -2e-4), (3e-4, 2e-4)):
        s, t = ds, dt
        rel = [sp.c[k] + sp.c_s[k] * s + sp.c_t[k] * t + sp.c_ss[k] * s * s + sp.c_st[k] * s * t + sp.c_tt[k] * t * t for k in range(3)]
        exact = (C.c_double * 3)()
        L.bto_coordinate_world_position(C.byref(orc), side.value, (C.c_double * 2)(uv[0] + ds, uv[1] + dt), 0.0, exact)
        err = np.linalg.norm(np.array(rel) - (np.array(exact[:]) - np.array(pos)))
        assert err < 0.5, err   # metres, on a 6371 km sphere, ~2 km from the view


@pytest.mark.parametrize("which", [0, 1, 2])
def test_tile_tree_matches_python_reference_over_a_camera_path(which):
    lib = F.load_library()
    prod, orc, positions = _models()[which]
    lod_count, atlas_size = 5, 400
    cfg = O.default_view_config()
    sides = 6 if which else 1
    existing = [(s, l, x, y) for s in range(sides) for l in range(lod_count) for x in range(1 << l) for y in range(1 << l)
                if not (l == 4 and (x + y) % 7 == 0)]   # some tiles missing on disk
    ref = ClipmapRef(orc, cfg, lod_count, atlas_size, existing)
    tt = C.c_void_p()
    assert lib.bt_tile_tree_create(C.byref(prod.c), C.byref(cfg), lod_count, atlas_size, C.byref(tt)) == F.BT_OK
    ex = np.array(existing, dtype=np.uint32)
    assert lib.bt_tile_tree_add_existing(tt, ex.ctypes.data_as(C.POINTER(F.Tile)), len(ex)) == F.BT_OK
    n_entries = sides * lod_count * cfg.tree_size ** 2
    try:
        for i, pos in enumerate(positions + positions[::-1]):
            exp_loads = ref.update(pos)
            buf = (F.TileLoad * atlas_size)(); n = F.u32()
            assert lib.bt_tile_tree_update(tt, (F.f64 * 3)(*pos), buf, atlas_size, C.byref(n)) == F.BT_OK
            got = [((l.tile.side, l.tile.lod, l.tile.x, l.tile.y), l.atlas_index) for l in buf[: n.value]]
            assert got == exp_loads, f"frame {i}"
            for j, (tc, _) in enumerate(exp_loads):
                if j % 5 != 4:   # leave some tiles "still loading": best-tile lookup must fall back to parents
                    ref.mark_loaded(tc)
                    assert lib.bt_tile_tree_mark_loaded(tt, C.byref(F.Tile(*tc))) == F.BT_OK
            ent = (F.TileTreeEntry * n_entries)(); org = (F.u32 * (2 * sides * lod_count))()
            assert lib.bt_tile_tree_entries(tt, ent, n_entries, org, len(org)) == F.BT_OK
            assert np.array_equal(np.frombuffer(ent, dtype=np.uint32).reshape(-1, 2), ref.entries())
            assert np.array_equal(np.frombuffer(org, dtype=np.uint32), ref.origins.reshape(-1))
    finally:
        lib.bt_tile_tree_destroy(tt)


def test_from_world_position_round_trip():
    L = O.lib()
    orc = O.terrain_model(F.BT_KIND_SPHERICAL, (10.0, -5.0, 3.0), 1000.0, 0.0, 1.0)
    rng = np.random.default_rng(6)
    for _ in range(200):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        p = np.array([10.0, -5.0, 3.0]) + 1000.0 * d
        side = C.c_uint32(); uv = (C.c_double * 2)()
        L.bto_coordinate_from_world_position(C.byref(orc), (C.c_double * 3)(*p), C.byref(side), uv)
        back = (C.c_double * 3)()
        L.bto_coordinate_world_position(C.byref(orc), side.value, uv, 0.0, back)
        assert np.allclose(back[:], p, atol=1e-6)
        assert 0.0 <= uv[0] <= 1.0 and 0.0 <= uv[1] <= 1.0


# ---- ellipsoid model (src/math/ellipsoid.rs) ---------------------------------------------------------------------------------
def _project_numpy(major, minor, q):
    """independent formulation of the closest point on the ellipsoid x^2/a^2 + y^2/b^2 + z^2/a^2 = 1 (b = polar axis along y):
    Newton on the Lagrange multiplier  sum (e_i q_i / (e_i^2 + t))^2 = 1,  x_i = e_i^2 q_i / (e_i^2 + t)"""
    e = np.array([major, minor, major], dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    t = 0.0 if np.sum((q / e) ** 2) <= 1.0 else np.linalg.norm(q) * e.min() - e.min() ** 2
    lo = -e.min() ** 2 + 1e-9
    for _ in range(200):
        f = np.sum((e * q / (e * e + t)) ** 2) - 1.0
        df = -2.0 * np.sum((e * q) ** 2 / (e * e + t) ** 3)
        step = f / df
        t_new = max(t - step, (t + lo) / 2.0)
        if abs(t_new - t) <= 1e-15 * max(1.0, abs(t)):
            t = t_new
            break
        t = t_new
    return e * e * q / (e * e + t)


def test_ellipsoid_projection_is_the_closest_surface_point():
    L = O.lib()
    major, minor = 6378137.0, 6356752.314245
    rng = np.random.default_rng(12)
    queries = [rng.normal(size=3) * s for s in (major * 0.3, major, major * 1.001, major * 4.0) for _ in range(60)]
    queries += [np.array(v, dtype=np.float64) for v in ((major * 2, 0.0, 0.0), (0.0, minor * 3, 0.0), (0.0, 0.0, -major * 1.5), (1.0, 2.0, 3.0),
                                                         (major * 0.5, 0.0, major * 0.5), (0.0, -minor * 0.9, major * 0.2))]
    for q in queries:
        out = (C.c_double * 3)()
        L.bto_project_point_ellipsoid(major, minor, (C.c_double * 3)(*q), out)
        x = np.array(out[:])
        assert abs((x[0] / major) ** 2 + (x[1] / minor) ** 2 + (x[2] / major) ** 2 - 1.0) < 5e-9           # on the surface (queries near the centre are ill-conditioned)
        grad = np.array([x[0] / major ** 2, x[1] / minor ** 2, x[2] / major ** 2])
        d = q - x
        if np.linalg.norm(d) > 1e-3:                                                                     # q - x is parallel to the normal
            assert np.linalg.norm(np.cross(d / np.linalg.norm(d), grad / np.linalg.norm(grad))) < 1e-8
        if np.sum((q / np.array([major, minor, major])) ** 2) > 1e-2:      # (deep inside, the closest point is not unique enough for Newton)
            assert np.linalg.norm(x - _project_numpy(major, minor, q)) < 1e-4 * max(1.0, np.linalg.norm(d) / major)   # == the twin


def test_ellipsoid_model_local_positions_and_scale():
    from bevy_terrain_b200 import api
    L = O.lib()
    major, minor = 6378137.0, 6356752.314245
    orc = O.terrain_model(F.BT_KIND_ELLIPSOIDAL, (0.0, 0.0, 0.0), major, -12000.0, 9000.0, minor_axis=minor)
    prod = api.TerrainModel.ellipsoid((0.0, 0.0, 0.0), major, minor, -12000.0, 9000.0)
    assert bytes(prod.c) == bytes(orc) and prod.is_spherical()
    assert prod.scale() == O.model_scale(orc) == (major + minor) / 2.0
    rng = np.random.default_rng(3)
    for _ in range(100):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        p = d * major * rng.uniform(1.0, 1.5)
        side = C.c_uint32(); uv = (C.c_double * 2)()
        L.bto_coordinate_from_world_position(C.byref(orc), (C.c_double * 3)(*p), C.byref(side), uv)
        back = (C.c_double * 3)()
        L.bto_coordinate_world_position(C.byref(orc), side.value, uv, 0.0, back)       # the surface point below p
        b = np.array(back[:])
        assert abs((b[0] / major) ** 2 + (b[1] / minor) ** 2 + (b[2] / major) ** 2 - 1.0) < 1e-9
        foot = _project_numpy(major, minor, p)
        assert np.linalg.norm(b - foot) < 1e-3         # metres: it is the foot point of p on the ellipsoid


def test_two_views_of_one_terrain_share_the_atlas_state():
    """TileAtlasState is per terrain (tile_atlas.rs:272-300): two TileTrees (views) created on one bt_atlas_state never get the
    same atlas slot for different tiles, a tile requested by both is loaded once, and it stays resident until both released it."""
    lib = F.load_library()
    prod, orc, positions = _models()[0]
    lod_count, atlas_size = 5, 300
    cfg = O.default_view_config()
    existing = np.array([(0, l, x, y) for l in range(lod_count) for x in range(1 << l) for y in range(1 << l)], dtype=np.uint32)
    state, a, b = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert lib.bt_atlas_state_create(atlas_size, C.byref(state)) == F.BT_OK
    assert lib.bt_tile_tree_create_shared(C.byref(prod.c), C.byref(cfg), lod_count, state, C.byref(a)) == F.BT_OK
    assert lib.bt_tile_tree_create_shared(C.byref(prod.c), C.byref(cfg), lod_count, state, C.byref(b)) == F.BT_OK
    assert lib.bt_tile_tree_add_existing(a, existing.ctypes.data_as(C.POINTER(F.Tile)), len(existing)) == F.BT_OK   # shared: visible to b too
    slot_of = {}
    n_entries = lod_count * cfg.tree_size ** 2

    def update(tree, pos):
        buf = (F.TileLoad * atlas_size)(); n = F.u32()
        assert lib.bt_tile_tree_update(tree, (F.f64 * 3)(*pos), buf, atlas_size, C.byref(n)) == F.BT_OK
        for l in buf[: n.value]:
            key = (l.tile.side, l.tile.lod, l.tile.x, l.tile.y)
            assert key not in slot_of, "a tile both views want is loaded once"
            slot_of[key] = l.atlas_index
            assert lib.bt_tile_tree_mark_loaded(tree, C.byref(l.tile)) == F.BT_OK
        return n.value

    try:
        pa, pb = positions[0], positions[-1]
        assert update(a, pa) > 0
        update(b, pb)
        assert len(set(slot_of.values())) == len(slot_of), "two views must never be handed the same atlas slot"
        ent_a = (F.TileTreeEntry * n_entries)(); ent_b = (F.TileTreeEntry * n_entries)()
        assert lib.bt_tile_tree_entries(a, ent_a, n_entries, None, 0) == F.BT_OK
        assert lib.bt_tile_tree_entries(b, ent_b, n_entries, None, 0) == F.BT_OK
        ea, eb = (np.frombuffer(e, dtype=np.uint32).reshape(-1, 2) for e in (ent_a, ent_b))
        assert ea[0, 0] != 0xFFFFFFFF and eb[0, 0] == ea[0, 0]       # the root tile: one slot, seen by both views
        valid_b = int((eb[:, 0] != 0xFFFFFFFF).sum())
        assert valid_b > lod_count
        # moving view b onto view a's position requests nothing new: the tiles are resident through view a
        before = len(slot_of)
        update(b, pa)
        assert len(slot_of) == before
        # destroying view a releases its requests; b still holds the same tiles, so they must not be recycled
        lib.bt_tile_tree_destroy(a); a = C.c_void_p()
        update(b, pa)
        assert lib.bt_tile_tree_entries(b, ent_b, n_entries, None, 0) == F.BT_OK
        after = np.frombuffer(ent_b, dtype=np.uint32).reshape(-1, 2)
        assert int((after[:, 0] != 0xFFFFFFFF).sum()) >= valid_b and after[0, 0] == ea[0, 0]
    finally:
        if a: lib.bt_tile_tree_destroy(a)
        lib.bt_tile_tree_destroy(b)
        lib.bt_atlas_state_destroy(state)

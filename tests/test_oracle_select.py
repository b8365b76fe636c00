"""Structural invariants of the oracle's select — the stand-in for missing goldens."""
import numpy as np

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O
from tests.scene import Scene, planar_view_positions, spherical_view_positions


def _disjoint_cover(tiles, sides):
    # every final tile is a leaf: no tile is an ancestor of another, and areas sum to the face count
    keys = {tuple(t) for t in tiles.tolist()}
    assert len(keys) == len(tiles)
    for s, l, x, y in keys:
        for up in range(1, l + 1):
            assert (s, l - up, x >> up, y >> up) not in keys
    assert abs(sum(4.0 ** -int(l) for l in tiles[:, 1]) - sides) < 1e-9


def test_planar_tiles_cover_the_domain_exactly():
    scene = Scene("planar", lod_count=4, T=16, B=2, src=32, grid_size=16)
    for pos in planar_view_positions(8):
        scene.move_view(pos)
        tiles, ind, res = scene.oracle_select()
        assert res.dropped_parents == 0 and not res.overflow
        _disjoint_cover(tiles, 1)
        assert ind.vertex_count == scene.vc.vertices_per_tile * len(tiles) and ind.instance_count == 1


def test_spherical_tiles_cover_all_faces():
    scene = Scene("spherical", lod_count=3, T=16, B=2, src=32, grid_size=16)
    for pos in spherical_view_positions():
        scene.move_view(pos)
        tiles, _, res = scene.oracle_select()
        assert res.dropped_parents == 0
        _disjoint_cover(tiles, 6)


def test_lod_ring_population_matches_survey_simulation():
    """planar side 2000, camera 20 above the surface -> ~4.8k tiles, ~830 per LOD ring"""
    scene = Scene("planar", lod_count=4, T=16, B=2, src=32, grid_size=64)
    scene.move_view((100.0, -100.0 + 250.0 + 20.0, -50.0))
    tiles, _, res = scene.oracle_select()
    assert 4500 < res.final_count < 5100
    rings = np.bincount(tiles[:, 1])
    assert all(700 < c < 900 for c in rings[6:10])


def test_divide_decision_agrees_with_f64_tile_distance():
    """cross-check of the GPU formulation (should_be_divided, f32) against the reference's independent CPU
    formulation compute_tile_distance (tile_tree.rs:193-221, f64): away from the threshold both must agree"""
    import ctypes as C
    scene = Scene("planar", lod_count=4, T=16, B=2, src=32, grid_size=16)
    pos = (321.0, 180.0, -77.0)
    scene.move_view(pos)
    tiles, _, _ = scene.oracle_select()
    L = O.lib()
    vp = (C.c_double * 3)(*pos)
    side = C.c_uint32(); uv = (C.c_double * 2)()
    L.bto_coordinate_from_world_position(C.byref(scene.model), vp, C.byref(side), uv)
    sd = scene.vc.subdivision_distance
    checked = 0
    for s, l, x, y in tiles.tolist():
        t = F.Tile(s, l, x, y)
        d = L.bto_tile_distance(C.byref(scene.model), C.byref(t), 0, uv, scene.approximate_height, vp)
        thr = sd / (1 << l)
        if abs(d - thr) > 1e-3 * thr:
            assert d >= thr, (s, l, x, y, d, thr)   # a final tile did not want to be divided
            checked += 1
        if l > 0:  # its parent wanted to be divided
            p = F.Tile(s, l - 1, x >> 1, y >> 1)
            dp = L.bto_tile_distance(C.byref(scene.model), C.byref(p), 0, uv, scene.approximate_height, vp)
            thrp = sd / (1 << (l - 1))
            if abs(dp - thrp) > 1e-3 * thrp:
                assert dp < thrp
    assert checked > 100


def test_capacity_overflow_flag():
    scene = Scene("planar", lod_count=4, T=16, B=2, src=32, grid_size=16)
    scene.move_view((0.0, 151.0, 0.0))
    tiles, ind, res = scene.oracle_select(capacity=100)
    assert res.overflow == 1 and res.final_count == 100

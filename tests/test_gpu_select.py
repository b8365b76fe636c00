"""GPU parity: bt_select_tiles vs the oracle's sequential execution of prepare_root / refine_tiles /
prepare_next / prepare_render.  Tile lists must match BIT-EXACTLY AND IN ORDER (the CUDA compaction
reproduces the order a sequential execution of the reference shaders yields)."""
import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests.scene import Scene, planar_view_positions, spherical_view_positions

pytestmark = pytest.mark.gpu


def _run_select(ctx, scene, positions, geometry_tile_count=None):
    from bevy_terrain_b200 import api
    cfg = scene.terrain_config()
    atlas = api.TileAtlas(ctx, cfg)
    vcfg = scene.cfg
    if geometry_tile_count is not None:
        vcfg.geometry_tile_count = geometry_tile_count
        scene.vc.geometry_tile_count = geometry_tile_count
    tree = api.TileTree(atlas, vcfg)
    results = []
    try:
        for pos in positions:
            scene.move_view(pos)
            tree.prepare(frame=scene.view_frame())
            tree.select_tiles()
            try:
                tiles, ind = tree.readback()
                status = F.BT_OK
            except F.BtError as e:
                status = e.status
                tiles, ind = None, None
            results.append((tiles, ind, status, tree.select_status() if status == F.BT_OK else None))
    finally:
        tree.close()
        atlas.close()
    return results


@pytest.mark.parametrize("grid_size", [16, 64])
def test_select_planar_matches_oracle_in_order(ctx, grid_size):
    scene = Scene("planar", lod_count=4, T=16, B=2, src=32, grid_size=grid_size)
    positions = planar_view_positions(8)
    got = _run_select(ctx, scene, positions)
    for pos, (tiles, ind, status, dropped) in zip(positions, got):
        scene.move_view(pos)
        exp_tiles, exp_ind, res = scene.oracle_select()
        assert status == F.BT_OK
        assert len(tiles) == res.final_count
        assert np.array_equal(tiles, exp_tiles), f"tile list differs at view {pos}"
        assert (ind.vertex_count, ind.instance_count, ind.base_vertex, ind.base_instance) == (exp_ind.vertex_count, 1, 0, 0)
        assert dropped == res.dropped_parents


def test_select_spherical_matches_oracle_in_order(ctx):
    scene = Scene("spherical", lod_count=3, T=16, B=2, src=32, grid_size=16)
    positions = spherical_view_positions()
    got = _run_select(ctx, scene, positions)
    for pos, (tiles, ind, status, dropped) in zip(positions, got):
        scene.move_view(pos)
        exp_tiles, exp_ind, res = scene.oracle_select()
        assert status == F.BT_OK
        assert np.array_equal(tiles, exp_tiles), f"tile list differs at view {pos}"
        assert ind.vertex_count == exp_ind.vertex_count
        # structural invariant: no parent dropped => the final tiles tile every face exactly
        if res.dropped_parents == 0:
            area = sum(4.0 ** -int(l) for l in tiles[:, 1])
            assert abs(area - 6.0) < 1e-9


def test_select_few_refinements_drops_parents(ctx):
    """Q2: parents still wanting division after the last pass are subdivided into temp and dropped."""
    scene = Scene("planar", lod_count=4, T=16, B=2, src=32, grid_size=16, refinement_count=5)
    pos = (0.0, 151.0, 0.0)
    (tiles, ind, status, dropped), = _run_select(ctx, scene, [pos])
    scene.move_view(pos)
    exp_tiles, _, res = scene.oracle_select()
    assert res.dropped_parents > 0
    assert dropped == res.dropped_parents
    assert np.array_equal(tiles, exp_tiles)
    assert sum(4.0 ** -int(l) for l in tiles[:, 1]) < 1.0  # holes, as in the reference


def test_select_capacity_overflow_is_reported(ctx):
    """geometry_tile_count overflow is unchecked in the reference; here it is BT_ERR_CAPACITY."""
    scene = Scene("planar", lod_count=4, T=16, B=2, src=32, grid_size=16)
    (tiles, ind, status, _), = _run_select(ctx, scene, [(0.0, 151.0, 0.0)], geometry_tile_count=1000)
    assert status == F.BT_ERR_CAPACITY


def test_select_spills_to_global_frontier(ctx):
    """a large morph_distance pushes frontiers beyond the 4096-tile shared-memory buffers"""
    scene = Scene("planar", lod_count=4, T=16, B=2, src=32, grid_size=16, morph_distance=48.0)
    pos = (10.0, 170.0, -20.0)
    (tiles, ind, status, dropped), = _run_select(ctx, scene, [pos])
    scene.move_view(pos)
    exp_tiles, _, res = scene.oracle_select()
    assert res.max_frontier > 4096
    assert status == F.BT_OK
    assert np.array_equal(tiles, exp_tiles)


@pytest.mark.parametrize("kind", ["planar", "spherical"])
def test_select_batch_equals_per_view_and_oracle(ctx, kind):
    """bt_select_tiles_batch: the views of a frame in ONE launch (one CTA per view) give exactly the per-view results, in order"""
    from bevy_terrain_b200 import api
    scene = Scene(kind, lod_count=3 if kind == "spherical" else 4, T=16, B=2, src=32, grid_size=16)
    positions = (spherical_view_positions() if kind == "spherical" else planar_view_positions(6))[:6]
    atlas = api.TileAtlas(ctx, scene.terrain_config())
    trees = [api.TileTree(atlas, scene.cfg) for _ in positions]
    batch = None
    try:
        for tree, pos in zip(trees, positions):
            scene.move_view(pos)
            tree.prepare(frame=scene.view_frame())
        batch = api.ViewBatch(trees)
        batch.select_tiles()
        for tree, pos in zip(trees, positions):
            scene.move_view(pos)
            exp_tiles, exp_ind, res = scene.oracle_select()
            tiles, ind = tree.readback()
            assert np.array_equal(tiles, exp_tiles), f"batched tile list differs at view {pos}"
            assert ind.vertex_count == exp_ind.vertex_count and tree.select_status() == res.dropped_parents
            tree.select_tiles()                         # the single-view launch of the same kernel
            again, _ = tree.readback()
            assert np.array_equal(again, tiles)
    finally:
        if batch:
            batch.close()
        for tree in trees:
            tree.close()
        atlas.close()

"""TileTree::approximate_height chain: host lookup (bt_tile_tree_sample_lookup) + device texels (bt_atlas_sample) against a
numpy restatement of sample_attachment / compute_blend / lookup_tile / AttachmentData::sample
(terrain_data/mod.rs:221-307, tile_tree.rs:223-266, tile_atlas.rs:249-258)."""

import numpy as np
import pytest

from tests.scene import Scene

pytestmark = pytest.mark.gpu
f32 = np.float32


from tests.twin import twin_sample  # noqa: E402


def test_sample_attachment_matches_cpu_twin(ctx):
    from bevy_terrain_b200 import api
    scene = Scene("planar", lod_count=4, T=64, B=2, src=256, grid_size=16)
    atlas = api.TileAtlas(ctx, scene.terrain_config())
    tree = api.TileTree(atlas, scene.cfg)
    try:
        tree.add_existing_tiles(np.array(sorted(scene.existing), dtype=np.uint32))
        rng = np.random.default_rng(4)
        seen_blend = False
        for pos in [(100.0, 170.0, -50.0), (-400.0, 400.0, 300.0), (0.0, 1500.0, 0.0), (700.0, 155.0, -650.0)]:
            loads = scene.move_view(pos)
            for tc, idx in loads:
                atlas.upload_tile(0, idx, scene.atlas[idx])
            for tile, atlas_index in tree.compute_requests(pos):
                tree.mark_loaded(tile)
            tree.adjust_to_tile_atlas()
            entries = np.frombuffer(tree.entries, dtype=np.uint32).reshape(-1, 2)
            assert np.array_equal(entries, scene.entries)
            for _ in range(12):
                sample_pos = (pos[0] + float(rng.uniform(-300, 300)), pos[1], pos[2] + float(rng.uniform(-300, 300)))
                sample_pos = tuple(float(np.clip(v, -990, 990)) if k != 1 else v for k, v in enumerate(sample_pos))
                got = tree.sample_attachment(0, sample_pos)
                exp, lod, ratio = twin_sample(scene, scene.entries, scene.cfg.tree_size, sample_pos, tree.approximate_height,
                                              scene.cfg.blend_distance * 1000.0, scene.cfg.blend_range, 4)
                seen_blend |= ratio > 0
                assert got.tobytes() == exp.tobytes(), (sample_pos, got, exp, lod, ratio)
            h = tree.update_approximate_height()      # TileTree::approximate_height: feeds the next frame's clipmap + Taylor series
            scene.clipmap.approximate_height = h      # keep the reference clipmap in step
            assert 0.0 <= h <= 500.0
        assert seen_blend
    finally:
        tree.close(); atlas.close()

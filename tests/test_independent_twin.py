"""The C oracle against a SECOND restatement of the reference's shaders (tests/independent_twin.py: numpy, written from the WGSL
without reading oracle/*.c).  Byte-exact pyramids (R16 + Rgba8, planar + cube, holes) and order-exact tile lists: a misreading of
the shaders would have to be made twice, in two unrelated formulations, to survive."""
import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from tests import independent_twin as TW
from tests import oracle as O
from tests.scene import Scene, planar_view_positions, spherical_view_positions


@pytest.mark.parametrize("fmt", [F.BT_FORMAT_R16, F.BT_FORMAT_RGBA8])
@pytest.mark.parametrize("kind", ["planar", "cube"])
def test_pyramid_matches_the_independent_twin(fmt, kind):
    T, B, L = 20, 2, 3
    spherical = kind == "cube"
    sides = range(6) if spherical else range(1)
    size = 50 if spherical else 70            # not a multiple of the tile grid: the split resamples at a fractional ratio
    sources = {s: O.synth_source(fmt, size, size, seed=77 + 13 * s, holes=True) for s in sides}
    for s in sources:                          # make sure every validity branch is exercised: a zero block and a zero-red speckle
        sources[s][5:9, 11:17] = 0
        sources[s][size // 2, ::7] = 0
    pyr = O.Pyramid(F.BT_KIND_SPHERICAL if spherical else F.BT_KIND_PLANAR, fmt, T, B, L)
    pyr.add_datasets([O.make_dataset(sources[s], side=s, lod_range=(0, L)) for s in sides])
    pyr.run()
    twin = TW.build_pyramid(sources, T, B, L, spherical)
    got = {tuple(int(v) for v in t) for t in pyr.tiles()}
    assert got == set(twin)
    for tc in sorted(got):
        a, b = pyr.tile(*tc), twin[tc]
        assert a.dtype == b.dtype and a.shape == b.shape
        assert np.array_equal(a, b), f"tile {tc}: {int((a != b).sum())} texel components differ (first at {np.argwhere(a != b)[0]})"
    pyr.close()


def test_split_overlay_and_partial_coverage_match_the_twin():
    """a dataset that covers part of the terrain (top_left / bottom_right) laid over an earlier one: the fall-through keeps the earlier texels"""
    T, B, lod = 24, 2, 2
    base = O.synth_source(F.BT_FORMAT_R16, 64, 64, seed=5)
    over = O.synth_source(F.BT_FORMAT_R16, 40, 40, seed=6, holes=True)
    tl, br = (0.25, 0.5), (0.75, 1.0)
    pyr = O.Pyramid(F.BT_KIND_PLANAR, F.BT_FORMAT_R16, T, B, lod + 1)
    pyr.add_datasets([O.make_dataset(base, lod_range=(lod, lod + 1))])
    pyr.add_datasets([O.make_dataset(over, top_left=tl, bottom_right=br, lod_range=(lod, lod + 1))])
    pyr.run_phases = getattr(pyr, "run_phases", None)
    pyr.run(phases=1)                                      # split (+ downsample: none at a single LOD); no stitch
    n = 1 << lod
    for x in range(n):
        for y in range(n):
            first = TW.split_tile(base, (x, y), lod, T, B)
            lo = (np.floor(np.float32(tl[0]) * n), np.floor(np.float32(tl[1]) * n))
            hi = (np.ceil(np.float32(br[0]) * n), np.ceil(np.float32(br[1]) * n))
            expect = first
            if lo[0] <= x < hi[0] and lo[1] <= y < hi[1]:   # PreprocessDataset::overlapping_tiles
                expect = TW.split_tile(over, (x, y), lod, T, B, top_left=tl, bottom_right=br, existing=first)
            assert np.array_equal(pyr.tile(0, lod, x, y), expect), (x, y)
    pyr.close()


@pytest.mark.parametrize("kind", ["planar", "spherical"])
def test_select_matches_the_independent_twin_in_order(kind):
    scene = Scene(kind, lod_count=3, T=16, B=2, src=32, grid_size=16)
    positions = (spherical_view_positions() if kind == "spherical" else planar_view_positions(6))[:5]
    for pos in positions:
        scene.move_view(pos)
        exp_tiles, _, res = scene.oracle_select()
        got, dropped = TW.select_tiles(scene.vc, scene.approx, [np.float32(v) for v in scene.view_pos], scene.mesh, scene.spherical)
        assert [tuple(int(v) for v in t) for t in exp_tiles] == got, f"tile list differs at {pos}"
        assert dropped == res.dropped_parents

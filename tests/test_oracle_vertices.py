"""Self-consistency of the vertex path (checks that stand in for the missing reference goldens), on the ORACLE;
the CUDA path is bit-identical to it (tests/test_gpu_vertices.py), so the properties carry over.

* vertices shared by adjacent tiles of equal LOD coincide bit-exactly (no cracks inside one LOD ring);
* with MORPH, a vertex whose morph ratio is 1 sits exactly on its even (parent-grid) neighbour;
* across a cube-face edge the two faces' edge vertices agree to f32 rounding of the 6 371 km radius."""
import numpy as np

from bevy_terrain_b200 import _ffi as F
from tests.scene import Scene


def _grid(verts, n_tiles, g):
    """strip-ordered vertices -> unique grid [tile][column 0..g][row 0..g]"""
    vpr = 2 * (g + 2)
    v = verts.reshape(n_tiles, g, vpr)
    left = v[:, :, 1:-1:2][:, :, :g + 1]      # (column c, rows 0..g)
    right = v[:, :, 2::2][:, :, :g + 1]       # (column c + 1, rows 0..g)
    return np.concatenate([left, right[:, -1:]], axis=1)   # columns 0..g


def test_equal_lod_neighbours_share_their_edge_vertices_bit_exactly():
    g = 16
    scene = Scene("planar", lod_count=4, T=64, B=2, src=256, grid_size=g)
    scene.move_view((100.0, 170.0, -50.0))
    tiles, ind, _ = scene.oracle_select()
    flags = F.BT_FLAG_NORMALS                      # no morph: the grid position is the plain tile uv
    grid = _grid(scene.oracle_vertices(tiles, flags), len(tiles), g)
    index = {tuple(int(v) for v in t): i for i, t in enumerate(tiles)}
    pairs = 0
    for (side, lod, x, y), i in index.items():
        last = (1 << lod) - 1
        # (tiles on the terrain boundary are left out: where a DIAGONAL neighbour does not exist the reference's stitch repeats
        #  the tile's own corner texel (stitch.wgsl:98-103) instead of the side neighbour's, so the outermost rows differ there)
        j = index.get((side, lod, x + 1, y))
        if j is not None and 0 < y < last:         # right neighbour at the same LOD: my column g == its column 0
            assert np.array_equal(grid[i, g]["position"], grid[j, 0]["position"]), (lod, x, y)
            pairs += 1
        j = index.get((side, lod, x, y + 1))
        if j is not None and 0 < x < last:         # lower neighbour: my row g == its row 0
            assert np.array_equal(grid[i, :, g]["position"], grid[j, :, 0]["position"]), (lod, x, y)
            pairs += 1
    assert pairs > 500


def test_fully_morphed_vertices_sit_on_the_parent_grid():
    g = 16
    scene = Scene("planar", lod_count=4, T=64, B=2, src=256, grid_size=g)
    scene.move_view((100.0, 170.0, -50.0))
    tiles, ind, _ = scene.oracle_select()
    grid = _grid(scene.oracle_vertices(tiles, F.BT_FLAG_MORPH), len(tiles), g)
    uv = grid["uv"]                                 # morphed tile uv
    plain = (np.arange(g + 1, dtype=np.float32) / np.float32(g))
    odd, even = slice(1, g, 2), slice(0, g - 1, 2)                  # odd columns 1, 3, .. and their even left neighbours
    target = plain[even][None, :, None]
    # odd grid vertices whose morphed uv equals the even neighbour's uv (morph ratio == 1)
    snapped = (uv[:, odd, :, 0] == target) & (uv[:, even, :, 0] == target) & (uv[:, odd, :, 1] == uv[:, even, :, 1])
    a, b = grid["position"][:, odd][snapped], grid["position"][:, even][snapped]
    assert snapped.sum() > 100, snapped.sum()                        # the far rings of the frame are fully morphed
    assert np.array_equal(a[:, 0], b[:, 0]) and np.array_equal(a[:, 2], b[:, 2])      # horizontally the two vertices are ONE point
    # the height is sampled at the LOD the UNMORPHED position asks for (vertex.wgsl:54-70: blend from the pre-morph distance),
    # so it is the same bits unless the two vertices straddle a sampling-LOD boundary
    assert (a[:, 1] == b[:, 1]).mean() > 0.9


def test_cube_face_edge_vertices_agree_across_faces():
    from scipy.spatial import cKDTree
    g = 16
    scene = Scene("spherical", lod_count=3, T=32, B=2, src=64, grid_size=g)
    R = 6371000.0
    d = np.array([1.0, 0.02, 1.0]); d /= np.linalg.norm(d)          # above the edge between two cube faces
    scene.move_view(tuple(d * (R + 200000.0)))
    tiles, ind, _ = scene.oracle_select()
    grid = _grid(scene.oracle_vertices(tiles, 0), len(tiles), g)
    pos = grid["position"][..., :3].astype(np.float64)
    # the height displaces a vertex along its own direction, so the DIRECTION is the horizontal position on the sphere
    direction = pos / np.linalg.norm(pos, axis=-1, keepdims=True)
    edges = np.concatenate([direction[:, 0], direction[:, g], direction[:, :, 0], direction[:, :, g]], axis=1)   # [tile][4 (g+1)][3]
    sides = np.repeat(tiles[:, 0], edges.shape[1])
    pts = edges.reshape(-1, 3) * R
    shared = 0
    for s in np.unique(sides):
        mine, other = pts[sides == s], pts[sides != s]
        dist, _ = cKDTree(other).query(mine, distance_upper_bound=50.0)      # inf when nothing lies within 50 m
        same = dist < 50.0                      # a vertex of another face within 50 m on the ground: the SAME edge vertex ...
        assert (dist[same] < 4.0).all(), dist[same].max()      # ... and it agrees to a few ulps of 6.4e6 (0.5 m each)
        shared += int(same.sum())
    assert shared > 40


def test_ellipsoidal_terrain_vertices_lie_on_the_ellipsoid_shell():
    """TerrainModel::ellipsoid (terrain_model.rs:106-128): the mesh transform carries the non-uniform scale (major, minor, major),
    so every generated vertex sits between the ellipsoid displaced by min_height and by max_height along its normal"""
    g = 16
    scene = Scene("ellipsoid", lod_count=3, T=32, B=2, src=64, grid_size=g)
    a, b = 6378137.0, 6356752.314245
    d = np.array([0.3, 0.9, -0.4]); d /= np.linalg.norm(d)          # high latitude: the flattening matters
    scene.move_view(tuple(d * (a + 400000.0)))
    tiles, ind, _ = scene.oracle_select()
    assert len({int(t[0]) for t in tiles}) >= 2
    v = scene.oracle_vertices(tiles, F.BT_FLAG_NORMALS | F.BT_FLAG_MORPH | F.BT_FLAG_BLEND)
    p = v["position"][:, :3].astype(np.float64)
    # height above the ellipsoid along the (approximately radial) normal: solve |p / (a+h, b+h, a+h)| = 1 to first order
    level = np.sqrt((p[:, 0] / a) ** 2 + (p[:, 1] / b) ** 2 + (p[:, 2] / a) ** 2)
    h = (level - 1.0) * np.linalg.norm(p, axis=1) / level
    assert h.min() > -12000.0 - 50.0 and h.max() < 9000.0 + 50.0
    # the same frame on the SPHERE of the mean radius differs visibly near the pole (sanity: the ellipsoid is not ignored)
    polar = np.abs(p[:, 1]) > 0.95 * np.linalg.norm(p, axis=1)
    if polar.any():
        assert np.linalg.norm(p[polar], axis=1).max() < b + 9000.0 + 3000.0     # polar radius, not the 6 371 km mean

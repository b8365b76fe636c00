"""Helpers shared by the full-size GPU tests (TEST INFRASTRUCTURE): structure checks of the strip-ordered vertex stream and the
bit-exact comparison of whole tiles with the oracle, fed with the product's own inputs (atlas + tile-tree table downloaded)."""
import ctypes as C

import numpy as np

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O


def check_strip_structure(tree, tile_index: int, grid: int = 64):
    vpt = 2 * (grid + 2) * grid
    v = tree.download_vertices(tile_index * vpt, vpt).reshape(grid, 2 * (grid + 2))
    raw = np.ascontiguousarray(v).view(np.uint8).reshape(grid, 2 * (grid + 2), 40)
    assert np.array_equal(raw[:, 0], raw[:, 1]) and np.array_equal(raw[:, -1], raw[:, -2])       # duplicated strip ends
    assert np.array_equal(v[:-1, 2::2]["position"][:, :grid + 1], v[1:, 1::2]["position"][:, :grid + 1])   # shared columns
    assert (v["tile_index"] == tile_index).all()
    return v


def oracle_inputs(ctx, cfg, atlas, tree, tiles, flags, pos, spherical: bool, T: int, B: int, min_h: float, max_h: float):
    host_atlas = np.empty((cfg.atlas_size, T, T), dtype=np.uint16)
    ctx.d2h(host_atlas, atlas.device_ptr(0))
    vi = O.VertexInputs()
    vi.view_config, vi.approximation = tree.frame.view_config, tree.frame.approximation
    vi.terrain_config = F.TerrainConfigUniform(cfg.lod_count, min_h, max_h, float(np.float32(cfg.model.scale())))
    Cn = T - 2 * B
    vi.attachments[0].size, vi.attachments[0].scale, vi.attachments[0].offset = float(Cn), float(np.float32(Cn) / np.float32(T)), float(np.float32(B) / np.float32(T))
    vi.mesh = cfg.model.mesh_transform()
    vi.view_world_position[:] = [float(np.float32(p)) for p in pos]
    vi.spherical, vi.texture_size, vi.atlas_size = int(spherical), T, cfg.atlas_size
    entries = np.frombuffer(tree.entries, dtype=np.uint32).reshape(-1, 2).copy()
    vi.tile_tree = entries.ctypes.data_as(C.POINTER(F.TileTreeEntry))
    vi.height_atlas = host_atlas.ctypes.data
    vi.tiles = tiles.ctypes.data_as(C.POINTER(F.Tile))
    vi.tile_count, vi.flags = len(tiles), flags
    vi._keep = (host_atlas, entries, tiles)
    return vi


def assert_tiles_bit_exact(tree, vi, tile_indices, grid: int = 64):
    vpt = 2 * (grid + 2) * grid
    for t in tile_indices:
        exp = O.gen_vertices(vi, int(t) * vpt, vpt)
        got = tree.download_vertices(int(t) * vpt, vpt)
        assert got.tobytes() == exp.tobytes(), f"tile {t} differs"

"""helpers that rebuild oracle / C-ABI inputs from the committed golden fixtures (tests/golden/*.npz)"""
import ctypes as C
import os

import numpy as np

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def struct_from(cls, arr):
    s = cls()
    C.memmove(C.byref(s), arr.tobytes(), C.sizeof(cls))
    return s


class GoldenFrame:
    def __init__(self, g, i, spherical):
        self.spherical = spherical
        self.lod_count, self.T, self.B = int(g["lod_count"]), int(g["T"]), int(g["B"])
        self.vc = struct_from(F.ViewConfigUniform, g[f"v{i}_view_config"])
        self.approx = struct_from(F.ModelApproximation, g[f"v{i}_approximation"])
        self.cull = struct_from(F.CullingUniform, g[f"v{i}_culling"])
        self.mesh = struct_from(F.MeshTransform, g[f"v{i}_mesh"])
        self.view_pos = tuple(g[f"v{i}_view_pos"])
        self.entries = np.ascontiguousarray(g[f"v{i}_entries"])
        self.atlas = np.ascontiguousarray(g[f"v{i}_atlas"])
        self.tiles = np.ascontiguousarray(g[f"v{i}_tiles"])
        self.indirect = g[f"v{i}_indirect"]
        self.flags = int(g[f"v{i}_flags"])
        self.vertex_index = g[f"v{i}_vertex_index"]
        self.vertices = g[f"v{i}_vertices"]

    def min_max_scale(self):
        return (-12000.0, 9000.0, 6371000.0) if self.spherical else (0.0, 500.0, 1000.0)

    def oracle_vertex_inputs(self):
        vi = O.VertexInputs()
        vi.view_config, vi.approximation, vi.mesh = self.vc, self.approx, self.mesh
        mn, mx, sc = self.min_max_scale()
        vi.terrain_config = F.TerrainConfigUniform(self.lod_count, mn, mx, float(np.float32(sc)))
        c = np.float32(self.T - 2 * self.B)
        vi.attachments[0].size, vi.attachments[0].scale = float(c), float(c / np.float32(self.T))
        vi.attachments[0].offset = float(np.float32(self.B) / np.float32(self.T))
        vi.view_world_position[:] = [float(np.float32(v)) for v in self.view_pos]
        vi.spherical, vi.texture_size = int(self.spherical), self.T
        vi.atlas_size = len(self.atlas)
        vi.tile_tree = self.entries.ctypes.data_as(C.POINTER(F.TileTreeEntry))
        vi.height_atlas = self.atlas.ctypes.data
        vi.tiles = self.tiles.ctypes.data_as(C.POINTER(F.Tile))
        vi.tile_count, vi.flags = len(self.tiles), self.flags
        return vi

    def view_frame(self):
        f = F.ViewFrame()
        f.view_config, f.approximation, f.culling = self.vc, self.approx, self.cull
        f.view_world_position[:] = [float(np.float32(v)) for v in self.view_pos]
        f.tile_tree_entry_count = len(self.entries)
        f.tile_tree = self.entries.ctypes.data_as(C.POINTER(F.TileTreeEntry))
        return f

"""Shared scenario builder for the parity tests: synthetic sources -> ORACLE pyramid -> clipmap table
-> per-frame inputs, in the byte formats both the oracle and the C ABI consume.  TEST INFRASTRUCTURE.

The scenarios are scaled-down versions of BASELINE.json's configs: same model
constants and view defaults as the reference examples, smaller textures so the CPU oracle
finishes in seconds.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O
from tests.clipmap_ref import ClipmapRef


class Scene:
    def __init__(self, kind="planar", lod_count=4, T=64, B=2, src=256, grid_size=16, atlas_size=None, seed=0x5EED0001,
                 holes=False, **view_overrides):
        self.kind = F.BT_KIND_PLANAR if kind == "planar" else F.BT_KIND_SPHERICAL     # what terrains / pyramids / kernels see
        self.spherical = self.kind == F.BT_KIND_SPHERICAL
        self.lod_count, self.T, self.B, self.src_size = lod_count, T, B, src
        if kind == "ellipsoid":   # examples/spherical.rs:5-9: MAJOR_AXES / MINOR_AXES (WGS84), an ellipsoid is "spherical" on the GPU
            self.model = O.terrain_model(F.BT_KIND_ELLIPSOIDAL, (0.0, 0.0, 0.0), 6378137.0, -12000.0, 9000.0, minor_axis=6356752.314245)
        elif self.spherical:      # examples/spherical.rs:5-9 constants
            self.model = O.terrain_model(self.kind, (0.0, 0.0, 0.0), 6371000.0, -12000.0, 9000.0)
        else:                # examples/planar.rs:6-9,53
            self.model = O.terrain_model(self.kind, (0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
        self.sides = 6 if self.spherical else 1
        self.cfg = O.default_view_config(grid_size=grid_size, **view_overrides)
        self.vc = O.view_config_uniform(self.cfg, self.model)
        self.mesh = O.mesh_transform(self.model)
        self.atlas_size = atlas_size or self.sides * lod_count * self.cfg.tree_size ** 2
        # sources + oracle pyramid (height only, R16)
        self.sources = [O.synth_source(F.BT_FORMAT_R16, src, src, seed=seed + 17 * s, holes=holes) for s in range(self.sides)]
        self.pyramid = O.Pyramid(self.kind, F.BT_FORMAT_R16, T, B, lod_count)
        ds = [O.make_dataset(self.sources[s], side=s, lod_range=(0, lod_count)) for s in range(self.sides)]
        self.pyramid.add_datasets(ds)
        self.pyramid.run()
        self.existing = {tuple(int(v) for v in t) for t in self.pyramid.tiles()}
        self.clipmap = ClipmapRef(self.model, self.cfg, lod_count, self.atlas_size, self.existing)
        self.atlas = np.zeros((self.atlas_size, T, T), dtype=np.uint16)
        self.new_loads = []

    # -- one frame of the main-world systems -----------------------------------------------------------
    def move_view(self, view_pos):
        self.view_pos = tuple(float(v) for v in view_pos)
        loads = self.clipmap.update(self.view_pos)
        for tc, idx in loads:
            self.atlas[idx] = self.pyramid.tile(*tc)
            self.clipmap.mark_loaded(tc)
        self.new_loads = loads
        self.entries = np.ascontiguousarray(self.clipmap.entries())
        self.approximate_height = self.clipmap.approximate_height
        self.approx = O.model_approximation(self.model, self.view_pos, self.cfg.origin_lod, self.approximate_height)
        self.cull = O.culling(self.view_pos)
        return loads

    def terrain_config_uniform(self) -> F.TerrainConfigUniform:
        return F.TerrainConfigUniform(self.lod_count, self.model.min_height, self.model.max_height, float(np.float32(O.model_scale(self.model))))

    def attachment_uniforms(self):
        arr = (F.AttachmentUniform * 8)()
        c = np.float32(self.T - 2 * self.B)
        arr[0].size = float(c)
        arr[0].scale = float(c / np.float32(self.T))
        arr[0].offset = float(np.float32(self.B) / np.float32(self.T))
        return arr

    # -- oracle calls -------------------------------------------------------------------------------------
    def oracle_select(self, capacity=None):
        return O.select_tiles(self.vc, self.approx, self.cull, self.mesh, self.spherical, capacity)

    def oracle_vertex_inputs(self, tiles: np.ndarray, flags: int) -> O.VertexInputs:
        vi = O.VertexInputs()
        vi.view_config, vi.approximation = self.vc, self.approx
        vi.terrain_config = self.terrain_config_uniform()
        vi.attachments = self.attachment_uniforms()
        vi.mesh = self.mesh
        vi.view_world_position[:] = [float(np.float32(v)) for v in self.view_pos]
        vi.spherical, vi.texture_size = int(self.spherical), self.T
        vi.atlas_size = len(self.atlas)
        self._tiles_keep = np.ascontiguousarray(tiles, dtype=np.uint32)
        vi.tile_tree = self.entries.ctypes.data_as(C.POINTER(F.TileTreeEntry))
        vi.height_atlas = self.atlas.ctypes.data
        vi.tiles = self._tiles_keep.ctypes.data_as(C.POINTER(F.Tile))
        vi.tile_count, vi.flags = len(tiles), flags
        return vi

    def oracle_vertices(self, tiles, flags, first=0, count=None):
        vi = self.oracle_vertex_inputs(tiles, flags)
        n = len(tiles) * self.vc.vertices_per_tile
        count = n - first if count is None else count
        return O.gen_vertices(vi, first, count)

    # -- the same frame in C-ABI form ---------------------------------------------------------------------
    def view_frame(self) -> F.ViewFrame:
        f = F.ViewFrame()
        f.view_config, f.approximation, f.culling = self.vc, self.approx, self.cull
        f.view_world_position[:] = [float(np.float32(v)) for v in self.view_pos]
        f.tile_tree_entry_count = len(self.entries)
        f.tile_tree = self.entries.ctypes.data_as(C.POINTER(F.TileTreeEntry))
        self._origins = np.ascontiguousarray(self.clipmap.origins.reshape(-1))
        f.origins = self._origins.ctypes.data_as(C.POINTER(F.u32))
        return f

    def terrain_config(self):
        """TerrainConfig for the product API with the same parameters."""
        from bevy_terrain_b200 import api
        if self.model.kind == F.BT_KIND_ELLIPSOIDAL:
            model = api.TerrainModel.ellipsoid((0.0, 0.0, 0.0), 6378137.0, 6356752.314245, -12000.0, 9000.0)
        elif self.spherical:
            model = api.TerrainModel.sphere((0.0, 0.0, 0.0), 6371000.0, -12000.0, 9000.0)
        else:
            model = api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
        cfg = api.TerrainConfig(lod_count=self.lod_count, model=model, atlas_size=self.atlas_size)
        cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=self.T, border_size=self.B, mip_level_count=1,
                                                format=api.AttachmentFormat.R16))
        return cfg


def planar_view_positions(n=6, seed=7):
    """camera positions inside the terrain box of examples/planar.rs, altitude in [1, 400] above mid height"""
    rng = np.random.default_rng(seed)
    out = [(100.0, -100.0 + 250.0 + 20.0, -50.0), (0.0, -100.0 + 250.0 + 1.0, 0.0)]
    for _ in range(n - 2):
        out.append((float(rng.uniform(-950, 950)), float(-100.0 + 250.0 + rng.uniform(5, 400)), float(rng.uniform(-950, 950))))
    return out


def spherical_view_positions():
    R = 6371000.0
    h = -1500.0
    out = [(-(R + h + 10000.0), 0.0, 0.0), (-(R + h + 100000.0), 0.0, 0.0), (-3.0 * R, 0.0, 0.0)]
    d = np.array([1.0, 1.0, 1.0]) / np.sqrt(3.0)          # cube corner
    out.append(tuple((R + h + 50000.0) * d))
    e = np.array([1.0, 0.0, 1.0]) / np.sqrt(2.0)          # face edge
    out.append(tuple((R + h + 20000.0) * e))
    out.append((0.0, R + h + 5000.0, 1000.0))             # top face
    return out

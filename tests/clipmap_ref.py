"""Python restatement of the clipmap residency logic, used to build tile-tree tables for the parity
tests and to check the product's C++ bt_tile_tree (host_math.cpp) against.  TEST INFRASTRUCTURE.

Follows TileTree::update (src/terrain_data/tile_tree.rs:268-333), TileAtlas::update
(tile_atlas.rs:574-601) and TileAtlasState::{request_tile,release_tile,get_best_tile}
(tile_atlas.rs:418-503); the f64 arithmetic is done by the C oracle (bt_oracle_host.c).
"""
from __future__ import annotations

import ctypes as C
from collections import deque

import numpy as np

from bevy_terrain_b200 import _ffi as F
from tests import oracle as O

INVALID = 0xFFFFFFFF


class ClipmapRef:
    def __init__(self, model: F.TerrainModel, cfg: F.TerrainViewConfig, lod_count: int, atlas_size: int, existing):
        self.model, self.cfg, self.lod_count, self.atlas_size = model, cfg, lod_count, atlas_size
        self.side_count = 1 if model.kind == F.BT_KIND_PLANAR else 6
        self.ts = cfg.tree_size
        self.load_distance = cfg.load_distance * O.model_scale(model)
        self.approximate_height = float(np.float32((np.float32(model.min_height) + np.float32(model.max_height)) / np.float32(2.0)))
        n = self.side_count * lod_count * self.ts * self.ts
        self.coord = [None] * n
        self.requested = [False] * n
        self.origins = np.zeros((self.side_count, lod_count, 2), dtype=np.uint32)
        self.existing = set(existing)
        self.tile_states = {}                # tile -> [requests, loaded, atlas_index]
        self.unused = deque((None, i) for i in range(atlas_size))

    def _slot(self, side, lod, x, y):
        return ((side * self.lod_count + lod) * self.ts + x % self.ts) * self.ts + y % self.ts

    def _allocate(self):
        coord, idx = self.unused.popleft()   # "Atlas out of indices" -> IndexError
        self.tile_states.pop(coord, None)
        return idx

    def update(self, view_pos):
        L = O.lib()
        vp = (C.c_double * 3)(*view_pos)
        side = C.c_uint32()
        uv = (C.c_double * 2)()
        L.bto_coordinate_from_world_position(C.byref(self.model), vp, C.byref(side), uv)
        released, requested, loads = [], [], []
        for s in range(self.side_count):
            puv = (C.c_double * 2)()
            L.bto_project_to_side(C.byref(self.model), side.value, uv, s, puv)
            for lod in range(self.lod_count):
                org = (C.c_uint32 * 2)()
                L.bto_tile_tree_origin(self.ts, lod, puv, org)
                self.origins[s, lod] = (org[0], org[1])
                for x in range(self.ts):
                    for y in range(self.ts):
                        tc = (s, lod, org[0] + x, org[1] + y)
                        t = F.Tile(*tc)
                        dist = L.bto_tile_distance(C.byref(self.model), C.byref(t), s, puv, self.approximate_height, vp)
                        want = lod == 0 or dist < self.load_distance / float(1 << lod)
                        i = self._slot(*tc)
                        if self.coord[i] != tc:
                            if self.requested[i]:
                                self.requested[i] = False
                                released.append(self.coord[i])
                            self.coord[i] = tc
                        if not self.requested[i] and want:
                            self.requested[i] = True
                            requested.append(tc)
                        elif self.requested[i] and not want:
                            self.requested[i] = False
                            released.append(tc)
        for tc in released:
            if tc in self.existing:
                st = self.tile_states[tc]
                st[0] -= 1
                if st[0] == 0:
                    self.unused.append((tc, st[2]))
        for tc in requested:
            if tc not in self.existing:
                continue
            st = self.tile_states.get(tc)
            if st is not None:
                if st[0] == 0:
                    self.unused = deque(u for u in self.unused if u[1] != st[2])
                st[0] += 1
            else:
                idx = self._allocate()
                self.tile_states[tc] = [1, False, idx]
                loads.append((tc, idx))
        return loads

    def mark_loaded(self, tc):
        self.tile_states[tc][1] = True

    def entries(self) -> np.ndarray:
        out = np.full((len(self.coord), 2), INVALID, dtype=np.uint32)
        for i, tc in enumerate(self.coord):
            k = tc
            while k is not None and k[1] != INVALID:
                st = self.tile_states.get(k)
                if st is not None and st[1]:
                    out[i] = (st[2], k[1])
                    break
                k = (k[0], k[1] - 1, k[2] >> 1, k[3] >> 1) if k[1] > 0 else None
        return out

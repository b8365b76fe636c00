"""numpy restatement of the reference's CPU-side sampling twins (sample_attachment, TileTree::compute_blend / lookup_tile,
AtlasAttachment::sample, AttachmentData::sample — terrain_data/mod.rs:221-307, tile_tree.rs:223-266, tile_atlas.rs:249-258).
TEST INFRASTRUCTURE: an independent second formulation of "which height does the terrain have at this position"."""
import ctypes as C
import math

import numpy as np

from tests import oracle as O

f32 = np.float32


def twin_sample(scene, tree_entries, ts, pos, approximate_height, blend_distance_scaled, blend_range, lod_count):
    L = O.lib()
    m = scene.model
    # TerrainModel::surface_position (planar): local = ((x - px)/s, 0, (z - pz)/s); world = s*local + p + h * up
    lx, lz = (pos[0] - m.position[0]) / m.scale[0], (pos[2] - m.position[2]) / m.scale[2]
    surface = (m.scale[0] * lx + m.position[0], m.scale[1] * 0.0 + m.position[1] + float(approximate_height), m.scale[2] * lz + m.position[2])
    d = math.sqrt(sum((a - b) ** 2 for a, b in zip(scene.view_pos, surface)))
    target = f32(min(math.log2(blend_distance_scaled / d), lod_count - 0.00001))
    lod = int(target) if target > 0 else 0
    ratio = f32(0.0)
    if lod != 0:
        a, b = f32(lod) + f32(blend_range), f32(lod)
        ratio = f32(min(max((target - a) / (b - a), f32(0.0)), f32(1.0)))
    side = C.c_uint32(); uv = (C.c_double * 2)()
    L.bto_coordinate_from_world_position(C.byref(m), (C.c_double * 3)(*surface), C.byref(side), uv)

    def sample(tree_lod):
        tc = float(1 << tree_lod)
        xy = [min(uv[k] * tc, tc - 0.000001) for k in range(2)]
        idx = ((side.value * lod_count + tree_lod) * ts + int(xy[0]) % ts) * ts + int(xy[1]) % ts
        atlas_index, atlas_lod = tree_entries[idx]
        if atlas_lod == 0xFFFFFFFF:
            return np.zeros(4, f32)
        div = float(1 << (tree_lod - atlas_lod))
        auv = [f32(math.fmod(xy[k] / div, 1.0)) for k in range(2)]
        T, B = scene.T, scene.B
        scale, offset = f32(T - 2 * B) / f32(T), f32(B) / f32(T)
        u = [f32(f32(auv[k] * scale + offset) * f32(T)) - f32(0.5) for k in range(2)]
        r = [f32(math.fmod(float(u[k]), 1.0)) for k in range(2)]
        i = [int(u[k]) for k in range(2)]
        t = scene.atlas[atlas_index]
        val = lambda x, y: f32(t[i[1] + y, i[0] + x]) / f32(65535.0)
        l0 = val(0, 0) + (val(0, 1) - val(0, 0)) * r[1]
        l1 = val(1, 0) + (val(1, 1) - val(1, 0)) * r[1]
        return np.array([l0 + (l1 - l0) * r[0], 0, 0, 0], f32)

    value = sample(lod)
    if ratio > 0:
        v2 = sample(lod - 1)
        value = (value + (v2 - value) * ratio).astype(f32)
    return value, lod, float(ratio)



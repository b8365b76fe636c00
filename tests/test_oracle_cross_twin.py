"""Cross-check of the two independent formulations the reference itself contains: the GPU vertex path
(vertex.wgsl + functions.wgsl, f32, restated in oracle/bt_oracle.c) against the CPU sampling twins (sample_attachment,
tile_tree.rs / terrain_data/mod.rs, f64 + f32, restated in tests/twin.py).  Both must report the same terrain height at the
same world position, up to the f32-vs-f64 coordinate rounding."""
import numpy as np

from bevy_terrain_b200 import _ffi as F
from tests.scene import Scene
from tests.twin import twin_sample


def test_vertex_heights_agree_with_cpu_sampling_twin():
    scene = Scene("planar", lod_count=4, T=64, B=2, src=256, grid_size=16)
    pos = (100.0, 170.0, -50.0)
    scene.move_view(pos)
    tiles, ind, res = scene.oracle_select()
    flags = F.BT_FLAG_BLEND          # blend as the CPU twin does; no morph so that uv is the plain grid position
    rng = np.random.default_rng(0)
    pick = rng.choice(len(tiles), 40, replace=False)
    vpt = scene.vc.vertices_per_tile
    checked, worst = 0, 0.0
    vi = scene.oracle_vertex_inputs(tiles, flags)
    from tests import oracle as O
    for t in pick:
        verts = O.gen_vertices(vi, int(t) * vpt, vpt)
        for v in verts[::37]:
            x, y, z = (float(c) for c in v["position"][:3])
            if abs(x) > 995 or abs(z) > 995:
                continue        # the terrain edge clamps differently in the two formulations (quirk Q5)
            value, lod, ratio = twin_sample(scene, scene.entries, scene.cfg.tree_size, (x, pos[1], z), scene.approximate_height,
                                            scene.cfg.blend_distance * 1000.0, scene.cfg.blend_range, scene.lod_count)
            h_twin = 0.0 + (500.0 - 0.0) * float(value[0])
            h_vertex = y - (-100.0)
            # the twin measures the distance from the view to the point at approximate height, the vertex path too: the blend
            # LOD/ratio agree up to rounding, so heights agree to f32 rounding of the 0..500 m range (measured worst: 3e-5 m)
            worst = max(worst, abs(h_twin - h_vertex))
            assert abs(h_twin - h_vertex) < 1e-3, (t, x, z, h_twin, h_vertex, lod, ratio)
            checked += 1
    assert checked > 300
    assert worst < 1e-3

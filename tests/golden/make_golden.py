"""Generates tests/golden/*.npz with the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference ships NO golden vectors, tests or fixtures and cannot be executed here, so these
fixtures are produced by OUR oracle: they pin regressions of the oracle and give the CUDA path fixed
inputs/outputs that travel to the GPU box — they do not pin the oracle to the reference ("parity unpinned").
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bevy_terrain_b200 import _ffi as F  # noqa: E402
from tests import oracle as O  # noqa: E402
from tests.scene import Scene, planar_view_positions, spherical_view_positions  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ALL = F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS


def frame_arrays(scene):
    return dict(view_config=np.frombuffer(bytes(scene.vc), np.uint8), approximation=np.frombuffer(bytes(scene.approx), np.uint8),
                culling=np.frombuffer(bytes(scene.cull), np.uint8), mesh=np.frombuffer(bytes(scene.mesh), np.uint8),
                view_pos=np.array(scene.view_pos, np.float64), entries=scene.entries.copy())


def main():
    for name, kind, positions, lod_count, T in (("planar", "planar", planar_view_positions(3), 4, 16),
                                                 ("spherical", "spherical", spherical_view_positions()[:3], 3, 16)):
        scene = Scene(kind, lod_count=lod_count, T=T, B=2, src=48, grid_size=16)
        out = dict(lod_count=lod_count, T=T, B=2, n_views=len(positions), sources=np.stack(scene.sources))
        for i, pos in enumerate(positions):
            scene.move_view(pos)
            tiles, ind, res = scene.oracle_select()
            flags = ALL | (F.BT_FLAG_HIGH_PRECISION if kind == "spherical" else 0)
            n = len(tiles) * scene.vc.vertices_per_tile
            idx = np.unique(np.concatenate([np.arange(min(n, 1200)), np.arange(0, n, 2003), np.arange(max(n - 600, 0), n)]))
            vi = scene.oracle_vertex_inputs(tiles, flags)
            verts = np.concatenate([O.gen_vertices(vi, int(j), 1) for j in idx])
            for k, v in frame_arrays(scene).items():
                out[f"v{i}_{k}"] = v
            out[f"v{i}_atlas"] = scene.atlas.copy()
            out[f"v{i}_tiles"] = tiles
            out[f"v{i}_indirect"] = np.array([ind.vertex_count, ind.instance_count, ind.base_vertex, ind.base_instance], np.uint32)
            out[f"v{i}_flags"] = np.uint32(flags)
            out[f"v{i}_vertex_index"] = idx.astype(np.uint64)
            out[f"v{i}_vertices"] = verts
        np.savez_compressed(os.path.join(OUT, f"frame_{name}.npz"), **out)
    # preprocess goldens: planar R16 + Rgba8 with holes, and a spherical cross-face pyramid
    pre = {}
    for fmt, tag in ((F.BT_FORMAT_R16, "r16"), (F.BT_FORMAT_RGBA8, "rgba8")):
        src = O.synth_source(fmt, 70, 61, seed=0xBEEF, holes=True)
        p = O.Pyramid(F.BT_KIND_PLANAR, fmt, 16, 2, 3)
        p.add_datasets([O.make_dataset(src, lod_range=(0, 3))])
        p.run()
        pre[f"{tag}_source"] = src
        pre[f"{tag}_tiles"] = p.tiles()
        pre[f"{tag}_data"] = np.stack([p.tile(*t) for t in p.tiles()])
    faces = [O.synth_source(F.BT_FORMAT_R16, 40, 40, seed=0xF00 + s) for s in range(6)]
    p = O.Pyramid(F.BT_KIND_SPHERICAL, F.BT_FORMAT_R16, 16, 2, 2)
    p.add_datasets([O.make_dataset(faces[s], side=s, lod_range=(0, 2)) for s in range(6)])
    p.run()
    pre["sph_sources"] = np.stack(faces)
    pre["sph_tiles"] = p.tiles()
    pre["sph_data"] = np.stack([p.tile(*t) for t in p.tiles()])
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **pre)
    print("wrote", [f for f in os.listdir(OUT) if f.endswith(".npz")])


if __name__ == "__main__":
    main()

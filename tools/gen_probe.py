"""Times gen_vertices on the bench workload (configs[1]: planar 4096^2, 8 LODs, grid 64, 8 seeded views) for the four common flag
sets and checks sampled whole tiles bit for bit against the oracle.  For kernel A/B work on the GPU box:

    python tools/gen_probe.py [--rounds 6] [--check-tiles 6]
    BT_LIB_PATH=build/variants/lib_x.so python tools/gen_probe.py

Prints one JSON line: {"lib", "ms": {flagset: ms per launch}, "Gverts": {...}, "bit_exact": true}.  TEST/PROFILING TOOL (it loads the oracle)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--check-tiles", type=int, default=6)
    ap.add_argument("--views", type=int, default=8)
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from bevy_terrain_b200 import _ffi as F, api
    from tests.fullsize_util import assert_tiles_bit_exact, oracle_inputs

    class A:
        source, lods = 4096, 8
    ctx = api.Context(0)
    cfg, pre, src, _ = bench.build_terrain(api, ctx, A)
    vcfg = api.TerrainViewConfig(grid_size=64)
    views = []
    for pos in bench.view_positions(0)[: args.views]:
        atlas = api.TileAtlas(ctx, cfg)
        tree = api.TileTree(atlas, vcfg, max_vertices=6400 * bench.VPT)
        tree.add_existing_tiles(pre.tiles())
        api.load_resident_tiles(tree, pre, pos)
        tree.prepare()
        tree.select_tiles()
        tiles, ind = tree.readback()
        views.append({"tree": tree, "atlas": atlas, "pos": pos, "verts": int(ind.vertex_count), "tiles": tiles})
    sets = {"NORMALS": F.BT_FLAG_NORMALS, "MORPH_BLEND_NORMALS": F.BT_FLAG_MORPH | F.BT_FLAG_BLEND | F.BT_FLAG_NORMALS, "NONE": 0,
            "MORPH_BLEND": F.BT_FLAG_MORPH | F.BT_FLAG_BLEND}
    ms, gv, ok = {}, {}, True
    for name, flags in sets.items():
        r = bench.time_views(torch, views, flags, args.rounds, select=False)
        ms[name] = round(r["gen_vertices_ms"], 4)
        gv[name] = round(r["verts_per_step"] / r["gen_vertices_ms"] / 1e6, 2)
        if args.check_tiles:
            v = views[0]
            v["tree"].gen_vertices(flags)
            vi = oracle_inputs(ctx, cfg, v["atlas"], v["tree"], v["tiles"], flags, v["pos"], False, 512, 2, 0.0, 500.0)
            rng = np.random.default_rng(7)
            try:
                assert_tiles_bit_exact(v["tree"], vi, rng.choice(len(v["tiles"]), args.check_tiles, replace=False))
            except AssertionError as e:
                ok = False
                print(f"MISMATCH {name}: {e}", file=sys.stderr)
    print(json.dumps({"lib": os.environ.get("BT_LIB_PATH", "default"), "ms": ms, "Gverts_per_s": gv, "bit_exact": ok}), flush=True)


if __name__ == "__main__":
    main()

"""Times the single-GPU preprocess of one attachment at a time (CUDA events, un-profiled) and checks that two builds of
the library produce identical pyramids (checksums of every tile).  Usage on the GPU box:

    python tools/preprocess_probe.py [--source 65536] [--lods 8] [--reps 3]
    BT_LIB_PATH=build/variants/lib_x.so python tools/preprocess_probe.py ...

Prints one JSON line per attachment: {"format", "ms": [..], "tiles", "checksum"}.
"""
import argparse
import json
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", type=int, default=65536)
    ap.add_argument("--lods", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--checksum-tiles", type=int, default=64, help="tiles (spread over the list) to checksum")
    args = ap.parse_args()
    import numpy as np
    import torch
    from bevy_terrain_b200 import _ffi as F, api
    ctx = api.Context(0)
    S, L = args.source, args.lods
    model = api.TerrainModel.planar((0.0, -100.0, 0.0), 2000.0, 0.0, 500.0)
    for name, fmt, afmt, seed in (("R16", F.BT_FORMAT_R16, api.AttachmentFormat.R16, 0x5EED0001),
                                  ("Rgba8", F.BT_FORMAT_RGBA8, api.AttachmentFormat.Rgba8, 0x5EED0002)):
        cfg = api.TerrainConfig(lod_count=L, model=model)
        cfg.add_attachment(api.AttachmentConfig(name=name, texture_size=512, border_size=2, format=afmt))
        src = ctx.synth_source(fmt, S, S, x0=0, y0=0, seed=seed)
        torch.cuda.synchronize()
        times, checksum, n = [], 0, 0
        for rep in range(args.reps):
            pre = api.Preprocessor(ctx, cfg)
            pre.preprocess_tile(api.PreprocessDataset(attachment_index=0, device_pixels=src, width=S, height=S, lod_range=range(0, L)))
            pre.reserve()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pre.run()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            if rep == args.reps - 1:
                tiles = pre.tiles(0)
                n = len(tiles)
                step = max(1, n // args.checksum_tiles)
                for t in tiles[::step]:
                    checksum = zlib.crc32(np.ascontiguousarray(pre.get_tile(0, t)).tobytes(), checksum)
            pre.close()
        ctx.free(src)
        print(json.dumps({"lib": os.environ.get("BT_LIB_PATH", "default"), "format": name, "source": S, "ms": [round(t, 3) for t in times],
                          "tiles": n, "checksum": checksum}), flush=True)


if __name__ == "__main__":
    main()

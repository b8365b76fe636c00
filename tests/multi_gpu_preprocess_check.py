"""Run under torchrun on N GPUs: the sharded CUDA preprocess (NCCL exchange) must reproduce the single-GPU
pyramid AND the oracle byte for byte.  Launched by tests/test_gpu_multi_rank.py and usable stand-alone:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_preprocess_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bevy_terrain_b200 import _ffi as F  # noqa: E402
from bevy_terrain_b200 import api  # noqa: E402
from tests import oracle as O  # noqa: E402


def sharded_run(pre, world):
    """the host-side phase sequence of one job: kernels / NCCL all_gather / kernels"""
    pre.run_split_downsample()
    for nbytes, pack, unpack in ((pre.roots_bytes(), pre.roots_pack, pre.roots_unpack_and_finish),
                                 (pre.halo_bytes(), pre.halo_pack, pre.halo_unpack)):
        if nbytes:
            send = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            recv = torch.empty(nbytes * world, dtype=torch.uint8, device="cuda")
            pack(send.data_ptr())
            dist.all_gather_into_tensor(recv, send)
            unpack(recv.data_ptr())
        else:
            unpack(0)
    pre.run_stitch()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    comm = api.Comm.from_torch_distributed(ctx)          # the library's own NCCL communicator (bt_comm_*)
    info = comm.info()
    assert info["rank"] == rank and info["world"] == world
    modes = {"library": lambda pre: pre.run_sharded(comm), "host_phases": lambda pre: sharded_run(pre, world)}
    T, B, L, S = 64, 2, 5, 1024
    for fmt, afmt in ((F.BT_FORMAT_R16, api.AttachmentFormat.R16), (F.BT_FORMAT_RGBA8, api.AttachmentFormat.Rgba8)):
        src = O.synth_source(fmt, S, S, seed=123, holes=True)
        truth = O.Pyramid(F.BT_KIND_PLANAR, fmt, T, B, L)
        truth.add_datasets([O.make_dataset(src, lod_range=(0, L))])
        truth.run()
        cfg = api.TerrainConfig(lod_count=L, model=api.TerrainModel.planar((0, 0, 0), 1.0, 0.0, 1.0))
        cfg.add_attachment(api.AttachmentConfig(name="a", texture_size=T, border_size=B, format=afmt))
        # (a) whole source on every rank, (b) only this rank's band of source rows on the device (row window)
        for windowed, mode in ((False, "library"), (True, "library"), (False, "host_phases")):
            pre = api.Preprocessor(ctx, cfg, rank=rank, world=world)
            if windowed:
                n_rows = 1 << (L - 1)
                y0 = (rank * n_rows // world) * (S // n_rows)
                y1 = ((rank + 1) * n_rows // world) * (S // n_rows)
                r0, r1 = max(y0 - 1, 0), min(y1 + 1, S)
                band = np.ascontiguousarray(src[r0:r1])
                dptr = ctx.alloc(band.nbytes)
                ctx.h2d(dptr, band)
                pre.preprocess_tile(api.PreprocessDataset(device_pixels=dptr, width=S, height=S, lod_range=range(0, L), row_begin=r0, row_count=r1 - r0))
            else:
                pre.preprocess_tile(api.PreprocessDataset(pixels=src, lod_range=range(0, L)))
            if mode == "library":
                pre.reserve()
            modes[mode](pre)
            torch.cuda.synchronize()
            mine = [tuple(int(v) for v in t) for t in pre.tiles()]
            for tc in mine:
                got, exp = pre.get_tile(0, tc), truth.tile(*tc)
                assert np.array_equal(got, exp), f"rank {rank} fmt {fmt} windowed {windowed} {mode}: tile {tc} differs ({int((got != exp).sum())} texels)"
            # the device-side content hash: sum over ranks == hash of the single-GPU pyramid of the same source
            total = comm.all_reduce_sum_u64([pre.content_hash(0)])[0]
            single = api.Preprocessor(ctx, cfg)
            single.preprocess_tile(api.PreprocessDataset(pixels=src, lod_range=range(0, L))).run()
            assert total == single.content_hash(0), f"rank {rank}: sharded hash {total:#x} != single-GPU hash {single.content_hash(0):#x}"
            single.close()
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            union = [t for lst in gathered for t in lst]
            assert len(union) == len(set(union)) == (4 ** L - 1) // 3, (len(union), len(set(union)))
            pre.close()
            if windowed:
                ctx.free(dptr)
    # ---- cube: six faces of one attachment per job; every rank owns the same row band of each face, ghosts across face edges
    Ts, Ls, Ss = 32, 4, 160
    for fmt, afmt in ((F.BT_FORMAT_R16, api.AttachmentFormat.R16), (F.BT_FORMAT_RGBA8, api.AttachmentFormat.Rgba8)):
        if (1 << (Ls - 1)) < world:
            continue
        faces = [O.synth_source(fmt, Ss, Ss, seed=900 + 31 * side, holes=True) for side in range(6)]
        truth = O.Pyramid(F.BT_KIND_SPHERICAL, fmt, Ts, B, Ls)
        truth.add_datasets([O.make_dataset(faces[side], side=side, lod_range=(0, Ls)) for side in range(6)])
        truth.run()
        cfg = api.TerrainConfig(lod_count=Ls, model=api.TerrainModel.sphere((0, 0, 0), 1.0, 0.0, 1.0))
        cfg.add_attachment(api.AttachmentConfig(name="a", texture_size=Ts, border_size=B, format=afmt))
        pre = api.Preprocessor(ctx, cfg, rank=rank, world=world)
        pre.preprocess_spherical(api.SphericalDataset(faces=faces, lod_range=range(0, Ls)))
        pre.run_sharded(comm)
        torch.cuda.synchronize()
        mine = [tuple(int(v) for v in t) for t in pre.tiles()]
        for tc in mine:
            got, exp = pre.get_tile(0, tc), truth.tile(*tc)
            assert np.array_equal(got, exp), f"rank {rank} fmt {fmt} cube: tile {tc} differs ({int((got != exp).sum())} texels)"
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        union = [t for lst in gathered for t in lst]
        assert len(union) == len(set(union)) == 6 * (4 ** Ls - 1) // 3, (len(union), len(set(union)))
        pre.close()
    # ---- two attachments in ONE wave (config-5 shape: height R16 + albedo Rgba8 -> one grouped collective)
    srcs = [O.synth_source(F.BT_FORMAT_R16, S, S, seed=321), O.synth_source(F.BT_FORMAT_RGBA8, S, S, seed=654)]
    cfg = api.TerrainConfig(lod_count=L, model=api.TerrainModel.planar((0, 0, 0), 1.0, 0.0, 1.0))
    cfg.add_attachment(api.AttachmentConfig(name="height", texture_size=T, border_size=B, format=api.AttachmentFormat.R16))
    cfg.add_attachment(api.AttachmentConfig(name="albedo", texture_size=T, border_size=B, format=api.AttachmentFormat.Rgba8))
    pre, single = api.Preprocessor(ctx, cfg, rank=rank, world=world), api.Preprocessor(ctx, cfg)
    for a in (0, 1):
        pre.preprocess_tile(api.PreprocessDataset(attachment_index=a, pixels=srcs[a], lod_range=range(0, L)))
        single.preprocess_tile(api.PreprocessDataset(attachment_index=a, pixels=srcs[a], lod_range=range(0, L)))
    pre.reserve().run_sharded(comm)
    single.run()
    torch.cuda.synchronize()
    for a in (0, 1):
        for tc in pre.tiles(a):
            assert np.array_equal(pre.get_tile(a, tc), single.get_tile(a, tc)), f"rank {rank}: two-attachment wave, attachment {a} tile {tc}"
    sums = comm.all_reduce_sum_u64([pre.content_hash(0), pre.content_hash(1)])
    assert sums == [single.content_hash(0), single.content_hash(1)]
    pre.close(); single.close()
    dist.barrier()
    if rank == 0:
        print(f"MULTI_GPU_PREPROCESS_OK world={world} (planar + cube + two-attachment wave; library run_sharded and host phases; "
              f"nccl {info['nccl_version']}; device hashes equal the single-GPU pyramid)")
    comm.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

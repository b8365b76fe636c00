"""World-size-2 and -4 `gloo` tests (CPU) of the N>1 preprocess path.

What runs where:
  * the PRODUCT's host-side sharding logic — row ownership, ghost/replica tiles, the byte layout of the
    sub-root gather and of the halo exchange — comes from the C-ABI library through a plan-only
    bt_preprocessor (ctx = NULL, no GPU): bt_preprocessor_plan / _plan_tiles / _plan_exchange / _halo_bytes /
    _roots_bytes.  The very same plan drives the CUDA pack/unpack kernels on the GPU.
  * the collectives are real torch.distributed all_gathers (gloo) between two processes.
  * the pixel work (what the CUDA kernels do on a GPU) is done by the oracle, tile by tile.
The result of every rank must equal the single-rank oracle pyramid byte for byte, and the ranks' owned
tiles must partition the pyramid.
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bevy_terrain_b200 import _ffi as F

T, B, L = 16, 2, 4
SRC = 96


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _plan(lib, rank, world, fmt, kind=F.BT_KIND_PLANAR):
    att = (F.AttachmentConfig * 1)(F.AttachmentConfig(T, B, 1, fmt))
    h = C.c_void_p()
    assert lib.bt_preprocessor_create(None, kind, att, 1, L, F.Shard(rank, world), C.byref(h)) == F.BT_OK
    sides = 6 if kind == F.BT_KIND_SPHERICAL else 1
    ds = (F.Dataset * sides)()
    for side, d in enumerate(ds):
        d.attachment_index, d.side = 0, side
        d.top_left[:] = [0.0, 0.0]
        d.bottom_right[:] = [1.0, 1.0]
        d.lod_begin, d.lod_end, d.width, d.height = 0, L, SRC, SRC
    assert lib.bt_preprocessor_add_datasets(h, ds, sides) == F.BT_OK
    assert lib.bt_preprocessor_plan(h) == F.BT_OK
    return h


def _exchange_plan(lib, h, kind, unpack):
    n = F.u32()
    assert lib.bt_preprocessor_plan_exchange(h, kind, unpack, None, 0, C.byref(n)) == F.BT_OK
    buf = (F.ExchangeCopy * max(n.value, 1))()
    assert lib.bt_preprocessor_plan_exchange(h, kind, unpack, buf, n.value, C.byref(n)) == F.BT_OK
    return buf[: n.value]


def _worker(rank, world, port, fmt, out_dir, kind=F.BT_KIND_PLANAR):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import oracle as O
        lib = F.load_library()
        OL = O.lib()
        px = 2 if fmt == F.BT_FORMAT_R16 else 4
        dt, shape = (np.uint16, (T, T)) if fmt == F.BT_FORMAT_R16 else (np.uint8, (T, T, 4))
        spherical = kind == F.BT_KIND_SPHERICAL
        sides = 6 if spherical else 1
        srcs = [O.synth_source(fmt, SRC, SRC, seed=77 + 13 * side, holes=True) for side in range(sides)]
        # truth: single-rank oracle pyramid; `pre`: the same after split+downsample only
        truth = O.Pyramid(kind, fmt, T, B, L)
        truth.add_datasets([O.make_dataset(srcs[side], side=side, lod_range=(0, L)) for side in range(sides)])
        truth.run()
        pre = O.Pyramid(kind, fmt, T, B, L)
        pre.add_datasets([O.make_dataset(srcs[side], side=side, lod_range=(0, L)) for side in range(sides)])
        pre.run(phases=1)

        h = _plan(lib, rank, world, fmt, kind)
        n = F.u32()
        lib.bt_preprocessor_plan_tiles(h, 0, None, None, 0, C.byref(n))
        tiles = (F.Tile * n.value)()
        kinds = (F.u32 * n.value)()
        lib.bt_preprocessor_plan_tiles(h, 0, tiles, kinds, n.value, C.byref(n))
        root_lod = int(np.log2(world))
        store = {}
        for t, k in zip(tiles, kinds):
            key = (t.side, t.lod, t.x, t.y)
            if k == 0 and t.lod >= root_lod:          # owned rows: what this rank's split/downsample kernels produce
                store[key] = pre.tile(*key).copy()
            else:                                      # ghosts, replicas and the redundant top levels start empty
                store[key] = np.zeros(shape, dtype=dt)
        kind_of = {(t.side, t.lod, t.x, t.y): int(k) for t, k in zip(tiles, kinds)}

        def flat(key):
            return store[key].reshape(-1).view(np.uint8)

        def exchange(kind, nbytes):
            if nbytes == 0:
                return
            send = np.zeros(nbytes, dtype=np.uint8)
            for c in _exchange_plan(lib, h, kind, 0):
                key = (c.tile.side, c.tile.lod, c.tile.x, c.tile.y)
                o = (c.tile_row * T + c.tile_col) * px
                send[c.buffer_offset:c.buffer_offset + c.pixels * px] = flat(key)[o:o + c.pixels * px]
            gathered = [torch.zeros(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(gathered, torch.from_numpy(send))
            recv = torch.cat(gathered).numpy()
            for c in _exchange_plan(lib, h, kind, 1):
                key = (c.tile.side, c.tile.lod, c.tile.x, c.tile.y)
                o = (c.tile_row * T + c.tile_col) * px
                flat(key)[o:o + c.pixels * px] = recv[c.buffer_offset:c.buffer_offset + c.pixels * px]

        rb, hb = F.u64(), F.u64()
        lib.bt_preprocessor_roots_bytes(h, C.byref(rb))
        lib.bt_preprocessor_halo_bytes(h, C.byref(hb))
        assert rb.value == sides * (1 << root_lod) * T * T * px
        exchange(1, rb.value)
        # the levels above the shard boundary, computed redundantly on every rank from the gathered root level
        for lod in range(root_lod - 1, -1, -1):
            for side in range(sides):
                for x in range(1 << lod):
                    for y in range(1 << lod):
                        child = (C.c_void_p * 4)(*[store[(side, lod + 1, 2 * x + c % 2, 2 * y + c // 2)].ctypes.data for c in range(4)])
                        OL.bto_downsample_tile_raw(fmt, T, B, store[(side, lod, x, y)].ctypes.data, child)
        exchange(0, hb.value)
        # stitch what this rank stitches on the GPU: its owned rows (and every tile of the redundant top levels)
        for key, k in kind_of.items():
            if k == 1 or (k == 2 and key[1] >= root_lod):
                continue
            nb = (F.Tile * 8)()
            t = F.Tile(*key)
            OL.bto_tile_neighbours(C.byref(t), int(spherical), nb)
            ptrs = (C.c_void_p * 8)()
            for i in range(8):
                nk = (nb[i].side, nb[i].lod, nb[i].x, nb[i].y)
                if nb[i].side != 0xFFFFFFFF:
                    assert nk in store, f"rank {rank}: neighbour {nk} of {key} is neither owned, ghost nor replica"
                    ptrs[i] = store[nk].ctypes.data
            OL.bto_stitch_tile_raw(fmt, T, B, C.byref(t), store[key].ctypes.data, nb, ptrs)
        # compare with the single-rank pyramid
        owned = sorted(k for k, v in kind_of.items() if v == 0)
        for key in owned:
            assert np.array_equal(store[key], truth.tile(*key)), f"rank {rank}: tile {key} differs from the single-rank result"
        for key, k in kind_of.items():
            if key[1] < root_lod:   # redundant top levels are identical everywhere
                assert np.array_equal(store[key], truth.tile(*key))
        all_owned = [None] * world
        dist.all_gather_object(all_owned, owned)
        union = [k for lst in all_owned for k in lst]
        assert len(union) == len(set(union)), "owned tiles overlap between ranks"
        assert set(union) == {tuple(int(v) for v in t) for t in truth.tiles()}
        lib.bt_preprocessor_destroy(h)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fmt,world,kind", [(F.BT_FORMAT_R16, 2, F.BT_KIND_PLANAR), (F.BT_FORMAT_RGBA8, 2, F.BT_KIND_PLANAR), (F.BT_FORMAT_R16, 4, F.BT_KIND_PLANAR),
                                            (F.BT_FORMAT_R16, 2, F.BT_KIND_SPHERICAL), (F.BT_FORMAT_RGBA8, 4, F.BT_KIND_SPHERICAL)])
def test_sharded_preprocess_exchange_gloo(fmt, world, kind, tmp_path):
    """world 2: the sub-root level is LOD 1; world 4: LOD 2, two redundant top levels, interior ranks with halos on both sides;
    spherical: six faces, every rank owns the same row band of each face, ghosts across face edges (rotated neighbours)"""
    mp.spawn(_worker, args=(world, _free_port(), fmt, str(tmp_path), kind), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_plan_partitions_rows_for_world_1_2_4_8():
    lib = F.load_library()
    for world in (1, 2, 4, 8):
        owned = []
        for rank in range(world):
            h = _plan(lib, rank, world, F.BT_FORMAT_R16) if (1 << (L - 1)) >= world else None
            if h is None:
                continue
            n = F.u32()
            lib.bt_preprocessor_plan_tiles(h, 0, None, None, 0, C.byref(n))
            tiles, kinds = (F.Tile * n.value)(), (F.u32 * n.value)()
            lib.bt_preprocessor_plan_tiles(h, 0, tiles, kinds, n.value, C.byref(n))
            owned += [(t.side, t.lod, t.x, t.y) for t, k in zip(tiles, kinds) if k == 0]
            hb = F.u64()
            lib.bt_preprocessor_halo_bytes(h, C.byref(hb))
            root = int(np.log2(world))
            exp = sum(2 * (1 << lod) * B * (T - 2 * B) * 2 for lod in range(root + 1, L)) if world > 1 else 0
            assert hb.value == exp
            lib.bt_preprocessor_destroy(h)
        assert len(owned) == len(set(owned)) == (4 ** L - 1) // 3, (world, len(owned))


def test_cube_shard_plan_partitions_all_six_faces():
    """plan-only preprocessor (no GPU): for world 1..8 the owned tiles of all ranks partition the six-face pyramid, every ghost
    lies above the root level and is a neighbour (possibly across a face edge) of an owned tile, and the unpack plan of every
    rank covers exactly its ghost tiles"""
    from tests import oracle as O
    lib, OL = F.load_library(), O.lib()
    for world in (1, 2, 4, 8):
        owned_all = []
        root = int(np.log2(world))
        for rank in range(world):
            h = _plan(lib, rank, world, F.BT_FORMAT_R16, F.BT_KIND_SPHERICAL)
            n = F.u32()
            lib.bt_preprocessor_plan_tiles(h, 0, None, None, 0, C.byref(n))
            tiles, kinds = (F.Tile * n.value)(), (F.u32 * n.value)()
            lib.bt_preprocessor_plan_tiles(h, 0, tiles, kinds, n.value, C.byref(n))
            keys = [(t.side, t.lod, t.x, t.y) for t in tiles]
            owned = {k for k, kd in zip(keys, kinds) if kd == 0}
            ghosts = {k for k, kd in zip(keys, kinds) if kd == 1}
            owned_all += sorted(owned)
            if world == 1:
                assert not ghosts
            neighbours_of_owned = set()
            for k in owned:
                nb = (F.Tile * 8)()
                t = F.Tile(*k)
                OL.bto_tile_neighbours(C.byref(t), 1, nb)
                neighbours_of_owned |= {(b.side, b.lod, b.x, b.y) for b in nb if b.side != 0xFFFFFFFF}
            for g in ghosts:
                assert g[1] > root and g in neighbours_of_owned and g not in owned
            # what stitch will read but this rank does not own must be a ghost (above the root level) or a replica
            missing = {k for k in neighbours_of_owned if k not in owned and k[1] > root} - ghosts
            assert not missing, (world, rank, sorted(missing)[:4])
            unpacked = {(c.tile.side, c.tile.lod, c.tile.x, c.tile.y) for c in _exchange_plan(lib, h, 0, 1)}
            assert unpacked == ghosts
            hb = F.u64()
            lib.bt_preprocessor_halo_bytes(h, C.byref(hb))
            packed = sum(c.pixels for c in _exchange_plan(lib, h, 0, 0)) * 2
            assert packed == hb.value
            lib.bt_preprocessor_destroy(h)
        assert len(owned_all) == len(set(owned_all)) == 6 * (4 ** L - 1) // 3, (world, len(owned_all))


def test_sharded_job_argument_errors():
    lib = F.load_library()
    att = (F.AttachmentConfig * 1)(F.AttachmentConfig(T, B, 1, F.BT_FORMAT_R16))
    h = C.c_void_p()
    assert lib.bt_preprocessor_create(None, F.BT_KIND_PLANAR, att, 1, L, F.Shard(0, 3), C.byref(h)) == F.BT_ERR_INVALID_ARG   # not a power of two
    assert lib.bt_preprocessor_create(None, F.BT_KIND_PLANAR, att, 1, L, F.Shard(2, 2), C.byref(h)) == F.BT_ERR_INVALID_ARG   # rank >= world

    def dataset(side, lod_end=L):
        d = F.Dataset()
        d.attachment_index, d.side = 0, side
        d.top_left[:] = [0.0, 0.0]
        d.bottom_right[:] = [1.0, 1.0]
        d.lod_begin, d.lod_end, d.width, d.height = 0, lod_end, SRC, SRC
        return d

    assert lib.bt_preprocessor_create(None, F.BT_KIND_SPHERICAL, att, 1, L, F.Shard(0, 2), C.byref(h)) == F.BT_OK
    five = (F.Dataset * 5)(*[dataset(s) for s in range(5)])
    assert lib.bt_preprocessor_add_datasets(h, five, 5) == F.BT_ERR_INVALID_ARG            # a sharded cube job needs all six faces
    mixed = (F.Dataset * 6)(*[dataset(s, L if s else L - 1) for s in range(6)])
    assert lib.bt_preprocessor_add_datasets(h, mixed, 6) == F.BT_ERR_INVALID_ARG           # ... with one lod_range
    twice = (F.Dataset * 6)(*[dataset(min(s, 4)) for s in range(6)])
    assert lib.bt_preprocessor_add_datasets(h, twice, 6) == F.BT_ERR_INVALID_ARG           # ... on distinct sides
    lib.bt_preprocessor_destroy(h)
    assert lib.bt_preprocessor_create(None, F.BT_KIND_PLANAR, att, 1, L, F.Shard(0, 2), C.byref(h)) == F.BT_OK
    two = (F.Dataset * 2)(dataset(0), dataset(0))
    assert lib.bt_preprocessor_add_datasets(h, two, 2) == F.BT_ERR_INVALID_ARG             # sharded planar: one dataset per job
    lib.bt_preprocessor_destroy(h)

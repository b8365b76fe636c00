"""Host-side mirror of the reference's plugin API for the hot path, on top of the C ABI.

Names, fields and argument meaning follow bevy_terrain's prelude (src/lib.rs:60-86) so that the
parity tests read like the reference's examples:

    TerrainModel.planar / .sphere      src/math/terrain_model.rs:84-104
    AttachmentFormat, AttachmentConfig src/terrain_data/mod.rs:38-109
    TerrainConfig(.add_attachment)     src/terrain.rs:27-56
    TerrainViewConfig                  src/terrain_view.rs:19-64
    TileAtlas, TileTree                src/terrain_data/tile_atlas.rs:531, tile_tree.rs:135
    Preprocessor, PreprocessDataset,
    SphericalDataset                   src/preprocess/preprocessor.rs:29-55,214-343

Everything that computes runs inside libbevy_terrain_b200.so (CUDA sm_100a); this module only
marshals structs.  There is no Python/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import enum
from dataclasses import dataclass, field
from typing import Iterable, Sequence

import numpy as np

from . import _ffi as F

VERTEX_DTYPE = np.dtype([("position", "<f4", 4), ("normal", "<f4", 3), ("uv", "<f4", 2), ("tile_index", "<u4")])


class Context:
    """bt_context: one CUDA device."""

    def __init__(self, device: int = 0):
        self.lib = F.load_library()
        self._h = C.c_void_p()
        st = self.lib.bt_context_create(device, C.byref(self._h))
        if st != F.BT_OK:
            raise F.BtError(st, "bt_context_create failed: no usable sm_100a CUDA device (there is no CPU fallback)")
        self.device = device

    def check(self, status: int) -> None:
        if status != F.BT_OK:
            raise F.BtError(status, (self.lib.bt_last_error(self._h) or b"").decode())

    def synchronize(self) -> None:
        self.check(self.lib.bt_context_synchronize(self._h))

    # raw device memory (tests/bench without a CUDA python binding)
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.check(self.lib.bt_device_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr: int) -> None:
        self.check(self.lib.bt_device_free(self._h, C.c_void_p(ptr)))

    def h2d(self, dst: int, src: np.ndarray, stream: int = 0) -> None:
        src = np.ascontiguousarray(src)
        self.check(self.lib.bt_memcpy_h2d(self._h, C.c_void_p(dst), src.ctypes.data, src.nbytes, C.c_void_p(stream)))

    def d2h(self, dst: np.ndarray, src: int, stream: int = 0) -> None:
        assert dst.flags["C_CONTIGUOUS"]
        self.check(self.lib.bt_memcpy_d2h(self._h, dst.ctypes.data, C.c_void_p(src), dst.nbytes, C.c_void_p(stream)))

    def synth_source(self, fmt: int, width: int, height: int, x0: int = 0, y0: int = 0, seed: int = 0x5EED0001,
                     holes: bool = False, out_ptr: int | None = None, stream: int = 0) -> int:
        px = 2 if fmt == F.BT_FORMAT_R16 else 4
        ptr = out_ptr if out_ptr is not None else self.alloc(width * height * px)
        self.check(self.lib.bt_synth_source(self._h, fmt, width, height, x0, y0, seed, int(holes), C.c_void_p(ptr), C.c_void_p(stream)))
        return ptr

    def close(self) -> None:
        if self._h:
            self.lib.bt_context_destroy(self._h)
            self._h = C.c_void_p()


class Comm:
    """bt_comm: the multi-GPU communicator of the sharded preprocessor (NCCL over NVLink, one process per GPU).

    Bootstrap like any NCCL program: rank 0 creates the 128-byte unique id (`Comm.unique_id()`), the host distributes it
    (`broadcast` is any callable bytes -> bytes that returns rank 0's value on every rank, e.g. a torch.distributed or MPI
    broadcast), then every rank constructs the communicator (collective)."""

    def __init__(self, ctx: Context, rank: int, world: int, unique_id: bytes):
        self.ctx, self.rank, self.world = ctx, rank, world
        self._h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        ctx.check(ctx.lib.bt_comm_create(ctx._h, rank, world, buf, 128, C.byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        lib = F.load_library()
        buf = C.create_string_buffer(128)
        st = lib.bt_comm_unique_id(buf, 128)
        if st != F.BT_OK:
            raise F.BtError(st, "bt_comm_unique_id: " + (lib.bt_comm_load_error() or b"").decode())
        return buf.raw

    @staticmethod
    def from_torch_distributed(ctx: Context) -> "Comm":
        """convenience bootstrap for hosts that already run torch.distributed: the unique id travels as a broadcast tensor"""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        raw = Comm.unique_id() if rank == 0 else bytes(128)
        t = torch.tensor(list(raw), dtype=torch.uint8, device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.broadcast(t, 0)
        return Comm(ctx, rank, world, bytes(t.cpu().tolist()))

    def info(self) -> dict:
        r, w, v, p = F.u32(), F.u32(), F.i32(), F.u32()
        self.ctx.check(self.ctx.lib.bt_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v), C.byref(p)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value, "peer_access": bool(p.value)}

    def barrier(self, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.bt_comm_barrier(self._h, C.c_void_p(stream)))

    def all_reduce_sum_u64(self, values, stream: int = 0) -> list:
        arr = (F.u64 * len(values))(*[int(v) & 0xFFFFFFFFFFFFFFFF for v in values])
        self.ctx.check(self.ctx.lib.bt_comm_all_reduce_sum_u64(self._h, arr, len(values), C.c_void_p(stream)))
        return list(arr)

    def close(self) -> None:
        if self._h:
            self.ctx.lib.bt_comm_destroy(self._h)
            self._h = C.c_void_p()


class AttachmentFormat(enum.IntEnum):  # terrain_data/mod.rs:38-57 (ids)
    Rgba8 = F.BT_FORMAT_RGBA8
    R16 = F.BT_FORMAT_R16


@dataclass
class AttachmentConfig:  # terrain_data/mod.rs:88-109
    name: str = ""
    texture_size: int = 512
    border_size: int = 1
    mip_level_count: int = 1
    format: AttachmentFormat = AttachmentFormat.R16

    def to_c(self) -> F.AttachmentConfig:
        return F.AttachmentConfig(self.texture_size, self.border_size, self.mip_level_count, int(self.format))

    @property
    def pixel_size(self) -> int:
        return 2 if self.format == AttachmentFormat.R16 else 4

    @property
    def dtype_shape(self):
        T = self.texture_size
        return (np.uint16, (T, T)) if self.format == AttachmentFormat.R16 else (np.uint8, (T, T, 4))


class TerrainModel:
    """TerrainModel (src/math/terrain_model.rs:40-203); planar and sphere constructors."""

    def __init__(self, c: F.TerrainModel):
        self.c = c

    @staticmethod
    def planar(position: Sequence[float], side_length: float, min_height: float, max_height: float) -> "TerrainModel":
        m = F.TerrainModel()
        F.load_library().bt_terrain_model_planar((F.f64 * 3)(*position), side_length, min_height, max_height, C.byref(m))
        return TerrainModel(m)

    @staticmethod
    def sphere(position: Sequence[float], radius: float, min_height: float, max_height: float) -> "TerrainModel":
        m = F.TerrainModel()
        F.load_library().bt_terrain_model_sphere((F.f64 * 3)(*position), radius, min_height, max_height, C.byref(m))
        return TerrainModel(m)

    @staticmethod
    def ellipsoid(position: Sequence[float], major_axis: float, minor_axis: float, min_height: float, max_height: float) -> "TerrainModel":
        """TerrainModel::ellipsoid (terrain_model.rs:106-128): polar (minor) axis along world y."""
        m = F.TerrainModel()
        F.load_library().bt_terrain_model_ellipsoid((F.f64 * 3)(*position), major_axis, minor_axis, min_height, max_height, C.byref(m))
        return TerrainModel(m)

    def is_spherical(self) -> bool:
        return self.c.kind != F.BT_KIND_PLANAR

    def scale(self) -> float:
        return F.load_library().bt_terrain_model_scale(C.byref(self.c))

    def mesh_transform(self) -> F.MeshTransform:
        t = F.MeshTransform()
        F.load_library().bt_terrain_model_mesh_transform(C.byref(self.c), C.byref(t))
        return t


@dataclass
class TerrainConfig:  # terrain.rs:27-56
    lod_count: int = 1
    model: TerrainModel | None = None
    atlas_size: int = 1024
    path: str = ""
    attachments: list = field(default_factory=list)

    def __post_init__(self):
        if self.model is None:
            self.model = TerrainModel.sphere((0.0, 0.0, 0.0), 1.0, 0.0, 1.0)  # terrain.rs:43

    def add_attachment(self, attachment: AttachmentConfig) -> "TerrainConfig":
        self.attachments.append(attachment)
        return self


def TerrainViewConfig(**overrides) -> F.TerrainViewConfig:
    """TerrainViewConfig::default() with field overrides (terrain_view.rs:47-64)."""
    c = F.TerrainViewConfig()
    F.load_library().bt_terrain_view_config_default(C.byref(c))
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


class TileAtlas:
    """TileAtlas + GpuTileAtlas (tile_atlas.rs:531-551, gpu_tile_atlas.rs:208-223): the device atlas."""

    def __init__(self, ctx: Context, config: TerrainConfig):
        self.ctx, self.config = ctx, config
        self.model = config.model
        self.lod_count = config.lod_count
        self.atlas_size = config.atlas_size
        self.attachments = list(config.attachments)
        d = F.TerrainDesc()
        d.kind = F.BT_KIND_SPHERICAL if self.model.is_spherical() else F.BT_KIND_PLANAR
        d.lod_count, d.atlas_size, d.attachment_count = config.lod_count, config.atlas_size, len(self.attachments)
        for i, a in enumerate(self.attachments):
            d.attachments[i] = a.to_c()
        d.min_height, d.max_height = self.model.c.min_height, self.model.c.max_height
        d.scale = self.model.scale()
        d.transform = self.model.mesh_transform()
        self._h = C.c_void_p()
        ctx.check(ctx.lib.bt_terrain_create(ctx._h, C.byref(d), C.byref(self._h)))
        # TileAtlasState is per terrain: every TileTree (view) of this atlas shares it (tile_atlas.rs:272-300)
        self._state = C.c_void_p()
        ctx.check(ctx.lib.bt_atlas_state_create(config.atlas_size, C.byref(self._state)))

    @staticmethod
    def new(ctx: Context, config: TerrainConfig) -> "TileAtlas":
        return TileAtlas(ctx, config)

    def upload_tile(self, attachment: int, atlas_index: int, texels: np.ndarray, stream: int = 0) -> None:
        texels = np.ascontiguousarray(texels)
        self.ctx.check(self.ctx.lib.bt_atlas_upload_tile(self._h, attachment, atlas_index, texels.ctypes.data, texels.nbytes, C.c_void_p(stream)))

    def copy_tile(self, attachment: int, atlas_index: int, device_ptr: int, stream: int = 0) -> None:
        a = self.attachments[attachment]
        nbytes = a.texture_size * a.texture_size * a.pixel_size
        self.ctx.check(self.ctx.lib.bt_atlas_copy_tile(self._h, attachment, atlas_index, C.c_void_p(device_ptr), nbytes, C.c_void_p(stream)))

    def load_tile_file(self, attachment: int, atlas_index: int, directory: str, tile, stream: int = 0) -> None:
        """AtlasTileAttachmentWithData::start_loading (tile_atlas.rs:118-149), raw .bin variant"""
        t = F.Tile(*[int(v) for v in tile])
        self.ctx.check(self.ctx.lib.bt_atlas_load_tile_file(self._h, attachment, atlas_index, directory.encode(), C.byref(t), C.c_void_p(stream)))

    def download_tile(self, attachment: int, atlas_index: int) -> np.ndarray:
        dt, shape = self.attachments[attachment].dtype_shape
        out = np.empty(shape, dtype=dt)
        self.ctx.check(self.ctx.lib.bt_atlas_download_tile(self._h, attachment, atlas_index, out.ctypes.data, out.nbytes))
        return out

    def device_ptr(self, attachment: int) -> int:
        p = C.c_void_p()
        self.ctx.check(self.ctx.lib.bt_atlas_device_ptr(self._h, attachment, C.byref(p)))
        return p.value

    def generate_mips(self, attachment: int, atlas_index: int, stream: int = 0) -> None:
        """AttachmentData::generate_mipmaps at tile load (tile_atlas.rs:141), on the GPU"""
        self.ctx.check(self.ctx.lib.bt_atlas_generate_mips(self._h, attachment, atlas_index, C.c_void_p(stream)))

    def update_minmax(self, atlas_index: int, stream: int = 0) -> None:
        """(min, max) table of the height tile in a slot, for BT_FLAG_MINMAX_CULL"""
        self.ctx.check(self.ctx.lib.bt_atlas_update_minmax(self._h, atlas_index, C.c_void_p(stream)))

    def minmax_table(self) -> np.ndarray:
        p = C.c_void_p()
        self.ctx.check(self.ctx.lib.bt_atlas_minmax_device_ptr(self._h, C.byref(p)))
        out = np.empty((self.atlas_size, 65, 2), dtype=np.uint16)
        self.ctx.d2h(out, p.value)
        return out

    def mips_device_ptr(self, attachment: int):
        p, n = C.c_void_p(), C.c_size_t()
        self.ctx.check(self.ctx.lib.bt_atlas_mips_device_ptr(self._h, attachment, C.byref(p), C.byref(n)))
        return p.value, n.value

    def close(self) -> None:
        if self._state:
            self.ctx.lib.bt_atlas_state_destroy(self._state)
            self._state = C.c_void_p()
        if self._h:
            self.ctx.lib.bt_terrain_destroy(self._h)
            self._h = C.c_void_p()


class TileTree:
    """TileTree (tile_tree.rs:94-386) + its render-world TerrainViewData (terrain_view_bind_group.rs:118-187).

    Owns the clipmap residency state (host, f64) and the per-view device buffers; `update` +
    `prepare` produce exactly the uniform/storage contents the reference uploads each frame.
    """

    def __init__(self, tile_atlas: TileAtlas, view_config: F.TerrainViewConfig, max_vertices: int = 0, device_clipmap: bool = False):
        """device_clipmap=True keeps the slots, origins and best-tile table on the GPU (bt_clipmap): compute_requests and
        adjust_to_tile_atlas then run the device twins and a frame uploads no table."""
        self.atlas, self.ctx, self.view_config = tile_atlas, tile_atlas.ctx, view_config
        lib = self.ctx.lib
        self.model = tile_atlas.model
        self.lod_count = tile_atlas.lod_count
        self.side_count = 6 if self.model.is_spherical() else 1
        self.entry_count = self.side_count * self.lod_count * view_config.tree_size ** 2
        self._tree = C.c_void_p()
        self.ctx.check(lib.bt_tile_tree_create_shared(C.byref(self.model.c), C.byref(view_config), self.lod_count, tile_atlas._state, C.byref(self._tree)))
        vd = F.ViewDesc(geometry_tile_count=view_config.geometry_tile_count, tree_size=view_config.tree_size,
                        grid_size=view_config.grid_size, max_vertices=max_vertices)
        self._view = C.c_void_p()
        self.ctx.check(lib.bt_view_create(tile_atlas._h, C.byref(vd), C.byref(self._view)))
        self.view_config_uniform = F.ViewConfigUniform()
        lib.bt_view_config_uniform_from(C.byref(view_config), C.byref(self.model.c), C.byref(self.view_config_uniform))
        self.approximate_height = (self.model.c.min_height + self.model.c.max_height) / 2.0  # tile_tree.rs:155
        self.view_world_position = (0.0, 0.0, 0.0)
        self.entries = (F.TileTreeEntry * self.entry_count)()
        self.origins = (F.u32 * (self.side_count * self.lod_count * 2))()
        self.frame = F.ViewFrame()
        self.h2d_bytes = 0
        self._clipmap = None
        if device_clipmap:
            self._clipmap = C.c_void_p()
            self.ctx.check(lib.bt_clipmap_create(self._view, self._tree, C.byref(self._clipmap)))

    @staticmethod
    def new(tile_atlas: TileAtlas, view_config: F.TerrainViewConfig, max_vertices: int = 0, device_clipmap: bool = False) -> "TileTree":
        return TileTree(tile_atlas, view_config, max_vertices, device_clipmap)

    def download_table(self, stream: int = 0):
        """(entries, origins) as the kernels see them: read back from the device with a clipmap, the host copies otherwise"""
        if self._clipmap is None:
            return np.frombuffer(self.entries, dtype=np.uint32).reshape(-1, 2).copy(), np.frombuffer(self.origins, dtype=np.uint32).copy()
        e = (F.TileTreeEntry * self.entry_count)()
        o = (F.u32 * len(self.origins))()
        self.ctx.check(self.ctx.lib.bt_clipmap_download(self._clipmap, e, self.entry_count, o, len(o), C.c_void_p(stream)))
        return np.frombuffer(e, dtype=np.uint32).reshape(-1, 2).copy(), np.frombuffer(o, dtype=np.uint32).copy()

    def clipmap_transfer_bytes(self):
        h, d = C.c_uint64(), C.c_uint64()
        self.ctx.lib.bt_clipmap_transfer_bytes(self._clipmap, C.byref(h), C.byref(d))
        return h.value, d.value

    # -- main-world systems (plugin.rs:46-56) -------------------------------------------------------
    def add_existing_tiles(self, tiles: np.ndarray) -> None:
        """TileAtlasState::existing_tiles (config.tc contents)."""
        tiles = np.ascontiguousarray(tiles, dtype=np.uint32).reshape(-1, 4)
        self.ctx.check(self.ctx.lib.bt_tile_tree_add_existing(self._tree, tiles.ctypes.data_as(C.POINTER(F.Tile)), len(tiles)))

    def compute_requests(self, view_world_position: Sequence[float]) -> list:
        """TileTree::compute_requests + TileAtlas::update: returns [(tile(side,lod,x,y), atlas_index)] to load."""
        self.view_world_position = tuple(float(v) for v in view_world_position)
        vp = (F.f64 * 3)(*self.view_world_position)
        n = F.u32()
        cap = self.atlas.atlas_size
        loads = (F.TileLoad * cap)()
        if self._clipmap is not None:
            self.ctx.check(self.ctx.lib.bt_clipmap_update(self._clipmap, vp, loads, cap, C.byref(n), None))
        else:
            self.ctx.check(self.ctx.lib.bt_tile_tree_update(self._tree, vp, loads, cap, C.byref(n)))
        return [((l.tile.side, l.tile.lod, l.tile.x, l.tile.y), l.atlas_index) for l in loads[: n.value]]

    def mark_loaded(self, tile) -> None:
        t = F.Tile(*tile)
        self.ctx.check(self.ctx.lib.bt_tile_tree_mark_loaded(self._tree, C.byref(t)))

    def adjust_to_tile_atlas(self) -> None:
        if self._clipmap is not None:
            self.ctx.check(self.ctx.lib.bt_clipmap_entries(self._clipmap, None))
            return
        self.ctx.check(self.ctx.lib.bt_tile_tree_entries(self._tree, self.entries, self.entry_count, self.origins, len(self.origins)))

    def sample_attachment(self, attachment: int, sample_world_position) -> np.ndarray:
        """sample_attachment (terrain_data/mod.rs:267-295): host lookup + device texels, blended with Vec4::lerp"""
        lk = (F.TileLookup * 2)()
        n, ratio = F.u32(), F.f32()
        self.ctx.check(self.ctx.lib.bt_tile_tree_sample_lookup(self._tree, (F.f64 * 3)(*[float(v) for v in sample_world_position]), lk, C.byref(n), C.byref(ratio)))
        out = (F.f32 * 4)()
        self.ctx.check(self.ctx.lib.bt_atlas_sample(self.atlas._h, attachment, C.byref(lk[0]), out))
        value = np.array(out[:], dtype=np.float32)
        if ratio.value > 0.0:
            self.ctx.check(self.ctx.lib.bt_atlas_sample(self.atlas._h, attachment, C.byref(lk[1]), out))
            v2 = np.array(out[:], dtype=np.float32)
            value = (value + (v2 - value) * np.float32(ratio.value)).astype(np.float32)
        return value

    def update_approximate_height(self) -> float:
        """TileTree::approximate_height (tile_tree.rs:376-386): sample_height at the view position"""
        m = self.model.c
        s = self.sample_attachment(0, self.view_world_position)[0]
        h = np.float32(m.min_height) + (np.float32(m.max_height) - np.float32(m.min_height)) * s   # f32::lerp
        self.set_approximate_height(float(h))
        return float(h)

    def set_approximate_height(self, h: float) -> None:
        """TileTree::approximate_height result (tile_tree.rs:376-386); the caller samples it."""
        self.approximate_height = float(h)
        self.ctx.lib.bt_tile_tree_set_approximate_height(self._tree, float(h))

    # -- render-world Prepare (terrain_view_bind_group.rs:238-262, culling_bind_group.rs:87-101) ------
    def prepare(self, stream: int = 0, frame: F.ViewFrame | None = None) -> F.ViewFrame:
        if frame is None:
            f = self.frame
            f.view_config = self.view_config_uniform
            vp = (F.f64 * 3)(*self.view_world_position)
            self.ctx.check(self.ctx.lib.bt_model_approximation_compute(C.byref(self.model.c), vp, self.view_config.origin_lod,
                                                                       self.approximate_height, C.byref(f.approximation)))
            pos32 = [float(np.float32(v)) for v in self.view_world_position]
            f.culling.world_position[:] = pos32     # CullingUniform::from(&ExtractedView)
            f.view_world_position[:] = pos32        # bevy view.world_position
            f.tile_tree_entry_count = self.entry_count
            f.tile_tree = C.cast(self.entries, C.POINTER(F.TileTreeEntry)) if self._clipmap is None else None
            f.origins = C.cast(self.origins, C.POINTER(F.u32))
        else:
            f = frame
        self.ctx.check(self.ctx.lib.bt_view_update(self._view, C.byref(f), C.c_void_p(stream)))
        return f

    def build_frame(self) -> F.ViewFrame:
        """the host half of Prepare: fills self.frame from the tree's current state (no device work)"""
        f = self.frame
        f.view_config = self.view_config_uniform
        vp = (F.f64 * 3)(*self.view_world_position)
        self.ctx.check(self.ctx.lib.bt_model_approximation_compute(C.byref(self.model.c), vp, self.view_config.origin_lod,
                                                                   self.approximate_height, C.byref(f.approximation)))
        pos32 = [float(np.float32(v)) for v in self.view_world_position]
        f.culling.world_position[:] = pos32
        f.view_world_position[:] = pos32
        f.tile_tree_entry_count = self.entry_count
        f.tile_tree = C.cast(self.entries, C.POINTER(F.TileTreeEntry)) if self._clipmap is None else None
        f.origins = C.cast(self.origins, C.POINTER(F.u32))
        return f

    def frame_step(self, gen_flags: int, select_flags: int = 0, stream: int = 0, tile_capacity: int = 8192):
        """one frame through bt_view_run_frame: upload + select + gen_vertices + overlapped tile-list read-back"""
        if getattr(self, "_tile_buf_cap", 0) < tile_capacity:
            self._tile_buf = np.empty((tile_capacity, 4), dtype=np.uint32)
            self._tile_buf_cap = tile_capacity
        n, ind, vb = F.u32(), F.Indirect(), F.VertexBuffer()
        self.ctx.check(self.ctx.lib.bt_view_run_frame(self._view, C.byref(self.build_frame()), select_flags, gen_flags, C.c_void_p(stream),
                                                  self._tile_buf.ctypes.data_as(C.POINTER(F.Tile)), tile_capacity, C.byref(n), C.byref(ind), C.byref(vb)))
        return self._tile_buf[: min(n.value, tile_capacity)], ind, vb

    # -- render graph: TilingPrepassNode::run + the draw's vertex stage -------------------------------
    def select_tiles(self, flags: int = 0, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.bt_select_tiles(self._view, flags, C.c_void_p(stream)))

    def gen_vertices(self, flags: int, stream: int = 0) -> F.VertexBuffer:
        vb = F.VertexBuffer()
        self.ctx.check(self.ctx.lib.bt_gen_vertices(self._view, flags, C.c_void_p(stream), C.byref(vb)))
        return vb

    def readback(self, stream: int = 0):
        cap = self.view_config.geometry_tile_count
        n, ind = F.u32(), F.Indirect()
        self.ctx.check(self.ctx.lib.bt_view_readback(self._view, None, 0, C.byref(n), C.byref(ind), C.c_void_p(stream)))
        tiles = np.empty((min(n.value, cap), 4), dtype=np.uint32)
        if len(tiles):
            self.ctx.check(self.ctx.lib.bt_view_readback(self._view, tiles.ctypes.data_as(C.POINTER(F.Tile)), len(tiles), C.byref(n), C.byref(ind), C.c_void_p(stream)))
        return tiles, ind

    def select_status(self, stream: int = 0) -> int:
        d = F.u32()
        self.ctx.check(self.ctx.lib.bt_view_select_status(self._view, C.byref(d), C.c_void_p(stream)))
        return d.value

    def download_vertices(self, first: int, count: int, stream: int = 0) -> np.ndarray:
        out = np.empty(count, dtype=VERTEX_DTYPE)
        self.ctx.check(self.ctx.lib.bt_view_download_vertices(self._view, out.ctypes.data, first, count, C.c_void_p(stream)))
        return out

    def download_albedo(self, first: int, count: int, stream: int = 0) -> np.ndarray:
        """the per-vertex pack4x8unorm colours of the last gen_vertices(BT_FLAG_ALBEDO)"""
        out = np.empty(count, dtype=np.uint32)
        self.ctx.check(self.ctx.lib.bt_view_download_albedo(self._view, out.ctypes.data, first, count, C.c_void_p(stream)))
        return out

    def close(self) -> None:
        if getattr(self, "_clipmap", None):
            self.ctx.lib.bt_clipmap_destroy(self._clipmap)
            self._clipmap = None
        if self._view:
            self.ctx.lib.bt_view_destroy(self._view)
            self._view = C.c_void_p()
        if self._tree:
            self.ctx.lib.bt_tile_tree_destroy(self._tree)
            self._tree = C.c_void_p()


class ViewBatch:
    """bt_view_batch: the views of one frame refined by ONE select_tiles launch (one CTA per view)."""

    def __init__(self, trees):
        self.trees = list(trees)
        self.ctx = self.trees[0].ctx
        arr = (C.c_void_p * len(self.trees))(*[t._view for t in self.trees])
        self._h = C.c_void_p()
        self.ctx.check(self.ctx.lib.bt_view_batch_create(arr, len(self.trees), C.byref(self._h)))

    def select_tiles(self, flags: int = 0, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.bt_select_tiles_batch(self._h, flags, C.c_void_p(stream)))

    def close(self) -> None:
        if self._h:
            self.ctx.lib.bt_view_batch_destroy(self._h)
            self._h = C.c_void_p()


def load_image(path: str) -> np.ndarray:
    """Decode a PNG / TIFF source image through the library (bt_image_load): (H, W) uint16 for 16-bit gray, (H, W, 4) uint8 for colour."""
    lib = F.load_library()
    img, err = F.Image(), C.create_string_buffer(512)
    st = lib.bt_image_load(os.fsencode(path), C.byref(img), err, len(err))
    if st != F.BT_OK:
        raise F.BtError(st, err.value.decode(errors="replace"))
    try:
        if img.format == F.BT_FORMAT_R16:
            a = np.ctypeslib.as_array(C.cast(img.pixels, C.POINTER(C.c_uint16)), shape=(img.height, img.width)).copy()
        else:
            a = np.ctypeslib.as_array(C.cast(img.pixels, C.POINTER(C.c_uint8)), shape=(img.height, img.width, 4)).copy()
    finally:
        lib.bt_image_free(C.byref(img))
    return a


@dataclass
class PreprocessDataset:  # preprocessor.rs:36-55 — `path` (decoded by the library's PNG / TIFF readers) or decoded `pixels`
    attachment_index: int = 0
    path: str | None = None               # source image file, as in the reference ...
    pixels: np.ndarray | None = None      # ... or a host image (H,W) uint16 / (H,W,4) uint8 ...
    device_pixels: int = 0                # ... or a device pointer with width/height
    width: int = 0
    height: int = 0
    side: int = 0
    top_left: tuple = (0.0, 0.0)
    bottom_right: tuple = (1.0, 1.0)
    lod_range: range = range(0, 1)
    row_begin: int = 0                    # row window of a sharded source (device_pixels points at row_begin)
    row_count: int = 0

    def to_c(self) -> F.Dataset:
        d = F.Dataset()
        d.attachment_index, d.side = self.attachment_index, self.side
        d.top_left[:] = [float(v) for v in self.top_left]
        d.bottom_right[:] = [float(v) for v in self.bottom_right]
        d.lod_begin, d.lod_end = self.lod_range.start, self.lod_range.stop
        if self.pixels is None and self.path is not None and not self.device_pixels:
            self.pixels = load_image(self.path)
        if self.pixels is not None:
            self.pixels = np.ascontiguousarray(self.pixels)
            d.height, d.width = self.pixels.shape[0], self.pixels.shape[1]
            d.pixels_on_device, d.pixels = 0, self.pixels.ctypes.data
        else:
            d.height, d.width = self.height, self.width
            d.pixels_on_device, d.pixels = 1, self.device_pixels
            d.row_begin, d.row_count = self.row_begin, self.row_count
        return d


@dataclass
class SphericalDataset:  # preprocessor.rs:29-34 — one image per cube face
    attachment_index: int = 0
    faces: list = field(default_factory=list)   # six host images ...
    device_faces: list = field(default_factory=list)   # ... or six device pointers with width/height
    width: int = 0
    height: int = 0
    lod_range: range = range(0, 1)


class Preprocessor:
    """Preprocessor + TerrainPreprocessPlugin's render node (preprocessor.rs:214-343, preprocess/mod.rs:143-217)."""

    def __init__(self, ctx: Context, config: TerrainConfig, rank: int = 0, world: int = 1):
        self.ctx, self.config = ctx, config
        self.attachments = list(config.attachments)
        arr = (F.AttachmentConfig * len(self.attachments))(*[a.to_c() for a in self.attachments])
        kind = F.BT_KIND_SPHERICAL if config.model.is_spherical() else F.BT_KIND_PLANAR
        self._h = C.c_void_p()
        ctx.check(ctx.lib.bt_preprocessor_create(ctx._h, kind, arr, len(self.attachments), config.lod_count, F.Shard(rank, world), C.byref(self._h)))
        self._keep = []
        self.rank, self.world = rank, world

    @staticmethod
    def new(ctx: Context, config: TerrainConfig, rank: int = 0, world: int = 1) -> "Preprocessor":
        return Preprocessor(ctx, config, rank, world)

    def clear_attachment(self, attachment_index: int) -> "Preprocessor":
        """preprocessor.rs:290-296 clears the on-disk data; the device pyramid of a new Preprocessor is empty."""
        return self

    def preprocess_tile(self, dataset: PreprocessDataset) -> "Preprocessor":
        c = dataset.to_c()
        self._keep.append((dataset, c))
        self.ctx.check(self.ctx.lib.bt_preprocessor_add_datasets(self._h, C.byref(c), 1))
        return self

    def preprocess_spherical(self, dataset: SphericalDataset) -> "Preprocessor":
        if dataset.device_faces:
            ds = [PreprocessDataset(attachment_index=dataset.attachment_index, device_pixels=dataset.device_faces[s], width=dataset.width,
                                    height=dataset.height, side=s, lod_range=dataset.lod_range) for s in range(6)]
        else:
            ds = [PreprocessDataset(attachment_index=dataset.attachment_index, pixels=dataset.faces[s], side=s, lod_range=dataset.lod_range) for s in range(6)]
        arr = (F.Dataset * 6)(*[d.to_c() for d in ds])
        self._keep.append((ds, arr))
        self.ctx.check(self.ctx.lib.bt_preprocessor_add_datasets(self._h, arr, 6))
        return self

    def reserve(self, stream: int = 0) -> "Preprocessor":
        """allocate the device pyramid store up front (keeps cudaMalloc out of run)"""
        self.ctx.check(self.ctx.lib.bt_preprocessor_reserve(self._h, C.c_void_p(stream)))
        return self

    def run(self, stream: int = 0) -> "Preprocessor":
        self.ctx.check(self.ctx.lib.bt_preprocessor_run(self._h, C.c_void_p(stream)))
        return self

    def run_sharded(self, comm: "Comm", stream: int = 0) -> "Preprocessor":
        """the whole sharded job in the library: kernels + ONE grouped NCCL all-gather per wave, overlapped with the interior stitch"""
        self.ctx.check(self.ctx.lib.bt_preprocessor_run_sharded(self._h, comm._h, C.c_void_p(stream)))
        return self

    # phase API (hosts that run their own collective between the phases; also what the CPU gloo tests drive)
    def run_split_downsample(self, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.bt_preprocessor_run_split_downsample(self._h, C.c_void_p(stream)))

    def run_stitch(self, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.bt_preprocessor_run_stitch(self._h, C.c_void_p(stream)))

    def halo_bytes(self) -> int:
        n = F.u64()
        self.ctx.check(self.ctx.lib.bt_preprocessor_halo_bytes(self._h, C.byref(n)))
        return n.value

    def roots_bytes(self) -> int:
        n = F.u64()
        self.ctx.check(self.ctx.lib.bt_preprocessor_roots_bytes(self._h, C.byref(n)))
        return n.value

    def halo_pack(self, send_ptr: int, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.bt_preprocessor_halo_pack(self._h, C.c_void_p(send_ptr), C.c_void_p(stream)))

    def halo_unpack(self, gathered_ptr: int, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.bt_preprocessor_halo_unpack(self._h, C.c_void_p(gathered_ptr), C.c_void_p(stream)))

    def roots_pack(self, send_ptr: int, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.bt_preprocessor_roots_pack(self._h, C.c_void_p(send_ptr), C.c_void_p(stream)))

    def roots_unpack_and_finish(self, gathered_ptr: int, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.bt_preprocessor_roots_unpack_and_finish(self._h, C.c_void_p(gathered_ptr), C.c_void_p(stream)))

    # results
    def tiles(self, attachment: int = 0) -> np.ndarray:
        n = F.u32()
        self.ctx.check(self.ctx.lib.bt_preprocessor_tile_count(self._h, attachment, C.byref(n)))
        out = np.empty((n.value, 4), dtype=np.uint32)
        if n.value:
            self.ctx.check(self.ctx.lib.bt_preprocessor_list_tiles(self._h, attachment, out.ctypes.data_as(C.POINTER(F.Tile)), n.value, C.byref(n)))
        return out

    def tile_device_ptr(self, attachment: int, tile) -> int:
        t, p = F.Tile(*[int(v) for v in tile]), C.c_void_p()
        self.ctx.check(self.ctx.lib.bt_preprocessor_tile_device_ptr(self._h, attachment, C.byref(t), C.byref(p)))
        return p.value

    def get_tile(self, attachment: int, tile) -> np.ndarray:
        dt, shape = self.attachments[attachment].dtype_shape
        out = np.empty(shape, dtype=dt)
        t = F.Tile(*[int(v) for v in tile])
        self.ctx.check(self.ctx.lib.bt_preprocessor_get_tile(self._h, attachment, C.byref(t), out.ctypes.data, out.nbytes))
        return out

    def content_hash(self, attachment: int = 0, stream: int = 0) -> int:
        """order-independent 64-bit hash of this rank's output tiles (sum over ranks == single-GPU hash iff byte-identical)"""
        h = F.u64()
        self.ctx.check(self.ctx.lib.bt_preprocessor_hash(self._h, attachment, C.byref(h), C.c_void_p(stream)))
        return h.value

    def save(self, attachment: int, directory: str) -> None:
        self.ctx.check(self.ctx.lib.bt_preprocessor_save(self._h, attachment, directory.encode()))

    def close(self) -> None:
        if self._h:
            self.ctx.lib.bt_preprocessor_destroy(self._h)
            self._h = C.c_void_p()


def load_tile_config(path: str) -> np.ndarray:
    """TileAtlas::load_tile_config (tile_atlas.rs:616-623): the tiles listed in config.tc as an (n, 4) uint32 array"""
    lib = F.load_library()
    n = F.u32()
    if lib.bt_tile_config_load(path.encode(), None, 0, C.byref(n)) != F.BT_OK:
        raise F.BtError(F.BT_ERR_IO, f"cannot read {path}")
    out = np.empty((n.value, 4), dtype=np.uint32)
    if n.value:
        lib.bt_tile_config_load(path.encode(), out.ctypes.data_as(C.POINTER(F.Tile)), n.value, C.byref(n))
    return out


def load_resident_tiles(tile_tree: TileTree, preprocessor: Preprocessor, view_world_position, attachments: Iterable[int] = (0,), stream: int = 0) -> int:
    """One frame of the streaming systems with the preprocessor's device pyramid as the tile source
    (stands in for the async file loads of tile_atlas.rs:118-149): compute_requests -> copy the
    requested tiles into their atlas slots -> adjust_to_tile_atlas.  Returns the number of loads."""
    loads = tile_tree.compute_requests(view_world_position)
    for tile, atlas_index in loads:
        for a in attachments:
            tile_tree.atlas.copy_tile(a, atlas_index, preprocessor.tile_device_ptr(a, tile), stream)
            if tile_tree.atlas.attachments[a].mip_level_count > 1:
                tile_tree.atlas.generate_mips(a, atlas_index, stream)     # the CPU mip generation at tile load, on the GPU
            if a == 0:
                tile_tree.atlas.update_minmax(atlas_index, stream)        # height bounds of the slot (min / max-height cull)
        tile_tree.mark_loaded(tile)
    tile_tree.adjust_to_tile_atlas()
    return len(loads)

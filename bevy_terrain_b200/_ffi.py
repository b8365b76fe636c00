"""ctypes mirrors of the PODs in include/bevy_terrain_b200.h and the loader of the C-ABI library.

The library is the product: there is NO Python/CPU fallback.  If ``libbevy_terrain_b200.so`` is
missing the import fails loudly (build it with ``python __graft_entry__.py`` or
``make -C bevy_terrain_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BT_LIB_PATH lets kernel A/B experiments point at an alternative build of the SAME library; there is still no fallback
LIB_PATH = os.environ.get("BT_LIB_PATH") or os.path.join(_HERE, "csrc", "libbevy_terrain_b200.so")

u32, i32, f32, f64, u64 = C.c_uint32, C.c_int32, C.c_float, C.c_double, C.c_uint64

# status codes
BT_OK, BT_ERR_INVALID_ARG, BT_ERR_CAPACITY, BT_ERR_CUDA, BT_ERR_NCCL, BT_ERR_IO, BT_ERR_OUT_OF_MEMORY, BT_ERR_NOT_FOUND = range(8)
BT_FORMAT_RGBA8, BT_FORMAT_R16 = 0, 1
BT_KIND_PLANAR, BT_KIND_SPHERICAL, BT_KIND_ELLIPSOIDAL = 0, 1, 2
BT_FLAG_MORPH, BT_FLAG_BLEND, BT_FLAG_HIGH_PRECISION, BT_FLAG_NORMALS, BT_FLAG_FRUSTUM_CULL = 1 << 1, 1 << 2, 1 << 3, 1 << 4, 1 << 6
BT_FLAG_FILTER_TRILINEAR_ANISO, BT_FLAG_ALBEDO, BT_FLAG_MINMAX_CULL = 1 << 7, 1 << 8, 1 << 9


class Tile(C.Structure):  # bt_tile / TileCoordinate (coordinate.rs:156-166)
    _fields_ = [("side", u32), ("lod", u32), ("x", u32), ("y", u32)]

    def key(self):
        return (self.side, self.lod, self.x, self.y)


class Indirect(C.Structure):
    _fields_ = [("vertex_count", u32), ("instance_count", u32), ("base_vertex", u32), ("base_instance", u32)]


class ViewConfigUniform(C.Structure):  # terrain_view_bind_group.rs:82-116
    _fields_ = [("tree_size", u32), ("geometry_tile_count", u32), ("refinement_count", u32), ("grid_size", f32),
                ("vertices_per_row", u32), ("vertices_per_tile", u32), ("morph_distance", f32),
                ("blend_distance", f32), ("load_distance", f32), ("subdivision_distance", f32),
                ("morph_range", f32), ("blend_range", f32), ("precision_threshold_distance", f32)]


class SideParameter(C.Structure):  # terrain_model.rs:228-249
    _fields_ = [("view_xy", i32 * 2), ("view_uv", f32 * 2),
                ("c", f32 * 3), ("_p0", f32), ("c_s", f32 * 3), ("_p1", f32), ("c_t", f32 * 3), ("_p2", f32),
                ("c_ss", f32 * 3), ("_p3", f32), ("c_st", f32 * 3), ("_p4", f32), ("c_tt", f32 * 3), ("_p5", f32)]


class ModelApproximation(C.Structure):  # terrain_model.rs:252-259
    _fields_ = [("origin_lod", u32), ("approximate_height", f32), ("_pad", u32 * 2), ("sides", SideParameter * 6)]


class CullingUniform(C.Structure):  # culling_bind_group.rs:41-55
    _fields_ = [("world_position", f32 * 3), ("_p0", f32), ("view_proj", f32 * 16), ("planes", (f32 * 4) * 5)]


class TerrainConfigUniform(C.Structure):  # terrain_bind_group.rs:77-93
    _fields_ = [("lod_count", u32), ("min_height", f32), ("max_height", f32), ("scale", f32)]


class AttachmentUniform(C.Structure):  # terrain_bind_group.rs:47-72
    _fields_ = [("size", f32), ("scale", f32), ("offset", f32), ("_pad", u32)]


class TileTreeEntry(C.Structure):  # tile_tree.rs:51-65
    _fields_ = [("atlas_index", u32), ("atlas_lod", u32)]


class MeshTransform(C.Structure):
    _fields_ = [("world_from_local", f32 * 12), ("normal_matrix", f32 * 9), ("_pad", f32 * 3)]


class Vertex(C.Structure):  # vertex.wgsl:14-20 minus clip_position
    _fields_ = [("position", f32 * 4), ("normal", f32 * 3), ("uv", f32 * 2), ("tile_index", u32)]


class AttachmentConfig(C.Structure):  # terrain_data/mod.rs:88-109
    _fields_ = [("texture_size", u32), ("border_size", u32), ("mip_level_count", u32), ("format", u32)]


class TerrainDesc(C.Structure):
    _fields_ = [("kind", u32), ("lod_count", u32), ("atlas_size", u32), ("attachment_count", u32),
                ("attachments", AttachmentConfig * 8), ("min_height", f32), ("max_height", f32),
                ("scale", f64), ("transform", MeshTransform)]


class ViewFrame(C.Structure):
    _fields_ = [("view_config", ViewConfigUniform), ("approximation", ModelApproximation),
                ("culling", CullingUniform), ("view_world_position", f32 * 3),
                ("tile_tree_entry_count", u32), ("tile_tree", C.POINTER(TileTreeEntry)),
                ("origins", C.POINTER(u32))]


class VertexBuffer(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("capacity", u64)]


class ViewDesc(C.Structure):
    _fields_ = [("geometry_tile_count", u32), ("tree_size", u32), ("grid_size", u32), ("readback_tile_capacity", u32), ("max_vertices", u64)]


class TerrainModel(C.Structure):  # terrain_model.rs:40-104
    _fields_ = [("kind", u32), ("_pad", u32), ("position", f64 * 3), ("scale", f64 * 3),
                ("min_height", f32), ("max_height", f32)]


class TerrainViewConfig(C.Structure):  # terrain_view.rs:19-64
    _fields_ = [("tree_size", u32), ("geometry_tile_count", u32), ("refinement_count", u32), ("grid_size", u32),
                ("subdivision_tolerance", f64), ("precision_threshold_distance", f64), ("load_distance", f64),
                ("morph_distance", f64), ("blend_distance", f64), ("morph_range", f32), ("blend_range", f32),
                ("origin_lod", u32), ("_pad", u32)]


class TileLoad(C.Structure):
    _fields_ = [("tile", Tile), ("atlas_index", u32), ("_pad", u32 * 3)]


class Image(C.Structure):  # a decoded source image (bt_image_load)
    _fields_ = [("width", u32), ("height", u32), ("format", u32), ("_pad", u32), ("pixels", C.c_void_p)]


class TileLookup(C.Structure):  # tile_tree.rs:67-83
    _fields_ = [("atlas_index", u32), ("atlas_lod", u32), ("atlas_uv", f32 * 2)]


class Dataset(C.Structure):  # preprocessor.rs:36-55
    _fields_ = [("attachment_index", u32), ("side", u32), ("top_left", f32 * 2), ("bottom_right", f32 * 2),
                ("lod_begin", u32), ("lod_end", u32), ("width", u32), ("height", u32),
                ("pixels_on_device", u32), ("row_begin", u32), ("row_count", u32), ("_pad", u32), ("pixels", C.c_void_p)]


class ExchangeCopy(C.Structure):
    _fields_ = [("tile", Tile), ("tile_row", u32), ("tile_col", u32), ("pixels", u32), ("src_rank", u32), ("buffer_offset", u64)]


class Shard(C.Structure):
    _fields_ = [("rank", u32), ("world", u32)]


_EXPECTED_SIZES = {Tile: 16, Indirect: 16, ViewConfigUniform: 52, SideParameter: 112, ModelApproximation: 688,
                   CullingUniform: 160, TerrainConfigUniform: 16, AttachmentUniform: 16, TileTreeEntry: 8,
                   MeshTransform: 96, Vertex: 40}
for _t, _n in _EXPECTED_SIZES.items():
    assert C.sizeof(_t) == _n, (_t.__name__, C.sizeof(_t), _n)


class BtError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"bt_status={status}: {message}")
        self.status = status


_lib = None


def load_library() -> C.CDLL:
    """Loads the C-ABI shared library (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA/C-ABI library is the product and has no fallback. "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(lib: C.CDLL) -> None:
    P, vp = C.POINTER, C.c_void_p
    sigs = {
        "bt_abi_version": (u32, []),
        "bt_context_create": (i32, [i32, P(vp)]),
        "bt_context_destroy": (None, [vp]),
        "bt_last_error": (C.c_char_p, [vp]),
        "bt_context_synchronize": (i32, [vp]),
        "bt_terrain_create": (i32, [vp, P(TerrainDesc), P(vp)]),
        "bt_terrain_destroy": (None, [vp]),
        "bt_atlas_upload_tile": (i32, [vp, u32, u32, vp, C.c_size_t, vp]),
        "bt_atlas_copy_tile": (i32, [vp, u32, u32, vp, C.c_size_t, vp]),
        "bt_atlas_download_tile": (i32, [vp, u32, u32, vp, C.c_size_t]),
        "bt_atlas_load_tile_file": (i32, [vp, u32, u32, C.c_char_p, P(Tile), vp]),
        "bt_tile_config_load": (i32, [C.c_char_p, P(Tile), u32, P(u32)]),
        "bt_image_load": (i32, [C.c_char_p, P(Image), C.c_char_p, C.c_size_t]),
        "bt_image_free": (None, [P(Image)]),
        "bt_atlas_device_ptr": (i32, [vp, u32, P(vp)]),
        "bt_atlas_generate_mips": (i32, [vp, u32, u32, vp]),
        "bt_atlas_update_minmax": (i32, [vp, u32, vp]),
        "bt_atlas_minmax_device_ptr": (i32, [vp, P(vp)]),
        "bt_atlas_mips_device_ptr": (i32, [vp, u32, P(vp), P(C.c_size_t)]),
        "bt_view_albedo_buffer": (i32, [vp, P(vp), P(u64)]),
        "bt_view_download_albedo": (i32, [vp, vp, u64, u64, vp]),
        "bt_view_create": (i32, [vp, P(ViewDesc), P(vp)]),
        "bt_view_destroy": (None, [vp]),
        "bt_view_update": (i32, [vp, P(ViewFrame), vp]),
        "bt_select_tiles": (i32, [vp, u32, vp]),
        "bt_gen_vertices": (i32, [vp, u32, vp, P(VertexBuffer)]),
        "bt_view_batch_create": (i32, [P(vp), u32, P(vp)]),
        "bt_view_batch_destroy": (None, [vp]),
        "bt_select_tiles_batch": (i32, [vp, u32, vp]),
        "bt_view_run_frame": (i32, [vp, P(ViewFrame), u32, u32, vp, P(Tile), u32, P(u32), P(Indirect), P(VertexBuffer)]),
        "bt_view_readback": (i32, [vp, P(Tile), u32, P(u32), P(Indirect), vp]),
        "bt_view_select_status": (i32, [vp, P(u32), vp]),
        "bt_view_download_vertices": (i32, [vp, vp, u64, u64, vp]),
        "bt_view_device_ptrs": (i32, [vp, P(vp), P(vp)]),
        "bt_terrain_model_planar": (None, [P(f64), f64, f32, f32, P(TerrainModel)]),
        "bt_terrain_model_sphere": (None, [P(f64), f64, f32, f32, P(TerrainModel)]),
        "bt_terrain_model_ellipsoid": (None, [P(f64), f64, f64, f32, f32, P(TerrainModel)]),
        "bt_terrain_view_config_default": (None, [P(TerrainViewConfig)]),
        "bt_terrain_model_scale": (f64, [P(TerrainModel)]),
        "bt_terrain_model_mesh_transform": (None, [P(TerrainModel), P(MeshTransform)]),
        "bt_view_config_uniform_from": (None, [P(TerrainViewConfig), P(TerrainModel), P(ViewConfigUniform)]),
        "bt_model_approximation_compute": (i32, [P(TerrainModel), P(f64), u32, f32, P(ModelApproximation)]),
        "bt_tile_tree_create": (i32, [P(TerrainModel), P(TerrainViewConfig), u32, u32, P(vp)]),
        "bt_atlas_state_create": (i32, [u32, P(vp)]),
        "bt_atlas_state_destroy": (None, [vp]),
        "bt_tile_tree_create_shared": (i32, [P(TerrainModel), P(TerrainViewConfig), u32, vp, P(vp)]),
        "bt_tile_tree_destroy": (None, [vp]),
        "bt_tile_tree_add_existing": (i32, [vp, P(Tile), u32]),
        "bt_tile_tree_update": (i32, [vp, P(f64), P(TileLoad), u32, P(u32)]),
        "bt_tile_tree_mark_loaded": (i32, [vp, P(Tile)]),
        "bt_tile_tree_entries": (i32, [vp, P(TileTreeEntry), u32, P(u32), u32]),
        "bt_clipmap_create": (i32, [vp, vp, P(vp)]),
        "bt_clipmap_destroy": (None, [vp]),
        "bt_clipmap_update": (i32, [vp, P(f64), P(TileLoad), u32, P(u32), vp]),
        "bt_clipmap_entries": (i32, [vp, vp]),
        "bt_clipmap_download": (i32, [vp, P(TileTreeEntry), u32, P(u32), u32, vp]),
        "bt_clipmap_transfer_bytes": (None, [vp, P(C.c_uint64), P(C.c_uint64)]),
        "bt_tile_tree_set_approximate_height": (None, [vp, f32]),
        "bt_tile_tree_sample_lookup": (i32, [vp, P(f64), P(TileLookup), P(u32), P(f32)]),
        "bt_atlas_sample": (i32, [vp, u32, P(TileLookup), P(f32)]),
        "bt_preprocessor_create": (i32, [vp, u32, P(AttachmentConfig), u32, u32, Shard, P(vp)]),
        "bt_preprocessor_destroy": (None, [vp]),
        "bt_preprocessor_add_datasets": (i32, [vp, P(Dataset), u32]),
        "bt_preprocessor_reserve": (i32, [vp, vp]),
        "bt_preprocessor_run_split_downsample": (i32, [vp, vp]),
        "bt_preprocessor_run_stitch": (i32, [vp, vp]),
        "bt_preprocessor_run": (i32, [vp, vp]),
        "bt_preprocessor_plan": (i32, [vp]),
        "bt_preprocessor_plan_tiles": (i32, [vp, u32, P(Tile), P(u32), u32, P(u32)]),
        "bt_preprocessor_plan_exchange": (i32, [vp, u32, u32, P(ExchangeCopy), u32, P(u32)]),
        "bt_preprocessor_halo_bytes": (i32, [vp, P(u64)]),
        "bt_preprocessor_halo_pack": (i32, [vp, vp, vp]),
        "bt_preprocessor_halo_unpack": (i32, [vp, vp, vp]),
        "bt_preprocessor_roots_bytes": (i32, [vp, P(u64)]),
        "bt_preprocessor_roots_pack": (i32, [vp, vp, vp]),
        "bt_preprocessor_roots_unpack_and_finish": (i32, [vp, vp, vp]),
        "bt_preprocessor_tile_count": (i32, [vp, u32, P(u32)]),
        "bt_preprocessor_list_tiles": (i32, [vp, u32, P(Tile), u32, P(u32)]),
        "bt_preprocessor_tile_device_ptr": (i32, [vp, u32, P(Tile), P(vp)]),
        "bt_preprocessor_get_tile": (i32, [vp, u32, P(Tile), vp, C.c_size_t]),
        "bt_preprocessor_save": (i32, [vp, u32, C.c_char_p]),
        "bt_preprocessor_hash": (i32, [vp, u32, P(u64), vp]),
        "bt_comm_unique_id": (i32, [vp, C.c_size_t]),
        "bt_comm_create": (i32, [vp, u32, u32, vp, C.c_size_t, P(vp)]),
        "bt_comm_destroy": (None, [vp]),
        "bt_comm_info": (i32, [vp, P(u32), P(u32), P(i32), P(u32)]),
        "bt_comm_barrier": (i32, [vp, vp]),
        "bt_comm_all_reduce_sum_u64": (i32, [vp, P(u64), u32, vp]),
        "bt_comm_load_error": (C.c_char_p, []),
        "bt_preprocessor_run_sharded": (i32, [vp, vp, vp]),
        "bt_culling_planes": (None, [P(f32), P(f32 * 4)]),
        "bt_generate_mipmaps": (i32, [vp, u32, u32, u32, vp, vp, vp]),
        "bt_synth_source": (i32, [vp, u32, u32, u32, u32, u32, u32, u32, vp, vp]),
        "bt_selftest_numerics": (i32, [vp, P(f32), P(f32), u32, P(u32)]),
        "bt_device_alloc": (i32, [vp, C.c_size_t, P(vp)]),
        "bt_device_free": (i32, [vp, vp]),
        "bt_memcpy_h2d": (i32, [vp, vp, vp, C.c_size_t, vp]),
        "bt_memcpy_d2h": (i32, [vp, vp, vp, C.c_size_t, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args


DECLARED_SYMBOLS = None  # filled lazily by tests from the header

//! Raw `extern "C"` declarations mirroring `include/bevy_terrain_b200.h` (ABI version 1).
//! Struct layouts equal the encase std140/std430 layouts the reference uploads, so the reference's own
//! `ShaderType` values can be byte-copied into them.  NOT COMPILED in this repository (no Rust toolchain).
#![allow(non_camel_case_types)]
use bytemuck::{Pod, Zeroable};
use std::os::raw::{c_char, c_void};

pub type bt_status = i32;
pub const BT_OK: bt_status = 0;
pub const BT_ERR_INVALID_ARG: bt_status = 1;
pub const BT_ERR_CAPACITY: bt_status = 2;
pub const BT_ERR_CUDA: bt_status = 3;
pub const BT_ERR_NCCL: bt_status = 4;
pub const BT_ERR_IO: bt_status = 5;
pub const BT_ERR_OUT_OF_MEMORY: bt_status = 6;
pub const BT_ERR_NOT_FOUND: bt_status = 7;

pub const BT_FORMAT_RGBA8: u32 = 0; // AttachmentFormat::id(), src/terrain_data/mod.rs:50-57
pub const BT_FORMAT_R16: u32 = 1;
pub const BT_KIND_PLANAR: u32 = 0;
pub const BT_KIND_SPHERICAL: u32 = 1;
pub const BT_KIND_ELLIPSOIDAL: u32 = 2; // host-side model only (bt_terrain_model)
pub const BT_FLAG_MORPH: u32 = 1 << 1;
pub const BT_FLAG_BLEND: u32 = 1 << 2;
pub const BT_FLAG_HIGH_PRECISION: u32 = 1 << 3;
pub const BT_FLAG_NORMALS: u32 = 1 << 4;
pub const BT_FLAG_FRUSTUM_CULL: u32 = 1 << 6; // extension, default off (the reference does not cull at this commit)
pub const BT_FLAG_FILTER_TRILINEAR_ANISO: u32 = 1 << 7; // configs[2] sampler: linear / linear / linear + anisotropy 16 per vertex
pub const BT_FLAG_ALBEDO: u32 = 1 << 8;          // per-vertex albedo stream (bt_view_albedo_buffer)
pub const BT_FLAG_MINMAX_CULL: u32 = 1 << 9;     // frustum cull with per-tile (min, max) height bounds

macro_rules! opaque { ($($n:ident),*) => { $( #[repr(C)] pub struct $n { _p: [u8; 0] } )* } }
opaque!(bt_context, bt_terrain, bt_view, bt_preprocessor, bt_tile_tree, bt_atlas_state, bt_comm, bt_view_batch, bt_clipmap);
pub const BT_COMM_UNIQUE_ID_BYTES: usize = 128;

/// `TileCoordinate` (src/math/coordinate.rs:156-166)
#[repr(C)] #[derive(Clone, Copy, Debug, Default, PartialEq, Eq, Hash, Pod, Zeroable)]
pub struct bt_tile { pub side: u32, pub lod: u32, pub x: u32, pub y: u32 }

/// `Indirect` (src/render/terrain_view_bind_group.rs:66-71)
#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_indirect { pub vertex_count: u32, pub instance_count: u32, pub base_vertex: u32, pub base_instance: u32 }

/// `TerrainViewConfigUniform` (terrain_view_bind_group.rs:82-116), 52 B
#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_view_config_uniform {
    pub tree_size: u32, pub geometry_tile_count: u32, pub refinement_count: u32, pub grid_size: f32,
    pub vertices_per_row: u32, pub vertices_per_tile: u32, pub morph_distance: f32, pub blend_distance: f32,
    pub load_distance: f32, pub subdivision_distance: f32, pub morph_range: f32, pub blend_range: f32,
    pub precision_threshold_distance: f32,
}

/// `SideParameter` (src/math/terrain_model.rs:228-249), 112 B
#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_side_parameter {
    pub view_xy: [i32; 2], pub view_uv: [f32; 2],
    pub c: [f32; 3], pub _p0: f32, pub c_s: [f32; 3], pub _p1: f32, pub c_t: [f32; 3], pub _p2: f32,
    pub c_ss: [f32; 3], pub _p3: f32, pub c_st: [f32; 3], pub _p4: f32, pub c_tt: [f32; 3], pub _p5: f32,
}

/// `TerrainModelApproximation` (terrain_model.rs:252-259), 688 B
#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_model_approximation { pub origin_lod: u32, pub approximate_height: f32, pub _pad: [u32; 2], pub sides: [bt_side_parameter; 6] }

/// `CullingUniform` (src/render/culling_bind_group.rs:41-55), 160 B
#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_culling_uniform { pub world_position: [f32; 3], pub _p0: f32, pub view_proj: [f32; 16], pub planes: [[f32; 4]; 5] }

/// `TileTreeEntry` (src/terrain_data/tile_tree.rs:51-65)
#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_tile_tree_entry { pub atlas_index: u32, pub atlas_lod: u32 }

#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_mesh_transform { pub world_from_local: [f32; 12], pub normal_matrix: [f32; 9], pub _pad: [f32; 3] }

/// `VertexOutput` minus clip_position (src/shaders/render/vertex.wgsl:14-20), 40 B
#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_vertex { pub position: [f32; 4], pub normal: [f32; 3], pub uv: [f32; 2], pub tile_index: u32 }

#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct bt_attachment_config { pub texture_size: u32, pub border_size: u32, pub mip_level_count: u32, pub format: u32 }

#[repr(C)] #[derive(Clone, Copy)]
pub struct bt_terrain_desc {
    pub kind: u32, pub lod_count: u32, pub atlas_size: u32, pub attachment_count: u32,
    pub attachments: [bt_attachment_config; 8], pub min_height: f32, pub max_height: f32, pub scale: f64,
    pub transform: bt_mesh_transform,
}

#[repr(C)] #[derive(Clone, Copy)]
pub struct bt_view_desc { pub geometry_tile_count: u32, pub tree_size: u32, pub grid_size: u32, pub readback_tile_capacity: u32, pub max_vertices: u64 }

#[repr(C)]
pub struct bt_view_frame {
    pub view_config: bt_view_config_uniform, pub approximation: bt_model_approximation, pub culling: bt_culling_uniform,
    pub view_world_position: [f32; 3], pub tile_tree_entry_count: u32, pub tile_tree: *const bt_tile_tree_entry, pub origins: *const u32,
}

#[repr(C)] pub struct bt_vertex_buffer { pub vertices: *mut bt_vertex, pub capacity: u64 }

/// `PreprocessDataset` (src/preprocess/preprocessor.rs:36-55) with decoded pixels instead of a path
#[repr(C)]
pub struct bt_dataset {
    pub attachment_index: u32, pub side: u32, pub top_left: [f32; 2], pub bottom_right: [f32; 2], pub lod_begin: u32, pub lod_end: u32,
    pub width: u32, pub height: u32, pub pixels_on_device: u32, pub row_begin: u32, pub row_count: u32, pub _pad: u32, pub pixels: *const c_void,
}

#[repr(C)] #[derive(Clone, Copy)] pub struct bt_shard { pub rank: u32, pub world: u32 }

#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_terrain_config_uniform { pub lod_count: u32, pub min_height: f32, pub max_height: f32, pub scale: f32 }   // terrain_bind_group.rs:77-93

#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_attachment_uniform { pub size: f32, pub scale: f32, pub offset: f32, pub _pad: u32 }                     // terrain_bind_group.rs:47-72

/// result of the CPU-side `lookup_tile` twin (tile_tree.rs:239-266): which atlas layer to sample and where
#[repr(C)] #[derive(Clone, Copy, Debug, Default, Pod, Zeroable)]
pub struct bt_tile_lookup { pub atlas_index: u32, pub atlas_lod: u32, pub atlas_uv: [f32; 2] }

/// `TerrainModel` (src/math/terrain_model.rs:61-128), f64 host side
#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct bt_terrain_model { pub kind: u32, pub _pad: u32, pub position: [f64; 3], pub scale: [f64; 3], pub min_height: f32, pub max_height: f32 }

/// `TerrainViewConfig` (src/terrain_view.rs:19-64)
#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct bt_terrain_view_config {
    pub tree_size: u32, pub geometry_tile_count: u32, pub refinement_count: u32, pub grid_size: u32,
    pub subdivision_tolerance: f64, pub precision_threshold_distance: f64, pub load_distance: f64, pub morph_distance: f64, pub blend_distance: f64,
    pub morph_range: f32, pub blend_range: f32, pub origin_lod: u32, pub _pad: u32,
}

/// one tile the clipmap asks the host to load into `atlas_index` (TileAtlas::update, tile_atlas.rs:574-601)
#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct bt_tile_load { pub tile: bt_tile, pub atlas_index: u32, pub _pad: [u32; 3] }

/// one segment of the multi-GPU halo / sub-root exchange plan (bt_preprocessor_plan_exchange)
#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct bt_exchange_copy { pub tile: bt_tile, pub tile_row: u32, pub tile_col: u32, pub pixels: u32, pub src_rank: u32, pub buffer_offset: u64 }

/// a decoded source image (`bt_image_load`): `pixels` is malloc'ed by the library, release it with `bt_image_free`
#[repr(C)] pub struct bt_image { pub width: u32, pub height: u32, pub format: u32, pub _pad: u32, pub pixels: *mut c_void }

// Every entry point of include/bevy_terrain_b200.h: GENERATED by `python tools/gen_rust_ffi.py --write` into ffi_extern.rs
// (tests/test_abi.py regenerates it and fails when the committed file is stale).
include!("ffi_extern.rs");

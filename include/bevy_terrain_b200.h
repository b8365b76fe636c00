/*
 * bevy_terrain_b200 — C ABI of the B200-native (sm_100a CUDA) implementation of
 * bevy_terrain's data-parallel hot path:
 *
 *   (1) per-frame quadtree tile selection + tile-list / indirect-draw emit
 *   (2) per-tile heightmap-sampled vertex-grid generation (positions + normals)
 *   (3) the offline tile-pyramid preprocessor (split / downsample / stitch)
 *
 * The reference (kurtkuehnert/bevy_terrain @ 5a88daf) has NO FFI boundary: the hot path sits
 * behind three Bevy render-graph call points and a set of plain-data uniform/storage structs.
 * Every entry point below names the reference call point (file:line under /root/reference)
 * it replaces.  All structs are little-endian PODs whose byte layout equals the encase
 * std140/std430 layout the reference uploads, so the Rust side can `bytemuck::cast` them.
 *
 * Conventions
 *   - every function returns a bt_status (0 = BT_OK); nothing throws/unwinds across the ABI
 *   - opaque handles own all device memory; host input STRUCTS are copied during the call (bulk texel
 *     uploads are stream-ordered copies from the caller's buffer: see bt_atlas_upload_tile)
 *   - `stream` arguments are a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *     per-frame calls are stream-ordered and non-blocking except the *_readback calls
 *   - no allocation on the per-frame path after bt_view_create
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *     with BT_ERR_CUDA (the CPU oracle under oracle/ is test infrastructure only)
 */
#ifndef BEVY_TERRAIN_B200_H
#define BEVY_TERRAIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BT_ABI_VERSION 1u

/* ------------------------------------------------------------------------------------------ */
/* status codes (reference: panics via unwrap/expect, tile_atlas.rs:384; silent frame skip,    */
/* tiling_prepass.rs:229-237 — replaced by integer codes + sticky error string)                */
/* ------------------------------------------------------------------------------------------ */
typedef int32_t bt_status;
enum {
    BT_OK               = 0,
    BT_ERR_INVALID_ARG  = 1,
    BT_ERR_CAPACITY     = 2, /* tile list / frontier overflow (UNCHECKED in the reference)   */
    BT_ERR_CUDA         = 3,
    BT_ERR_NCCL         = 4,
    BT_ERR_IO           = 5,
    BT_ERR_OUT_OF_MEMORY = 6,
    BT_ERR_NOT_FOUND    = 7
};

/* ------------------------------------------------------------------------------------------ */
/* POD mirrors of the reference's GPU-visible structs                                          */
/* ------------------------------------------------------------------------------------------ */

/* TileCoordinate — src/math/coordinate.rs:156-166, src/shaders/types.wgsl:26-30. 16 B.
 * lod 0 = root (1 tile per side), count(lod) = 1 << lod.  INVALID = all 0xFFFFFFFF.          */
typedef struct bt_tile { uint32_t side, lod, x, y; } bt_tile;

/* Indirect draw record — src/render/terrain_view_bind_group.rs:66-71, prepare_prepass.wgsl:38-43 */
typedef struct bt_indirect {
    uint32_t vertex_count, instance_count, base_vertex, base_instance;
} bt_indirect;

/* TerrainViewConfigUniform — terrain_view_bind_group.rs:82-116, types.wgsl:10-24. 52 B.      */
typedef struct bt_view_config_uniform {
    uint32_t tree_size;
    uint32_t geometry_tile_count;
    uint32_t refinement_count;
    float    grid_size;
    uint32_t vertices_per_row;
    uint32_t vertices_per_tile;
    float    morph_distance;
    float    blend_distance;
    float    load_distance;
    float    subdivision_distance;
    float    morph_range;
    float    blend_range;
    float    precision_threshold_distance;
} bt_view_config_uniform;

/* SideParameter — src/math/terrain_model.rs:228-249, types.wgsl:78-87. 112 B (std140).       */
typedef struct bt_side_parameter {
    int32_t view_xy[2];           /* Rust: origin_xy */
    float   view_uv[2];           /* Rust: origin_uv */
    float   c[3];    float _p0;
    float   c_s[3];  float _p1;
    float   c_t[3];  float _p2;
    float   c_ss[3]; float _p3;
    float   c_st[3]; float _p4;
    float   c_tt[3]; float _p5;
} bt_side_parameter;

/* TerrainModelApproximation — terrain_model.rs:252-259, types.wgsl:89-93. 688 B.             */
typedef struct bt_model_approximation {
    uint32_t origin_lod;
    float    approximate_height;
    uint32_t _pad[2];
    bt_side_parameter sides[6];
} bt_model_approximation;

/* CullingUniform — src/render/culling_bind_group.rs:41-55, types.wgsl:99-103. 160 B.
 * Only world_position is read by the reference shaders (frustum culling is not implemented at
 * this commit); `planes` is consumed only by the opt-in BT_FLAG_FRUSTUM_CULL extension.      */
typedef struct bt_culling_uniform {
    float world_position[3]; float _p0;
    float view_proj[16];           /* column-major; holds inverse(world_from_view) in the ref */
    float planes[5][4];
} bt_culling_uniform;

/* TerrainConfigUniform — src/render/terrain_bind_group.rs:77-93, types.wgsl:3-8. 16 B.       */
typedef struct bt_terrain_config_uniform {
    uint32_t lod_count;
    float    min_height, max_height, scale;
} bt_terrain_config_uniform;

/* AttachmentConfig (GPU) — terrain_bind_group.rs:47-72, types.wgsl:71-76. 16 B each, x8.     */
typedef struct bt_attachment_uniform { float size, scale, offset; uint32_t _pad; } bt_attachment_uniform;

/* TileTreeEntry — src/terrain_data/tile_tree.rs:51-65, types.wgsl:55-58. 8 B.
 * table index = ((side*lod_count + lod)*tree_size + x%tree_size)*tree_size + y%tree_size     */
typedef struct bt_tile_tree_entry { uint32_t atlas_index, atlas_lod; } bt_tile_tree_entry;

/* TileLookup — src/terrain_data/tile_tree.rs:67-83 (CPU-side sampling); INVALID = {0xFFFFFFFF, 0xFFFFFFFF, (0,0)} */
typedef struct bt_tile_lookup { uint32_t atlas_index, atlas_lod; float atlas_uv[2]; } bt_tile_lookup;

/* Mesh transform consumed by position_local_to_world / normal_local_to_world
 * (functions.wgsl:14-29; bevy_pbr MeshUniform built at terrain_bind_group.rs:185-205).
 * world_from_local: column-major 3x4 affine (cols x,y,z,translation);
 * normal_matrix   : column-major inverse-transpose of the upper 3x3.                          */
typedef struct bt_mesh_transform {
    float world_from_local[12];
    float normal_matrix[9];
    float _pad[3];
} bt_mesh_transform;

/* One generated vertex = the fields of VertexOutput (src/shaders/render/vertex.wgsl:14-20)
 * minus clip_position.  40 B, interleaved.  With BT_FLAG_NORMALS `normal` is the height-derived
 * surface normal (sample_normal, attachments.wgsl:51-107, blended as fragment.wgsl:99-107),
 * otherwise the model normal the reference's vertex stage outputs.                            */
typedef struct bt_vertex {
    float    position[4];   /* world_position, w = 1 */
    float    normal[3];
    float    uv[2];         /* coordinate_uv (after morph) */
    uint32_t tile_index;
} bt_vertex;

/* AttachmentFormat ids — src/terrain_data/mod.rs:50-57 (Rgb8 = 5 and Rg16 = 3 are unimplemented
 * in the reference's shaders, preprocessing.wgsl:3-5, and rejected here).                     */
enum { BT_FORMAT_RGBA8 = 0, BT_FORMAT_R16 = 1 };

/* TerrainKind — src/math/terrain_model.rs:23-36 */
enum { BT_KIND_PLANAR = 0, BT_KIND_SPHERICAL = 1,
       BT_KIND_ELLIPSOIDAL = 2 /* host-side model only (bt_terrain_model): terrains, views and preprocessors are SPHERICAL */ };

/* shader-def flags (src/render/terrain_material.rs:73-228, tiling_prepass.rs:28-79) */
enum {
    BT_FLAG_MORPH          = 1u << 1,
    BT_FLAG_BLEND          = 1u << 2,
    BT_FLAG_HIGH_PRECISION = 1u << 3,
    BT_FLAG_NORMALS        = 1u << 4,  /* fuse sample_normal into the vertex record            */
    BT_FLAG_FRUSTUM_CULL   = 1u << 6,  /* bt_select_tiles extension; default off = reference behaviour: tiles whose
                                          ten bounding points (corners + centre at min/max height) lie outside
                                          one of culling.planes[5] are neither refined nor emitted            */
    /* bt_select_tiles extension, default off (docs/implementation.md:69 "The frustum culling uses a 2D min-max height data
     * attachment to approximate the bounding volumes of each tile"): like BT_FLAG_FRUSTUM_CULL, but the height range of a tile's
     * bounding points comes from the best RESIDENT atlas tile that contains it — the (min, max) table bt_atlas_update_minmax keeps
     * per atlas slot (whole tile + 8 x 8 cells) — instead of the terrain-wide [min_height, max_height].                          */
    BT_FLAG_MINMAX_CULL    = 1u << 9,
    /* BASELINE configs[2] ("trilinear + aniso sampling"): the reference's SAMPLER (linear / linear / linear, anisotropy clamp 16,
     * terrain_bind_group.rs:120-126; its fragment stage samples with textureSampleGrad, attachments.wgsl:16-20) evaluated per
     * vertex with the footprint of one grid step: N = min(ceil(Pmax / Pmin), 16) taps along the major axis, lod =
     * clamp(log2(Pmax / N), 0, mips - 1), every tap mix(bilinear(floor(lod)), bilinear(floor(lod) + 1), fract(lod)) over the
     * generate_mipmaps chain (terrain_data/mod.rs:143-219).  Default off: the reference's vertex stage samples mip 0.          */
    BT_FLAG_FILTER_TRILINEAR_ANISO = 1u << 7,
    /* per-vertex albedo: the Rgba8 attachment (sRGB-decoded per texel before filtering) sampled at the vertex, blended across
     * LODs like the height, written as pack4x8unorm to the view's albedo stream (bt_view_albedo_buffer), 4 bytes per vertex   */
    BT_FLAG_ALBEDO         = 1u << 8
};

/* ------------------------------------------------------------------------------------------ */
/* configuration structs (host side)                                                           */
/* ------------------------------------------------------------------------------------------ */

/* AttachmentConfig — src/terrain_data/mod.rs:88-109 */
typedef struct bt_attachment_config {
    uint32_t texture_size;
    uint32_t border_size;
    uint32_t mip_level_count;
    uint32_t format;          /* BT_FORMAT_* */
} bt_attachment_config;

/* TerrainConfig + TerrainModel — src/terrain.rs:27-49, src/math/terrain_model.rs:40-104 */
typedef struct bt_terrain_desc {
    uint32_t kind;            /* BT_KIND_* */
    uint32_t lod_count;
    uint32_t atlas_size;      /* number of atlas layers (tile slots) */
    uint32_t attachment_count;
    bt_attachment_config attachments[8];
    float    min_height, max_height;
    double   scale;           /* TerrainModel::scale(): side_length/2 (planar) or radius       */
    bt_mesh_transform transform;
} bt_terrain_desc;

/* Per-frame inputs of one (terrain, view) pair = the Prepare-stage write_buffers of the
 * reference (terrain_view_bind_group.rs:238-262, gpu_tile_tree.rs:84-95, culling_bind_group.rs:87-101) */
typedef struct bt_view_frame {
    bt_view_config_uniform view_config;
    bt_model_approximation approximation;
    bt_culling_uniform     culling;
    float                  view_world_position[3]; /* bevy `view.world_position` (vertex stage) */
    uint32_t               tile_tree_entry_count;  /* sides*lod_count*tree_size^2               */
    const bt_tile_tree_entry* tile_tree;           /* host pointer, copied                      */
    const uint32_t*        origins;                /* host pointer, sides*lod_count uvec2, may be NULL */
} bt_view_frame;

typedef struct bt_vertex_buffer {
    bt_vertex* vertices;      /* DEVICE pointer owned by the view, valid until next gen/destroy */
    uint64_t   capacity;      /* vertices the buffer can hold                                   */
} bt_vertex_buffer;

typedef struct bt_context bt_context;
typedef struct bt_terrain bt_terrain;
typedef struct bt_view    bt_view;
typedef struct bt_preprocessor bt_preprocessor;

/* ------------------------------------------------------------------------------------------ */
/* context                                                                                     */
/* ------------------------------------------------------------------------------------------ */
uint32_t    bt_abi_version(void);
bt_status   bt_context_create(int32_t device, bt_context** out);
void        bt_context_destroy(bt_context* ctx);
const char* bt_last_error(const bt_context* ctx);   /* sticky, per context */
bt_status   bt_context_synchronize(bt_context* ctx);

/* ------------------------------------------------------------------------------------------ */
/* terrain = TileAtlas GPU mirror (src/terrain_data/gpu_tile_atlas.rs:208-223):                 */
/* per attachment one pitched array [atlas_size][texture_size][texture_size] (mip 0)            */
/* ------------------------------------------------------------------------------------------ */
bt_status bt_terrain_create(bt_context* ctx, const bt_terrain_desc* desc, bt_terrain** out);
void      bt_terrain_destroy(bt_terrain* terrain);
/* = GpuTileAtlas::upload_tiles (gpu_tile_atlas.rs:309-336): host texels (mip 0, row-major) -> slot.
 * The copy is issued asynchronously on `stream` straight from `texels`: with PINNED host memory the buffer must stay valid
 * and unmodified until the stream has reached the copy (pageable memory is staged by the driver before the call returns). */
bt_status bt_atlas_upload_tile(bt_terrain* terrain, uint32_t attachment, uint32_t atlas_index,
                               const void* texels, size_t bytes, void* stream);
/* device-to-device variant (tile produced by bt_preprocessor stays on the GPU) */
bt_status bt_atlas_copy_tile(bt_terrain* terrain, uint32_t attachment, uint32_t atlas_index,
                             const void* device_texels, size_t bytes, void* stream);
bt_status bt_atlas_download_tile(bt_terrain* terrain, uint32_t attachment, uint32_t atlas_index,
                                 void* texels, size_t bytes);
/* Tile streaming from disk = AtlasTileAttachmentWithData::start_loading (tile_atlas.rs:118-149) with STORE_PNG = false:
 * reads <directory>/<side>_<lod>_<x>_<y>.bin (raw mip-0 texels, as written by bt_preprocessor_save / the reference's
 * Save tasks) through a pinned staging buffer into the atlas slot.  Returns BT_ERR_IO if the file is missing or short. */
bt_status bt_atlas_load_tile_file(bt_terrain* terrain, uint32_t attachment, uint32_t atlas_index, const char* directory,
                                  const bt_tile* tile, void* stream);
/* TC::load_file / decode_alloc (formats/mod.rs:14-31): bincode-standard Vec<TileCoordinate>; *count receives the total */
bt_status bt_tile_config_load(const char* path, bt_tile* tiles, uint32_t capacity, uint32_t* count);
/* AtlasAttachment::sample + AttachmentData::sample (tile_atlas.rs:249-258, terrain_data/mod.rs:221-264): bilinear sample of
 * one atlas tile at lookup.atlas_uv with the twin's exact f32 arithmetic; fetches the four texels from the device (blocking).
 * An INVALID lookup yields zeros, as in the reference.                                                                   */
bt_status bt_atlas_sample(bt_terrain* terrain, uint32_t attachment, const bt_tile_lookup* lookup, float out[4]);
/* Mip levels 1.. of one atlas slot from its mip 0 on the GPU = AttachmentData::generate_mipmaps, which the reference runs on
 * the CPU at tile load (tile_atlas.rs:141, terrain_data/mod.rs:143-219): R16 = mean of the non-zero texels (integer division),
 * Rgba8 = sum / 4.  Call after uploading / copying a tile when the attachment has mip_level_count > 1 and the filtered
 * sampler (BT_FLAG_FILTER_TRILINEAR_ANISO) is used.  The chain of every slot lives in a second array per attachment.           */
bt_status bt_atlas_generate_mips(bt_terrain* terrain, uint32_t attachment, uint32_t atlas_index, void* stream);
bt_status bt_atlas_mips_device_ptr(bt_terrain* terrain, uint32_t attachment, void** out, size_t* chain_bytes_per_slot);
/* (min, max) table of the height tile in one atlas slot for BT_FLAG_MINMAX_CULL: 65 pairs of uint16 (the tile's centre region, then an
 * 8 x 8 grid of cells over it; every region grown by one texel, zeros = no-data ignored).  Call after uploading / copying a height tile. */
bt_status bt_atlas_update_minmax(bt_terrain* terrain, uint32_t atlas_index, void* stream);
bt_status bt_atlas_minmax_device_ptr(bt_terrain* terrain, void** out);   /* uint16 [atlas_size][65][2] */
/* device base pointer of an attachment's atlas array (for zero-copy interop / tests) */
bt_status bt_atlas_device_ptr(bt_terrain* terrain, uint32_t attachment, void** out);

/* ------------------------------------------------------------------------------------------ */
/* view = TerrainViewData (terrain_view_bind_group.rs:118-187): final/temporary tile buffers,   */
/* indirect record, uniforms, tile-tree table, vertex buffer                                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct bt_view_desc {
    uint32_t geometry_tile_count;  /* capacity of final_tiles / temporary_tiles (ref default 1e6) */
    uint32_t tree_size;
    uint32_t grid_size;
    uint32_t readback_tile_capacity; /* tiles bt_view_run_frame can return per call (pinned host buffer, allocated by
                                        bt_view_create); 0 = min(geometry_tile_count, 65536)                        */
    uint64_t max_vertices;         /* vertex buffer capacity in vertices; 0 = grow on demand in bt_gen_vertices */
} bt_view_desc;

bt_status bt_view_create(bt_terrain* terrain, const bt_view_desc* desc, bt_view** out);
void      bt_view_destroy(bt_view* view);
/* Prepare-stage uploads; host structs may live in pageable or pinned memory */
bt_status bt_view_update(bt_view* view, const bt_view_frame* frame, void* stream);

/* Replaces TilingPrepassNode::run (src/render/tiling_prepass.rs:204-271): prepare_root,
 * (refine_tiles, prepare_next) x refinement_count, refine_tiles, prepare_render — ONE launch.
 * Output order is the order a sequential execution of the reference shaders would produce.    */
bt_status bt_select_tiles(bt_view* view, uint32_t flags, void* stream);

/* The views of one frame in ONE launch (gridDim.x = views): TilingPrepassNode::run loops over (terrain, view) pairs on the
 * host (tiling_prepass.rs:224-263, 63 dispatches each); here every view gets its own CTA of the same kernel, so N views cost
 * the latency of one.  The views must share a context and a terrain kind; results are exactly those of N bt_select_tiles calls. */
typedef struct bt_view_batch bt_view_batch;
bt_status bt_view_batch_create(bt_view* const* views, uint32_t count, bt_view_batch** out);
void      bt_view_batch_destroy(bt_view_batch* batch);      /* before the views it references */
bt_status bt_select_tiles_batch(bt_view_batch* batch, uint32_t flags, void* stream);

/* Replaces the vertex stage of the indirect draw (DrawTerrainCommand,
 * terrain_view_bind_group.rs:274-298; vertex.wgsl:85-98) by writing the strip-ordered vertex
 * stream (indirect.vertex_count records) to a device buffer.                                  */
bt_status bt_gen_vertices(bt_view* view, uint32_t flags, void* stream, bt_vertex_buffer* out);

/* One frame in one call = bt_view_update + bt_select_tiles + bt_gen_vertices + read-back of the tile list and indirect
 * record (Prepare + TilingPrepassNode::run + the draw's vertex stage of the reference's frame).  The read-back overlaps
 * gen_vertices on an internal copy stream; on return the tile list is on the host, the vertex buffer is (being) written
 * on the device, stream-ordered on `stream`.  Needs a view created with max_vertices > 0.  Allocates nothing: the copy
 * stream, events and pinned buffers belong to the view (bt_view_create); at most readback_tile_capacity tiles are
 * returned (*count still receives the true number).  The 1.5 GB vertex stream is NOT copied back: like the reference's
 * draw, its consumer reads it on the device.  The frame's inputs (uniform block, tile-tree table) are written to the view's
 * pinned staging blocks and FETCHED by the select kernel itself over PCIe (no copy operation in front of the kernels).      */
bt_status bt_view_run_frame(bt_view* view, const bt_view_frame* frame, uint32_t select_flags, uint32_t gen_flags, void* stream,
                        bt_tile* tiles, uint32_t capacity, uint32_t* count, bt_indirect* indirect, bt_vertex_buffer* vertices);
/* blocking D2H of final_tiles + indirect (the reference never reads these back; the draw
 * consumes them on-device)                                                                    */
bt_status bt_view_readback(bt_view* view, bt_tile* tiles, uint32_t capacity, uint32_t* count,
                           bt_indirect* indirect, void* stream);
/* status word written by the last bt_select_tiles (BT_OK / BT_ERR_CAPACITY), plus the number
 * of parents that still wanted subdivision after the last pass (dropped by the reference, Q2) */
bt_status bt_view_select_status(bt_view* view, uint32_t* dropped_parents, void* stream);
bt_status bt_view_download_vertices(bt_view* view, bt_vertex* out, uint64_t first, uint64_t count,
                                    void* stream);
/* BT_FLAG_ALBEDO: the per-vertex colour stream written beside the vertex stream by the last bt_gen_vertices (strip order, one
 * pack4x8unorm word per vertex; allocated on the first ALBEDO frame, with the vertex buffer's capacity) */
bt_status bt_view_albedo_buffer(bt_view* view, uint32_t** device_albedo, uint64_t* capacity);
bt_status bt_view_download_albedo(bt_view* view, uint32_t* out, uint64_t first, uint64_t count, void* stream);
/* device pointers (final_tiles, indirect) for zero-copy consumers */
bt_status bt_view_device_ptrs(bt_view* view, bt_tile** final_tiles, bt_indirect** indirect);

/* ------------------------------------------------------------------------------------------ */
/* host-side f64 producers of the per-frame inputs (no GPU needed)                              */
/* ------------------------------------------------------------------------------------------ */

/* TerrainModel — src/math/terrain_model.rs:40-128 (planar / sphere / ellipsoid) */
typedef struct bt_terrain_model {
    uint32_t kind;             /* BT_KIND_* */
    uint32_t _pad;
    double   position[3];
    double   scale[3];         /* side_length (planar) or radius (sphere), splatted; (major, minor, major) for an ellipsoid */
    float    min_height, max_height;
} bt_terrain_model;

/* TerrainViewConfig — src/terrain_view.rs:19-64 */
typedef struct bt_terrain_view_config {
    uint32_t tree_size, geometry_tile_count, refinement_count, grid_size;
    double   subdivision_tolerance, precision_threshold_distance, load_distance,
             morph_distance, blend_distance;
    float    morph_range, blend_range;
    uint32_t origin_lod;
    uint32_t _pad;
} bt_terrain_view_config;

void bt_terrain_model_planar(const double position[3], double side_length, float min_height,
                             float max_height, bt_terrain_model* out);   /* terrain_model.rs:84-93  */
void bt_terrain_model_sphere(const double position[3], double radius, float min_height,
                             float max_height, bt_terrain_model* out);   /* terrain_model.rs:95-104 */
/* TerrainModel::ellipsoid (terrain_model.rs:106-128): scale (major, minor, major), the polar (minor) axis is world y;
 * position_world_to_local projects onto the ellipsoid first (src/math/ellipsoid.rs); scale() = (major + minor) / 2     */
void bt_terrain_model_ellipsoid(const double position[3], double major_axis, double minor_axis, float min_height,
                                float max_height, bt_terrain_model* out);
void bt_terrain_view_config_default(bt_terrain_view_config* out);       /* terrain_view.rs:47-64   */
/* TerrainModel::scale (terrain_model.rs:183-193) */
double bt_terrain_model_scale(const bt_terrain_model* model);
/* MeshUniform equivalent of TerrainModel::transform (terrain_model.rs:196-203) */
void bt_terrain_model_mesh_transform(const bt_terrain_model* model, bt_mesh_transform* out);
/* TerrainViewConfigUniform::from_tile_tree o TileTree::new (terrain_view_bind_group.rs:98-115, tile_tree.rs:135-173) */
void bt_view_config_uniform_from(const bt_terrain_view_config* cfg, const bt_terrain_model* model,
                                 bt_view_config_uniform* out);
/* TerrainModelApproximation::compute (terrain_model.rs:263-360) */
bt_status bt_model_approximation_compute(const bt_terrain_model* model, const double view_world_position[3],
                                         uint32_t origin_lod, float approximate_height,
                                         bt_model_approximation* out);

/* Clipmap residency = TileTree (src/terrain_data/tile_tree.rs:94-386) + the request/release/
 * best-tile bookkeeping of TileAtlasState (src/terrain_data/tile_atlas.rs:369-503), without the
 * async file I/O: tiles the tree requests are reported to the caller, who uploads them and
 * marks them loaded.                                                                          */
typedef struct bt_tile_tree bt_tile_tree;
/* TileAtlasState (tile_atlas.rs:272-300) is per TERRAIN in the reference and shared by the TileTree of every view:
 * slot allocation, the LRU of unused slots, existing tiles, the load queue.  Views of one terrain must share one state
 * (bt_tile_tree_create_shared), otherwise each tree would hand out atlas indices 0..atlas_size-1 on its own and the views
 * would overwrite each other's slots.  bt_tile_tree_create makes a private state for single-view terrains.           */
typedef struct bt_atlas_state bt_atlas_state;
bt_status bt_atlas_state_create(uint32_t atlas_size, bt_atlas_state** out);
void      bt_atlas_state_destroy(bt_atlas_state* state);     /* freed once the last tree that references it is gone */
bt_status bt_tile_tree_create_shared(const bt_terrain_model* model, const bt_terrain_view_config* cfg,
                                     uint32_t lod_count, bt_atlas_state* state, bt_tile_tree** out);
bt_status bt_tile_tree_create(const bt_terrain_model* model, const bt_terrain_view_config* cfg,
                              uint32_t lod_count, uint32_t atlas_size, bt_tile_tree** out);
void      bt_tile_tree_destroy(bt_tile_tree* tree);
/* register a tile as existing on disk / in the pyramid (TileAtlasState::existing_tiles) */
bt_status bt_tile_tree_add_existing(bt_tile_tree* tree, const bt_tile* tiles, uint32_t count);
/* TileTree::update (tile_tree.rs:268-333) + TileAtlasState::request_tile/release_tile.
 * Writes up to `cap` (tile, atlas_index) load requests; *count receives the total.            */
typedef struct bt_tile_load { bt_tile tile; uint32_t atlas_index; uint32_t _pad[3]; } bt_tile_load;
bt_status bt_tile_tree_update(bt_tile_tree* tree, const double view_world_position[3],
                              bt_tile_load* loads, uint32_t cap, uint32_t* count);
/* TileAtlasState::loaded_tile_attachment for all attachments of a tile */
bt_status bt_tile_tree_mark_loaded(bt_tile_tree* tree, const bt_tile* tile);
/* TileTree::adjust_to_tile_atlas (tile_tree.rs:363-374) -> entries + origins for bt_view_frame */
bt_status bt_tile_tree_entries(bt_tile_tree* tree, bt_tile_tree_entry* entries, uint32_t entry_cap,
                               uint32_t* origins_xy, uint32_t origin_cap);

/* ---- the same clipmap on the DEVICE --------------------------------------------------------------------
 * bt_clipmap is the GPU twin of the per-frame half of bt_tile_tree for ONE view: the slots (TileState per side x lod x
 * tree_size^2), the origins and the best-tile table live in device memory, the table directly in the view's buffer that
 * select_tiles / gen_vertices read.  A frame costs one 128-byte parameter upload, one single-CTA kernel (origins, f64 tile
 * distances, the request / release state machine of tile_tree.rs:268-333, lists compacted in the reference's loop order)
 * and the list read-back; TileAtlasState (slot allocation, LRU, load queue) stays on the host because only the host can
 * load tiles.  Results are those of bt_tile_tree_update / bt_tile_tree_entries on the same inputs (tests/test_gpu_clipmap.py).
 * Create the clipmap on a tree that has not been updated yet; from then on drive the tree through the clipmap only.
 * With a clipmap attached, bt_view_update accepts frame->tile_tree == NULL (no per-frame table upload).                    */
typedef struct bt_clipmap bt_clipmap;
bt_status bt_clipmap_create(bt_view* view, bt_tile_tree* tree, bt_clipmap** out);
void      bt_clipmap_destroy(bt_clipmap* clipmap);     /* releases the view's outstanding requests, like bt_tile_tree_destroy;
                                                          destroy the clipmap BEFORE its view and its tree                      */
/* twin of bt_tile_tree_update (same outputs, same order of the load requests) */
bt_status bt_clipmap_update(bt_clipmap* clipmap, const double view_world_position[3],
                            bt_tile_load* loads, uint32_t cap, uint32_t* count, void* stream);
/* twin of bt_tile_tree_entries: recomputes the view's device table (call after bt_tile_tree_mark_loaded of the new tiles);
 * the set of loaded tiles is re-uploaded only when it changed                                                              */
bt_status bt_clipmap_entries(bt_clipmap* clipmap, void* stream);
/* read-back of the device table + origins (tests, debugging) */
bt_status bt_clipmap_download(bt_clipmap* clipmap, bt_tile_tree_entry* entries, uint32_t entry_cap,
                              uint32_t* origins_xy, uint32_t origin_cap, void* stream);
/* bytes the last bt_clipmap_update / bt_clipmap_entries call moved host->device and device->host */
void      bt_clipmap_transfer_bytes(const bt_clipmap* clipmap, uint64_t* h2d, uint64_t* d2h);
void      bt_tile_tree_set_approximate_height(bt_tile_tree* tree, float h);
/* CPU-side sampling chain of TileTree::approximate_height (tile_tree.rs:376-386 -> sample_attachment,
 * terrain_data/mod.rs:267-295): surface position, TileTree::compute_blend (tile_tree.rs:223-237) and TileTree::lookup_tile
 * (:239-266) for the blend LOD and, when blend_ratio > 0, the LOD below it.  The texels live on the device: feed the lookups
 * to bt_atlas_sample and mix the two results with blend_ratio (sample_attachment does Vec4::lerp).                      */
bt_status bt_tile_tree_sample_lookup(bt_tile_tree* tree, const double sample_world_position[3], bt_tile_lookup lookups[2],
                                     uint32_t* lookup_count, float* blend_ratio);

/* ------------------------------------------------------------------------------------------ */
/* preprocessor = Preprocessor + TerrainPreprocessNode (src/preprocess/preprocessor.rs:214-343,  */
/* src/preprocess/mod.rs:143-217) with the pixel kernels of src/shaders/preprocess/{split,downsample,stitch}.wgsl */
/* ------------------------------------------------------------------------------------------ */

/* PreprocessDataset — preprocessor.rs:36-55.  The source image is given as pixels instead of a
 * path (decoders are out of scope): `pixels` is a HOST pointer unless `pixels_on_device`.      */
typedef struct bt_dataset {
    uint32_t attachment_index;
    uint32_t side;
    float    top_left[2];
    float    bottom_right[2];
    uint32_t lod_begin, lod_end;     /* lod_range = lod_begin..lod_end (exclusive) */
    uint32_t width, height;          /* source image size in pixels */
    uint32_t pixels_on_device;
    /* row window (sharded hosts hold only their band of a huge source): `pixels` points at source
     * row `row_begin` and holds `row_count` rows; row_count == 0 means the whole image.          */
    uint32_t row_begin, row_count;
    uint32_t _pad;
    const void* pixels;              /* R16: uint16, Rgba8: 4 x uint8; row-major */
} bt_dataset;

/* Sharding of the deepest-LOD tile rows over ranks (new; the reference is single-GPU).
 * world = 1 reproduces the reference schedule.                                               */
typedef struct bt_shard { uint32_t rank, world; } bt_shard;

bt_status bt_preprocessor_create(bt_context* ctx, uint32_t kind /*BT_KIND_*/,
                                 const bt_attachment_config* attachments, uint32_t attachment_count,
                                 uint32_t lod_count, bt_shard shard, bt_preprocessor** out);
void      bt_preprocessor_destroy(bt_preprocessor* p);
/* count == 1: Preprocessor::preprocess_tile (preprocessor.rs:298-312) — records the dataset and its
 * task list; successive calls for one attachment overlay in call order (split.wgsl:36-42
 * fall-through).  count > 1 (e.g. the 6 cube faces): Preprocessor::preprocess_spherical ordering
 * (:314-343) — split/downsample ALL datasets, barrier, then per LOD per dataset stitch, so that
 * cross-face neighbours exist when the stitch tasks resolve them.
 * Sharded preprocessors (shard.world > 1) take full-coverage datasets: one per job (planar) or the six cube
 * faces of one attachment with a common lod_range (spherical).                                */
bt_status bt_preprocessor_add_datasets(bt_preprocessor* p, const bt_dataset* datasets, uint32_t count);
/* optional: allocate the device pyramid store of all pending jobs now (keeps cudaMalloc out of run_*) */
bt_status bt_preprocessor_reserve(bt_preprocessor* p, void* stream);
/* Runs split -> downsample (per level) -> [halo exchange hook] -> stitch on `stream`.
 * Phases are exposed separately so that a multi-GPU host can run the collective between them. */
bt_status bt_preprocessor_run_split_downsample(bt_preprocessor* p, void* stream);
bt_status bt_preprocessor_run_stitch(bt_preprocessor* p, void* stream);
bt_status bt_preprocessor_run(bt_preprocessor* p, void* stream);          /* both, world == 1 */
/* Host-side plan of the pending job's exchange (no GPU needed: `ctx` may be NULL at creation for a
 * plan-only preprocessor, which is how the world_size-2 gloo tests exercise this logic on CPU).
 * One descriptor = `pixels` contiguous texels starting at (tile_row, tile_col) of `tile`
 * <-> `buffer_offset` bytes into the send buffer (pack) or the gathered [world][bytes] buffer (unpack). */
typedef struct bt_exchange_copy {
    bt_tile  tile;
    uint32_t tile_row, tile_col, pixels;
    uint32_t src_rank;             /* rank whose data this segment carries */
    uint64_t buffer_offset;
} bt_exchange_copy;
bt_status bt_preprocessor_plan(bt_preprocessor* p);   /* allocate the pending job's tile table (implicit in run_*) */
/* every slot incl. ghosts/replicas; kinds: 0 = owned output, 1 = ghost (halo receiver), 2 = replica */
bt_status bt_preprocessor_plan_tiles(bt_preprocessor* p, uint32_t attachment, bt_tile* tiles, uint32_t* kinds, uint32_t cap, uint32_t* count);
/* kind: 0 = halo strips, 1 = sub-root tiles; unpack: 0 = pack plan, 1 = unpack plan */
bt_status bt_preprocessor_plan_exchange(bt_preprocessor* p, uint32_t kind, uint32_t unpack, bt_exchange_copy* out, uint32_t cap, uint32_t* count);
/* multi-GPU hooks: sub-root gather + halo strips (see DESIGN.md "multi-GPU") */
bt_status bt_preprocessor_halo_bytes(bt_preprocessor* p, uint64_t* send_bytes_per_rank);
bt_status bt_preprocessor_halo_pack(bt_preprocessor* p, void* device_send, void* stream);
bt_status bt_preprocessor_halo_unpack(bt_preprocessor* p, const void* device_gathered, void* stream);
bt_status bt_preprocessor_roots_bytes(bt_preprocessor* p, uint64_t* send_bytes_per_rank);
bt_status bt_preprocessor_roots_pack(bt_preprocessor* p, void* device_send, void* stream);
bt_status bt_preprocessor_roots_unpack_and_finish(bt_preprocessor* p, const void* device_gathered, void* stream);
/* ---- multi-GPU communicator (one process per GPU of one box; NCCL over NVLink / NVSwitch, loaded with dlopen at
 * bt_comm_create so single-GPU hosts never touch it).  Bootstrap like any NCCL program: rank 0 calls bt_comm_unique_id and
 * the host distributes the 128 bytes (MPI, a TCP store, a file ...), then every rank calls bt_comm_create (collective).     */
#define BT_COMM_UNIQUE_ID_BYTES 128u
#define BT_COMM_MAX_RANKS 8u
typedef struct bt_comm bt_comm;
bt_status   bt_comm_unique_id(void* out, size_t bytes);                     /* = ncclGetUniqueId */
bt_status   bt_comm_create(bt_context* ctx, uint32_t rank, uint32_t world, const void* unique_id, size_t id_bytes, bt_comm** out);
void        bt_comm_destroy(bt_comm* comm);
bt_status   bt_comm_info(const bt_comm* comm, uint32_t* rank, uint32_t* world, int32_t* nccl_version, uint32_t* peer_access);
bt_status   bt_comm_barrier(bt_comm* comm, void* stream);                   /* stream-ordered, all ranks */
bt_status   bt_comm_all_reduce_sum_u64(bt_comm* comm, uint64_t* values, uint32_t count, void* stream);   /* host in/out, blocking */
const char* bt_comm_load_error(void);                                       /* why libnccl could not be loaded, if it could not */
/* The sharded preprocess in ONE call (replaces the Python-driven phase sequence): per wave of pending jobs
 * split + downsample of this rank's band -> pack -> ONE grouped all-gather of all sub-root tiles and border strips on the
 * communicator's stream, overlapped with the stitch of every tile whose neighbours are local -> unpack -> the replicated top
 * levels -> stitch of the boundary tiles (stitch fused behind its collective).  Stream-ordered on `stream`;
 * bt_preprocessor_reserve beforehand keeps every allocation and table upload out of it.  Returns BT_ERR_NCCL on transport errors. */
bt_status   bt_preprocessor_run_sharded(bt_preprocessor* p, bt_comm* comm, void* stream);
/* results */
bt_status bt_preprocessor_tile_count(bt_preprocessor* p, uint32_t attachment, uint32_t* count);
bt_status bt_preprocessor_list_tiles(bt_preprocessor* p, uint32_t attachment, bt_tile* tiles, uint32_t cap, uint32_t* count);
/* device pointer of one tile's texels (texture_size^2 pixels, mip 0) */
bt_status bt_preprocessor_tile_device_ptr(bt_preprocessor* p, uint32_t attachment, const bt_tile* tile, void** out);
bt_status bt_preprocessor_get_tile(bt_preprocessor* p, uint32_t attachment, const bt_tile* tile, void* host_out, size_t bytes);
/* order-independent 64-bit content hash of this rank's output tiles of one attachment (coordinate-keyed, computed on the
 * device): the wrapping sum over the ranks of a sharded run equals the hash of the single-GPU pyramid iff all tiles agree */
bt_status bt_preprocessor_hash(bt_preprocessor* p, uint32_t attachment, uint64_t* out, void* stream);
/* Save (tile_atlas.rs:77-116, formats/mod.rs:9-36): <dir>/<side>_<lod>_<x>_<y>.bin + config.tc */
bt_status bt_preprocessor_save(bt_preprocessor* p, uint32_t attachment, const char* directory);

/* ------------------------------------------------------------------------------------------ */
/* utilities                                                                                   */
/* ------------------------------------------------------------------------------------------ */
/* planes() of src/render/culling_bind_group.rs:25-38 from a column-major clip_from_world matrix (never called by the
 * reference at this commit; feeds bt_culling_uniform.planes for BT_FLAG_FRUSTUM_CULL) */
void bt_culling_planes(const float clip_from_world[16], float planes[5][4]);
/* CPU mip chain at tile load (AttachmentData::generate_mipmaps, terrain_data/mod.rs:143-219)
 * is replaced by a device kernel operating on an uploaded tile; out is a device pointer pair.  */
bt_status bt_generate_mipmaps(bt_context* ctx, uint32_t format, uint32_t texture_size, uint32_t mip_level_count,
                              const void* device_mip0, void* device_mips_out, void* stream);
/* deterministic integer hash-noise source generator used by bench.py/tests (not part of the
 * reference; see DESIGN.md "synthetic inputs")                                                */
bt_status bt_synth_source(bt_context* ctx, uint32_t format, uint32_t width, uint32_t height,
                          uint32_t x0, uint32_t y0, uint32_t seed, uint32_t holes, void* device_out, void* stream);
/* Source-image decoders for preprocessor inputs (host only; PreprocessDataset::path in the reference is an asset path that
 * Bevy's ImageLoader / src/formats/tiff.rs decode): PNG 16-bit gray -> BT_FORMAT_R16, PNG 8-bit RGB / RGBA -> BT_FORMAT_RGBA8,
 * TIFF 16-bit gray (strips or tiles; none / LZW / Deflate; predictor 1 or 2; II or MM) -> R16 with the sample bytes kept as they
 * are (tiff.rs casts every integer result to R16Unorm texels), TIFF 8-bit RGB / RGBA -> Rgba8.  `pixels` is malloc'ed; release
 * it with bt_image_free.  On failure returns BT_ERR_IO and writes a message into `error` (may be NULL).                      */
typedef struct bt_image { uint32_t width, height, format /* BT_FORMAT_* */, _pad; void* pixels; } bt_image;
bt_status bt_image_load(const char* path, bt_image* out, char* error, size_t error_capacity);
void      bt_image_free(bt_image* image);
/* numerics self-test (all counters must be 0): mismatches[0]/[1] = 16-/8-bit unorm values whose fast exact decode
 * (scalar and packed-f32 forms) differs from the IEEE division it replaces, plus ([1]) values whose round-down pack
 * differs from floor(0.5 + M*clamp(v,0,1)); mismatches[2] = operand pairs whose branch-free exact division differs from
 * `/`, or whose packed-f32 product / difference / quotient differs from the scalar IEEE operations (a contracted
 * multiply-add would show up here); mismatches[3] = inputs whose branch-free square root / packed length differs from
 * sqrtf.  log2_out[i] = the kernels' deterministic log2 of x[i] (host arrays), compared with the oracle by
 * tests/test_gpu_numerics.py                                                                      */
bt_status bt_selftest_numerics(bt_context* ctx, const float* x, float* log2_out, uint32_t n, uint32_t mismatches[4]);
/* raw device memory helpers for hosts without a CUDA binding (ctypes tests) */
bt_status bt_device_alloc(bt_context* ctx, size_t bytes, void** out);
bt_status bt_device_free(bt_context* ctx, void* ptr);
bt_status bt_memcpy_h2d(bt_context* ctx, void* dst_device, const void* src_host, size_t bytes, void* stream);
bt_status bt_memcpy_d2h(bt_context* ctx, void* dst_host, const void* src_device, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BEVY_TERRAIN_B200_H */

/*
 * bt_oracle.h — CPU oracle for the bevy_terrain hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (bevy_terrain_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference (kurtkuehnert/bevy_terrain @ 5a88daf) ships no tests, golden
 * vectors or fixtures, and cannot be compiled here (Rust + WGSL, no cargo/rustc/naga).  The
 * oracle is therefore a line-by-line restatement of the cited WGSL / Rust, cross-checked
 * against the reference's own independent CPU twins where they exist (see tests/test_oracle_*).
 */
#ifndef BT_ORACLE_H
#define BT_ORACLE_H

#include "../include/bevy_terrain_b200.h" /* POD struct layouts only */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- numerics primitives (exposed for unit tests) ---- */
float    bto_log2f(float x);
uint32_t bto_f32_to_u32(float x);
void     bto_coordinate_change_lod(uint32_t* lod, uint32_t xy[2], float uv[2], uint32_t new_lod);
void     bto_local_position(int spherical, uint32_t side, uint32_t lod, const uint32_t xy[2],
                            const float uv[2], float out[3]);

/* ---- per-frame: select ---- */
typedef struct bto_select_result {
    uint32_t final_count;
    uint32_t dropped_parents;   /* parents that wanted subdivision in the last pass (Q2) */
    uint32_t max_frontier;
    uint32_t overflow;          /* 1 if capacity was exceeded (reference: undefined behaviour) */
} bto_select_result;

void bto_select_tiles(const bt_view_config_uniform* vc, const bt_model_approximation* approx,
                      const bt_culling_uniform* culling, const bt_mesh_transform* mesh,
                      int spherical, bt_tile* final_tiles, uint32_t capacity,
                      bt_indirect* indirect, bto_select_result* result);

/* with the opt-in frustum-cull extension (BT_FLAG_FRUSTUM_CULL) */
void bto_select_tiles_ex(const bt_view_config_uniform* vc, const bt_model_approximation* approx,
                         const bt_culling_uniform* culling, const bt_mesh_transform* mesh,
                         int spherical, int frustum_cull, float min_height, float max_height,
                         bt_tile* final_tiles, uint32_t capacity, bt_indirect* indirect, bto_select_result* result);

/* min / max-height culling (BT_FLAG_MINMAX_CULL): per-atlas-tile bounds table (65 (min, max) pairs: whole tile + 8 x 8 cells) and the
 * select variant that bounds every tile with it (tile_tree / minmax_table NULL = terrain-wide bounds = bto_select_tiles_ex) */
void bto_tile_minmax(const uint16_t* tile, uint32_t T, uint32_t B, uint16_t out[130]);
void bto_select_tiles_minmax(const bt_view_config_uniform* vc, const bt_model_approximation* approx,
                             const bt_culling_uniform* culling, const bt_mesh_transform* mesh,
                             int spherical, int frustum_cull, float min_height, float max_height,
                             const bt_tile_tree_entry* tile_tree, const uint16_t* minmax_table, uint32_t lod_count, uint32_t atlas_size,
                             bt_tile* final_tiles, uint32_t capacity, bt_indirect* indirect, bto_select_result* result);

/* ---- per-frame: vertex generation ---- */
typedef struct bto_vertex_inputs {
    bt_view_config_uniform    view_config;
    bt_model_approximation    approximation;
    bt_terrain_config_uniform terrain_config;
    bt_attachment_uniform     attachments[8];
    bt_mesh_transform         mesh;
    float                     view_world_position[3];
    int32_t                   spherical;
    uint32_t                  texture_size;      /* of attachment 0 (height, R16) */
    uint32_t                  atlas_size;        /* layers in height_atlas; array indices clamp to atlas_size-1 (0 = unchecked) */
    const bt_tile_tree_entry* tile_tree;
    const uint16_t*           height_atlas;      /* [atlas_index][T][T] */
    const bt_tile*            tiles;
    uint32_t                  tile_count;
    uint32_t                  flags;             /* BT_FLAG_* */
} bto_vertex_inputs;

/* generates vertices [first, first+count) of the strip-ordered stream; OpenMP over vertices */
void bto_gen_vertices(const bto_vertex_inputs* in, uint64_t first, uint64_t count, bt_vertex* out,
                      int threads);

/* configs[2] sampler (BT_FLAG_FILTER_TRILINEAR_ANISO) and per-vertex albedo (BT_FLAG_ALBEDO): extra inputs beside bto_vertex_inputs.
 * Mip chains use bto_generate_mipmaps' layout (levels 1.. appended), one chain per atlas slot.                        */
typedef struct bto_filter_inputs {
    const uint16_t* height_mips;     /* [atlas][chain] or NULL when mip_level_count == 1 */
    const uint8_t*  albedo_atlas;    /* [atlas][T][T][4] mip 0 (Rgba8, sRGB-encoded rgb) */
    const uint8_t*  albedo_mips;
    uint32_t        mip_level_count;
    uint32_t        albedo_attachment;   /* index into bto_vertex_inputs.attachments for the albedo's size / scale / offset */
} bto_filter_inputs;
void bto_gen_vertices_filtered(const bto_vertex_inputs* in, const bto_filter_inputs* fi, uint64_t first, uint64_t count,
                               bt_vertex* out, uint32_t* albedo_out /* may be NULL */, int threads);

/* ---- preprocess ---- */
typedef struct bto_pyramid bto_pyramid;
bto_pyramid* bto_pyramid_create(uint32_t kind, uint32_t format, uint32_t texture_size,
                                uint32_t border_size, uint32_t lod_count);
void     bto_pyramid_destroy(bto_pyramid* p);
/* preprocess_tile: records one dataset (source pixels are borrowed until run) */
void     bto_pyramid_add_dataset(bto_pyramid* p, const bt_dataset* dataset);
/* n datasets queued as ONE job with Preprocessor::preprocess_spherical ordering
 * (preprocessor.rs:314-343): split/downsample all, barrier, then per LOD per dataset stitch */
void     bto_pyramid_add_datasets(bto_pyramid* p, const bt_dataset* datasets, uint32_t n);
/* runs split -> downsample -> stitch for all recorded datasets, reference task order */
void     bto_pyramid_run(bto_pyramid* p, int threads);
void     bto_pyramid_run_phases(bto_pyramid* p, int threads, int phases); /* bit0 split+downsample, bit1 stitch */
void     bto_stitch_tile_raw(uint32_t format, uint32_t T, uint32_t B, const bt_tile* tile, void* data,
                             const bt_tile nb_tiles[8], void* const nb[8]);
void     bto_downsample_tile_raw(uint32_t format, uint32_t T, uint32_t B, void* data, void* const child[4]);
uint32_t bto_pyramid_tile_count(const bto_pyramid* p);
uint32_t bto_pyramid_list_tiles(const bto_pyramid* p, bt_tile* out, uint32_t cap);
/* returns NULL when the tile does not exist */
const void* bto_pyramid_tile(const bto_pyramid* p, const bt_tile* tile);
/* neighbour list of TileCoordinate::neighbours (coordinate.rs:208-279) */
void     bto_tile_neighbours(const bt_tile* tile, int spherical, bt_tile out[8]);

/* ---- mip generation at tile load ---- */
/* out receives mips 1..mip_level_count-1 appended (as the reference appends them after mip 0) */
void bto_generate_mipmaps(uint32_t format, uint32_t texture_size, uint32_t mip_level_count,
                          const void* mip0, void* out);

/* ---- f64 host math ---- */
void bto_model_approximation(const bt_terrain_model* model, const double view_world_position[3],
                             uint32_t origin_lod, float approximate_height, bt_model_approximation* out);
/* src/math/ellipsoid.rs:12-63 with e = (major, major, minor) */
void bto_project_point_ellipsoid(double major_axis, double minor_axis, const double query[3], double out[3]);
void bto_coordinate_from_world_position(const bt_terrain_model* model, const double world[3],
                                        uint32_t* side, double uv[2]);
void bto_coordinate_world_position(const bt_terrain_model* model, uint32_t side, const double uv[2],
                                   float height, double out[3]);
/* TileTree::compute_origin / compute_tile_distance (tile_tree.rs:180-221) */
void   bto_tile_tree_origin(uint32_t tree_size, uint32_t lod, const double uv[2], uint32_t out[2]);
void   bto_project_to_side(const bt_terrain_model* model, uint32_t self_side, const double self_uv[2], uint32_t side, double out_uv[2]);
double bto_tile_distance(const bt_terrain_model* model, const bt_tile* tile, uint32_t view_side,
                         const double view_uv[2], float approximate_height, const double view_world_position[3]);
/* CPU sampler twin: AttachmentData::sample for R16 (terrain_data/mod.rs:221-264) */
float  bto_sample_r16(const uint16_t* tile, uint32_t size, float u, float v);

/* ---- synthetic sources (integer hash noise; NOT part of the reference) ---- */
void bto_synth_source(uint32_t format, uint32_t width, uint32_t height, uint32_t x0, uint32_t y0,
                      uint32_t seed, uint32_t holes, void* out, int threads);

int bto_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

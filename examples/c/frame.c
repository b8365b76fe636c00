/* Plain-C consumer of the whole hot path through the C ABI (needs a B200): source heightmap -> tile pyramid on the GPU
 * (split / downsample / stitch) -> clipmap residency on the device (bt_clipmap) -> tiles into the atlas -> one frame
 * (bt_view_run_frame: select + vertex generation) -> read-back.  This is the sequence a maintainer's binding drives from Rust;
 * tests/test_c_consumer.py runs it and compares the printed hashes with the same frame produced through the Python mirror.
 *
 *   gcc -std=c11 -Wall -Wextra -Werror -I include examples/c/frame.c -L bevy_terrain_b200/csrc -lbevy_terrain_b200 -lm -o frame
 *   LD_LIBRARY_PATH=bevy_terrain_b200/csrc ./frame [device_clipmap: 0|1]                                                        */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bevy_terrain_b200.h"

#define CHECK(ctx, call) do { bt_status st_ = (call); if (st_ != BT_OK) { \
    fprintf(stderr, "%s -> status %d: %s\n", #call, (int)st_, (ctx) ? bt_last_error(ctx) : ""); return 1; } } while (0)

static uint64_t fnv1a(const void* data, size_t bytes, uint64_t h) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

static uint64_t sum32(const void* data, size_t bytes) {   /* order-dependent enough for a cross-check, and trivial to redo in numpy */
    const uint32_t* p = (const uint32_t*)data;
    uint64_t s = 0;
    for (size_t i = 0; i < bytes / 4; ++i) s += (uint64_t)p[i] * (uint64_t)(i % 251u + 1u);
    return s;
}

int main(int argc, char** argv) {
    const int device_clipmap = argc > 1 ? atoi(argv[1]) : 1;
    enum { SRC = 512, LODS = 4, T = 64, B = 2, ATLAS = 256, GRID = 16 };

    bt_context* ctx = NULL;
    if (bt_context_create(0, &ctx) != BT_OK) { fprintf(stderr, "no sm_100a device (there is no CPU fallback)\n"); return 2; }

    /* ---- a deterministic R16 heightmap on the host (no zeros: zero means "no data") ---- */
    uint16_t* source = (uint16_t*)malloc(sizeof(uint16_t) * SRC * SRC);
    for (uint32_t y = 0; y < SRC; ++y)
        for (uint32_t x = 0; x < SRC; ++x) {
            uint32_t v = (x * 37u + y * 91u + ((x * y) >> 3)) % 60000u + 1000u;
            source[y * SRC + x] = (uint16_t)v;
        }

    /* ---- Preprocessor::preprocess_tile + run (preprocessor.rs:298-312, preprocess/mod.rs:143-217) ---- */
    bt_attachment_config height = {T, B, 1, BT_FORMAT_R16};
    bt_preprocessor* pre = NULL;
    CHECK(ctx, bt_preprocessor_create(ctx, BT_KIND_PLANAR, &height, 1, LODS, (bt_shard){0, 1}, &pre));
    bt_dataset ds;
    memset(&ds, 0, sizeof(ds));
    ds.bottom_right[0] = ds.bottom_right[1] = 1.0f;
    ds.lod_end = LODS;
    ds.width = ds.height = SRC;
    ds.pixels = source;
    CHECK(ctx, bt_preprocessor_add_datasets(pre, &ds, 1));
    CHECK(ctx, bt_preprocessor_run(pre, NULL));
    uint32_t n_tiles = 0;
    CHECK(ctx, bt_preprocessor_tile_count(pre, 0, &n_tiles));
    bt_tile* existing = (bt_tile*)malloc(sizeof(bt_tile) * n_tiles);
    CHECK(ctx, bt_preprocessor_list_tiles(pre, 0, existing, n_tiles, &n_tiles));
    uint64_t pyramid_hash = 0;
    CHECK(ctx, bt_preprocessor_hash(pre, 0, &pyramid_hash, NULL));
    printf("pyramid: %u tiles, content hash %016llx\n", n_tiles, (unsigned long long)pyramid_hash);

    /* ---- TerrainConfig / TileAtlas / TerrainViewConfig / TileTree / TerrainViewData ---- */
    const double origin[3] = {0.0, -100.0, 0.0};
    bt_terrain_model model;
    bt_terrain_model_planar(origin, 2000.0, 0.0f, 500.0f, &model);
    bt_terrain_view_config vcfg;
    bt_terrain_view_config_default(&vcfg);
    vcfg.grid_size = GRID;
    vcfg.geometry_tile_count = 1u << 16;

    bt_terrain_desc td;
    memset(&td, 0, sizeof(td));
    td.kind = BT_KIND_PLANAR;
    td.lod_count = LODS;
    td.atlas_size = ATLAS;
    td.attachment_count = 1;
    td.attachments[0] = height;
    td.min_height = model.min_height;
    td.max_height = model.max_height;
    td.scale = bt_terrain_model_scale(&model);
    bt_terrain_model_mesh_transform(&model, &td.transform);
    bt_terrain* terrain = NULL;
    CHECK(ctx, bt_terrain_create(ctx, &td, &terrain));

    bt_tile_tree* tree = NULL;
    CHECK(ctx, bt_tile_tree_create(&model, &vcfg, LODS, ATLAS, &tree));
    CHECK(ctx, bt_tile_tree_add_existing(tree, existing, n_tiles));

    const uint32_t vertices_per_tile = 2u * GRID * (GRID + 2u);
    bt_view_desc vd = {vcfg.geometry_tile_count, vcfg.tree_size, GRID, 0, (uint64_t)8192 * vertices_per_tile};
    bt_view* view = NULL;
    CHECK(ctx, bt_view_create(terrain, &vd, &view));
    bt_clipmap* clipmap = NULL;
    if (device_clipmap) CHECK(ctx, bt_clipmap_create(view, tree, &clipmap));

    /* ---- one frame of the streaming systems: requests -> tiles into their atlas slots -> table ---- */
    const double camera[3] = {100.0, 170.0, -50.0};
    bt_tile_load* loads = (bt_tile_load*)malloc(sizeof(bt_tile_load) * ATLAS);
    uint32_t n_loads = 0;
    if (clipmap) CHECK(ctx, bt_clipmap_update(clipmap, camera, loads, ATLAS, &n_loads, NULL));
    else CHECK(ctx, bt_tile_tree_update(tree, camera, loads, ATLAS, &n_loads));
    for (uint32_t i = 0; i < n_loads; ++i) {
        void* tile_ptr = NULL;
        CHECK(ctx, bt_preprocessor_tile_device_ptr(pre, 0, &loads[i].tile, &tile_ptr));
        CHECK(ctx, bt_atlas_copy_tile(terrain, 0, loads[i].atlas_index, tile_ptr, (size_t)T * T * sizeof(uint16_t), NULL));
        CHECK(ctx, bt_atlas_update_minmax(terrain, loads[i].atlas_index, NULL));
        CHECK(ctx, bt_tile_tree_mark_loaded(tree, &loads[i].tile));
    }
    const uint32_t n_entries = LODS * vcfg.tree_size * vcfg.tree_size;
    bt_tile_tree_entry* entries = (bt_tile_tree_entry*)malloc(sizeof(bt_tile_tree_entry) * n_entries);
    uint32_t origins[2 * LODS];
    if (clipmap) {
        CHECK(ctx, bt_clipmap_entries(clipmap, NULL));
        CHECK(ctx, bt_clipmap_download(clipmap, entries, n_entries, origins, 2 * LODS, NULL));
    } else {
        CHECK(ctx, bt_tile_tree_entries(tree, entries, n_entries, origins, 2 * LODS));
    }
    printf("clipmap (%s): %u loads, table hash %016llx\n", clipmap ? "device" : "host", n_loads,
           (unsigned long long)fnv1a(entries, sizeof(bt_tile_tree_entry) * n_entries, 0xcbf29ce484222325ull));

    /* ---- Prepare stage payload + the frame itself ---- */
    bt_view_frame frame;
    memset(&frame, 0, sizeof(frame));
    bt_view_config_uniform_from(&vcfg, &model, &frame.view_config);
    CHECK(ctx, bt_model_approximation_compute(&model, camera, vcfg.origin_lod, 250.0f, &frame.approximation));
    for (int k = 0; k < 3; ++k) frame.culling.world_position[k] = frame.view_world_position[k] = (float)camera[k];
    frame.tile_tree_entry_count = n_entries;
    frame.tile_tree = clipmap ? NULL : entries;   /* with a device clipmap the table is already where the kernels read it */
    frame.origins = origins;

    bt_tile* tiles = (bt_tile*)malloc(sizeof(bt_tile) * 8192);
    uint32_t n_final = 0;
    bt_indirect indirect;
    bt_vertex_buffer vb;
    const uint32_t gen_flags = BT_FLAG_MORPH | BT_FLAG_BLEND | BT_FLAG_NORMALS;
    CHECK(ctx, bt_view_run_frame(view, &frame, 0, gen_flags, NULL, tiles, 8192, &n_final, &indirect, &vb));
    bt_vertex* verts = (bt_vertex*)malloc(sizeof(bt_vertex) * indirect.vertex_count);
    CHECK(ctx, bt_view_download_vertices(view, verts, 0, indirect.vertex_count, NULL));
    printf("frame: %u tiles, %u vertices\n", n_final, indirect.vertex_count);
    printf("tile list sum %llu\n", (unsigned long long)sum32(tiles, sizeof(bt_tile) * n_final));
    printf("vertex sum %llu\n", (unsigned long long)sum32(verts, sizeof(bt_vertex) * indirect.vertex_count));
    printf("tile list hash %016llx\n", (unsigned long long)fnv1a(tiles, sizeof(bt_tile) * n_final, 0xcbf29ce484222325ull));
    printf("vertex hash %016llx\n", (unsigned long long)fnv1a(verts, sizeof(bt_vertex) * indirect.vertex_count, 0xcbf29ce484222325ull));

    free(verts); free(tiles); free(entries); free(loads); free(existing); free(source);
    if (clipmap) bt_clipmap_destroy(clipmap);
    bt_view_destroy(view);
    bt_tile_tree_destroy(tree);
    bt_terrain_destroy(terrain);
    bt_preprocessor_destroy(pre);
    bt_context_destroy(ctx);
    return 0;
}

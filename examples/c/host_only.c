/* Plain-C consumer of the C ABI (no GPU needed for this part): proves that include/bevy_terrain_b200.h is valid C11 and that
 * the host-side entry points link and run from C.  It builds a terrain model, runs the clipmap for one camera position,
 * computes the Taylor approximation, plans a sharded cube preprocess, and (when a file is given) decodes a source image.
 *
 *   gcc -std=c11 -Wall -Wextra -Werror -I include examples/c/host_only.c -L bevy_terrain_b200/csrc -lbevy_terrain_b200 -o host_only
 *   LD_LIBRARY_PATH=bevy_terrain_b200/csrc ./host_only [image.png|image.tif]
 * With a GPU the same program continues with bt_context_create / bt_terrain_create / bt_view_run_frame. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bevy_terrain_b200.h"

#define CHECK(call) do { bt_status st_ = (call); if (st_ != BT_OK) { fprintf(stderr, "%s -> status %d\n", #call, (int)st_); return 1; } } while (0)

int main(int argc, char** argv) {
    printf("ABI version %u\n", bt_abi_version());

    /* TerrainModel::ellipsoid + TerrainViewConfig::default (terrain_model.rs:106-128, terrain_view.rs:47-64) */
    const double origin[3] = {0.0, 0.0, 0.0};
    bt_terrain_model model;
    bt_terrain_model_ellipsoid(origin, 6378137.0, 6356752.314245, -12000.0f, 9000.0f, &model);
    bt_terrain_view_config view_config;
    bt_terrain_view_config_default(&view_config);
    view_config.grid_size = 64;
    printf("model scale %.3f\n", bt_terrain_model_scale(&model));

    /* clipmap for one frame: which tiles does the view want, and the table the vertex kernel reads */
    enum { LODS = 5, ATLAS = 6 * LODS * 64 };
    bt_tile_tree* tree = NULL;
    CHECK(bt_tile_tree_create(&model, &view_config, LODS, ATLAS, &tree));
    bt_tile* existing = (bt_tile*)malloc(sizeof(bt_tile) * 6 * 341);
    uint32_t n_existing = 0;
    for (uint32_t side = 0; side < 6; ++side)
        for (uint32_t lod = 0; lod < LODS; ++lod)
            for (uint32_t x = 0; x < (1u << lod); ++x)
                for (uint32_t y = 0; y < (1u << lod); ++y) existing[n_existing++] = (bt_tile){side, lod, x, y};
    CHECK(bt_tile_tree_add_existing(tree, existing, n_existing));
    const double camera[3] = {-(6378137.0 + 30000.0), 1200.0, -800.0};
    bt_tile_load* loads = (bt_tile_load*)malloc(sizeof(bt_tile_load) * ATLAS);
    uint32_t n_loads = 0;
    CHECK(bt_tile_tree_update(tree, camera, loads, ATLAS, &n_loads));
    for (uint32_t i = 0; i < n_loads; ++i) CHECK(bt_tile_tree_mark_loaded(tree, &loads[i].tile));
    const uint32_t n_entries = 6 * LODS * view_config.tree_size * view_config.tree_size;
    bt_tile_tree_entry* entries = (bt_tile_tree_entry*)malloc(sizeof(bt_tile_tree_entry) * n_entries);
    uint32_t origins[6 * LODS * 2];
    CHECK(bt_tile_tree_entries(tree, entries, n_entries, origins, 6 * LODS * 2));
    printf("clipmap: %u tiles requested, table of %u entries\n", n_loads, n_entries);

    /* TerrainModelApproximation::compute -> the 688-byte uniform of the view */
    bt_model_approximation approximation;
    CHECK(bt_model_approximation_compute(&model, camera, view_config.origin_lod, -1500.0f, &approximation));
    printf("approximation: origin_lod %u, side 0 view_xy (%d, %d)\n", approximation.origin_lod, approximation.sides[0].view_xy[0],
           approximation.sides[0].view_xy[1]);

    /* a sharded cube preprocess, planned on the host (ctx = NULL): rank 1 of 4 */
    const bt_attachment_config attachment = {512, 2, 4, BT_FORMAT_R16};
    bt_preprocessor* pre = NULL;
    const bt_shard shard = {1, 4};
    CHECK(bt_preprocessor_create(NULL, BT_KIND_SPHERICAL, &attachment, 1, 8, shard, &pre));
    bt_dataset faces[6];
    memset(faces, 0, sizeof(faces));
    for (uint32_t side = 0; side < 6; ++side) {
        faces[side].side = side;
        faces[side].bottom_right[0] = faces[side].bottom_right[1] = 1.0f;
        faces[side].lod_end = 6;
        faces[side].width = faces[side].height = 16384;
    }
    CHECK(bt_preprocessor_add_datasets(pre, faces, 6));
    CHECK(bt_preprocessor_plan(pre));
    uint32_t n_tiles = 0;
    uint64_t halo = 0, roots = 0;
    CHECK(bt_preprocessor_plan_tiles(pre, 0, NULL, NULL, 0, &n_tiles));
    CHECK(bt_preprocessor_halo_bytes(pre, &halo));
    CHECK(bt_preprocessor_roots_bytes(pre, &roots));
    printf("sharded cube plan (rank 1 of 4): %u tile slots, halo %llu B, sub-root gather %llu B per rank\n", n_tiles,
           (unsigned long long)halo, (unsigned long long)roots);
    bt_preprocessor_destroy(pre);

    if (argc > 1) {   /* PreprocessDataset::path: decode a PNG / TIFF source */
        bt_image image;
        char error[256];
        if (bt_image_load(argv[1], &image, error, sizeof(error)) != BT_OK) { fprintf(stderr, "%s\n", error); return 2; }
        printf("%s: %u x %u, %s\n", argv[1], image.width, image.height, image.format == BT_FORMAT_R16 ? "R16" : "Rgba8");
        bt_image_free(&image);
    }

    bt_context* ctx = NULL;   /* without a GPU this reports BT_ERR_CUDA: there is no CPU fallback */
    const bt_status st = bt_context_create(0, &ctx);
    printf("bt_context_create -> %d%s\n", (int)st, st == BT_OK ? "" : " (no CUDA device: the per-frame path is not available)");
    if (st == BT_OK) bt_context_destroy(ctx);

    bt_tile_tree_destroy(tree);
    free(entries); free(loads); free(existing);
    return 0;
}

/*
 * bt_oracle.c — CPU oracle (plain C) for the bevy_terrain hot path.  TEST INFRASTRUCTURE ONLY
 * (see bt_oracle.h).  PARITY UNPINNED: the reference has no tests / golden vectors.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 * Numerics contract (WGSL semantics, fixed where the spec leaves latitude):
 *   - all device math is IEEE binary32, one rounding per operation, NO fused multiply-add
 *     (build with -ffp-contract=off), evaluation order exactly as written below
 *   - saturate(x) = min(max(x,0),1); mix(a,b,t) = a*(1-t) + b*t; f32 %1.0 = x - trunc(x)
 *   - u32(f32) saturates to [0, 2^32-1], NaN -> 0
 *   - normalize(v) = v / sqrt((x*x + y*y) + z*z); sqrt and / correctly rounded
 *   - log2 = bto_log2f below (deterministic, <= 2 ulp vs libm), because vendor log2 differ
 *   - unorm decode = (float)v / 65535.0f (or 255.0f); pack = floor(0.5 + M*clamp(v,0,1))
 *   - bilinear = exact f32 weights, texel centres at i+0.5, clamp-to-edge
 */
#include "bt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ============================================================================================ */
/* numerics primitives                                                                          */
/* ============================================================================================ */

static inline uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* log2 for positive finite normal x: fdlibm-style logf core, then * 1/ln2 + exponent.
 * Uses only + - * / in a fixed order so that the CUDA kernels reproduce it bit-exactly.      */
float bto_log2f(float x) {
    if (!(x > 0.0f)) return (x == 0.0f) ? -INFINITY : NAN;
    if (isinf(x)) return x;
    uint32_t ix = f2bits(x);
    int32_t k = 0;
    if (ix < 0x00800000u) { /* subnormal */
        x = x * 8388608.0f;
        ix = f2bits(x);
        k = -23;
    }
    /* reduce to m in [sqrt(2)/2, sqrt(2)) */
    ix += 0x3f800000u - 0x3f3504f3u;
    k += (int32_t)(ix >> 23) - 0x7f;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    float m = bits2f(ix);
    float f = m - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float w = z * z;
    float t1 = w * (0.40000972152f + w * 0.24279078841f);
    float t2 = z * (0.66666662693f + w * 0.28498786688f);
    float R = t2 + t1;
    float hfsq = 0.5f * f * f;
    float lg = f - (hfsq - s * (hfsq + R)); /* ln(m) */
    return (float)k + lg * 1.44269504089f;
}

uint32_t bto_f32_to_u32(float x) {
    if (!(x > 0.0f)) return 0u; /* negative, zero, NaN */
    if (x >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)x;
}

static inline float o_saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
static inline float o_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float o_fract_trunc(float x) { return x - truncf(x); }
/* tile_count(lod) = f32(1u << lod)  (functions.wgsl:156) */
static inline float o_tile_count(uint32_t lod) { return (float)(1u << (lod & 31u)); }
/* inverse_mix (functions.wgsl:31-33, util.rs:8-10) */
static inline float o_inverse_mix(float a, float b, float value) { return o_saturate((value - a) / (b - a)); }

/* C_SQR = 0.87 * 0.87 is an abstract-float constant in WGSL (functions.wgsl:12): evaluated in
 * f64, then converted.  `1.0 + C_SQR` is likewise folded before conversion.                   */
static const float C_SQR_F = (float)(0.87 * 0.87);
static const float ONE_PLUS_C_SQR_F = (float)(1.0 + 0.87 * 0.87);

typedef struct coord { uint32_t side, lod, xy[2]; float uv[2]; } coord;

/* coordinate_change_lod — functions.wgsl:164-188 */
static void change_lod(coord* c, uint32_t new_lod) {
    int32_t d = (int32_t)new_lod - (int32_t)c->lod;
    if (d == 0) return;
    uint32_t ad = (uint32_t)(d < 0 ? -d : d);
    uint32_t delta_count = 1u << (ad & 31u);
    float delta_size = ldexpf(1.0f, d); /* pow(2.0, f32(d)), exact */
    c->lod = new_lod;
    if (d > 0) {
        for (int k = 0; k < 2; ++k) {
            float scaled = c->uv[k] * delta_size;
            c->xy[k] = c->xy[k] * delta_count + bto_f32_to_u32(scaled);
            c->uv[k] = o_fract_trunc(scaled);
        }
    } else {
        for (int k = 0; k < 2; ++k) {
            uint32_t xy = c->xy[k];
            c->xy[k] = xy / delta_count;
            c->uv[k] = ((float)(xy % delta_count) + c->uv[k]) * delta_size;
        }
    }
}

void bto_coordinate_change_lod(uint32_t* lod, uint32_t xy[2], float uv[2], uint32_t new_lod) {
    coord c = {0, *lod, {xy[0], xy[1]}, {uv[0], uv[1]}};
    change_lod(&c, new_lod);
    *lod = c.lod; xy[0] = c.xy[0]; xy[1] = c.xy[1]; uv[0] = c.uv[0]; uv[1] = c.uv[1];
}

static inline void o_normalize(const float v[3], float out[3]) {
    float len = sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    out[0] = v[0] / len; out[1] = v[1] / len; out[2] = v[2] / len;
}
static inline float o_length(const float v[3]) { return sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

/* compute_local_position — functions.wgsl:73-96 */
static void local_position(int spherical, const coord* c, float out[3]) {
    float tc = o_tile_count(c->lod);
    float u = ((float)c->xy[0] + c->uv[0]) / tc;
    float v = ((float)c->xy[1] + c->uv[1]) / tc;
    if (!spherical) {
        out[0] = u - 0.5f; out[1] = 0.0f; out[2] = v - 0.5f;
        return;
    }
    u = (u - 0.5f) / 0.5f;
    v = (v - 0.5f) / 0.5f;
    u = u / sqrtf(ONE_PLUS_C_SQR_F - C_SQR_F * u * u);
    v = v / sqrtf(ONE_PLUS_C_SQR_F - C_SQR_F * v * v);
    float p[3] = {0.0f, 0.0f, 0.0f};
    switch (c->side) {
        case 0: p[0] = -1.0f; p[1] = -v;    p[2] = u;     break;
        case 1: p[0] = u;     p[1] = -v;    p[2] = 1.0f;  break;
        case 2: p[0] = u;     p[1] = 1.0f;  p[2] = v;     break;
        case 3: p[0] = 1.0f;  p[1] = -u;    p[2] = v;     break;
        case 4: p[0] = v;     p[1] = -u;    p[2] = -1.0f; break;
        case 5: p[0] = v;     p[1] = -1.0f; p[2] = u;     break;
        default: break;
    }
    o_normalize(p, out);
}

void bto_local_position(int spherical, uint32_t side, uint32_t lod, const uint32_t xy[2],
                        const float uv[2], float out[3]) {
    coord c = {side, lod, {xy[0], xy[1]}, {uv[0], uv[1]}};
    local_position(spherical, &c, out);
}

/* position_local_to_world — functions.wgsl:26-29 (affine3_to_square(world_from_local) * vec4(p,1)) */
static inline void position_local_to_world(const bt_mesh_transform* m, const float p[3], float out[3]) {
    const float* a = m->world_from_local;
    for (int r = 0; r < 3; ++r)
        out[r] = ((a[0 + r] * p[0] + a[3 + r] * p[1]) + a[6 + r] * p[2]) + a[9 + r];
}

/* normal_local_to_world — functions.wgsl:14-24 */
static inline void normal_local_to_world(const bt_mesh_transform* m, int spherical, const float lp[3], float out[3]) {
    float n[3];
    if (spherical) { n[0] = lp[0]; n[1] = lp[1]; n[2] = lp[2]; }
    else { n[0] = 0.0f; n[1] = 1.0f; n[2] = 0.0f; }
    const float* a = m->normal_matrix;
    float t[3];
    for (int r = 0; r < 3; ++r) t[r] = (a[0 + r] * n[0] + a[3 + r] * n[1]) + a[6 + r] * n[2];
    o_normalize(t, out);
}

/* compute_relative_position — functions.wgsl:98-115 */
static void relative_position(const bt_model_approximation* ap, coord c, float out[3]) {
    change_lod(&c, ap->origin_lod);
    const bt_side_parameter* p = &ap->sides[c.side];
    float tc = o_tile_count(ap->origin_lod);
    float s = (((float)((int32_t)c.xy[0] - p->view_xy[0]) + c.uv[0]) - p->view_uv[0]) / tc;
    float t = (((float)((int32_t)c.xy[1] - p->view_xy[1]) + c.uv[1]) - p->view_uv[1]) / tc;
    for (int k = 0; k < 3; ++k)
        out[k] = ((((p->c[k] + p->c_s[k] * s) + p->c_t[k] * t) + p->c_ss[k] * s * s) + p->c_st[k] * s * t) + p->c_tt[k] * t * t;
}

/* ============================================================================================ */
/* select: prepare_root / refine_tiles / prepare_next / prepare_render                          */
/* (prepare_prepass.wgsl:4-43, refine_tiles.wgsl:5-44), sequential-invocation semantics          */
/* ============================================================================================ */

/* compute_subdivision_coordinate — functions.wgsl:133-154 */
static coord subdivision_coordinate(const bt_model_approximation* ap, const bt_tile* tile) {
    const bt_side_parameter* p = &ap->sides[tile->side];
    coord v = {tile->side, ap->origin_lod, {(uint32_t)p->view_xy[0], (uint32_t)p->view_xy[1]}, {p->view_uv[0], p->view_uv[1]}};
    change_lod(&v, tile->lod);
    int32_t ox = (int32_t)v.xy[0] - (int32_t)tile->x;
    int32_t oy = (int32_t)v.xy[1] - (int32_t)tile->y;
    float u = v.uv[0], w = v.uv[1];
    if (ox < 0) u = 0.0f; else if (ox > 0) u = 1.0f;
    if (oy < 0) w = 0.0f; else if (oy > 0) w = 1.0f;
    coord out = {tile->side, tile->lod, {tile->x, tile->y}, {u, w}};
    return out;
}

/* approximate_view_distance without HIGH_PRECISION (compiled out in the prepass:
 * tiling_prepass.rs:61-78) — functions.wgsl:117-131 */
static float approximate_view_distance(const bt_model_approximation* ap, const bt_mesh_transform* mesh,
                                       int spherical, const coord* c, const float view[3]) {
    float lp[3], wp[3], wn[3], q[3];
    local_position(spherical, c, lp);
    position_local_to_world(mesh, lp, wp);
    normal_local_to_world(mesh, spherical, lp, wn);
    for (int k = 0; k < 3; ++k) q[k] = (wp[k] + ap->approximate_height * wn[k]) - view[k];
    return o_length(q);
}

/* should_be_divided — refine_tiles.wgsl:17-22 */
static int should_be_divided(const bt_view_config_uniform* vc, const bt_model_approximation* ap,
                             const bt_culling_uniform* cull, const bt_mesh_transform* mesh,
                             int spherical, const bt_tile* tile) {
    coord c = subdivision_coordinate(ap, tile);
    float d = approximate_view_distance(ap, mesh, spherical, &c, cull->world_position);
    return d < vc->subdivision_distance / o_tile_count(tile->lod);
}

/* EXTENSION (default off, not in the reference at this commit: SURVEY F4): frustum culling of tiles against the five
 * planes of culling_bind_group.rs:25-38, with the bounding volume docs/implementation.md:69-70 describes approximated
 * by ten points: the four tile corners and the tile centre, each at min_height and max_height.  A tile is culled when
 * all ten points lie on the negative side of one plane. */
/* ---- min / max-height culling (BT_FLAG_MINMAX_CULL; docs/implementation.md:69: "The frustum culling uses a 2D min-max height data
 * attachment to approximate the bounding volumes of each tile") ----
 * Per resident atlas tile a small table: entry 0 = (min, max) of the tile's centre region, entries 1..64 = an 8 x 8 grid of cells over
 * it, every region grown by one texel (bilinear taps reach the neighbours) and zeros (no-data) ignored.  A geometry tile is bounded by
 * the best resident atlas tile that CONTAINS it (the tile-tree entry of its centre at lod min(tile.lod, lod_count - 1)): the cell it
 * falls into when it is at most 1/8 of the atlas tile wide, else the whole-tile entry; without a resident tile the terrain-wide
 * [min_height, max_height] stays.  Conservative by construction: the bound only ever covers a superset of the tile. */
#define BTO_MINMAX_CELLS 8
void bto_tile_minmax(const uint16_t* tile, uint32_t T, uint32_t B, uint16_t out[(1 + BTO_MINMAX_CELLS * BTO_MINMAX_CELLS) * 2]) {
    const uint32_t C_ = T - 2 * B;
    for (uint32_t e = 0; e < 1 + BTO_MINMAX_CELLS * BTO_MINMAX_CELLS; ++e) {
        uint32_t x0, x1, y0, y1;
        if (e == 0) { x0 = B; x1 = B + C_; y0 = B; y1 = B + C_; }
        else {
            const uint32_t cx = (e - 1) % BTO_MINMAX_CELLS, cy = (e - 1) / BTO_MINMAX_CELLS;
            x0 = B + cx * C_ / BTO_MINMAX_CELLS; x1 = B + (cx + 1) * C_ / BTO_MINMAX_CELLS;
            y0 = B + cy * C_ / BTO_MINMAX_CELLS; y1 = B + (cy + 1) * C_ / BTO_MINMAX_CELLS;
        }
        x0 = x0 > 0 ? x0 - 1 : 0; y0 = y0 > 0 ? y0 - 1 : 0;
        x1 = x1 + 1 < T ? x1 + 1 : T; y1 = y1 + 1 < T ? y1 + 1 : T;
        uint32_t mn = 0xFFFFu, mx = 0u;
        for (uint32_t y = y0; y < y1; ++y) for (uint32_t x = x0; x < x1; ++x) {
            const uint32_t v = tile[(size_t)y * T + x];
            if (v == 0) continue;
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        if (mx == 0u) mn = 0u;   /* no valid texel: (0, 0) */
        out[2 * e] = (uint16_t)mn; out[2 * e + 1] = (uint16_t)mx;
    }
}

typedef struct minmax_ctx {
    const bt_tile_tree_entry* tile_tree;
    const uint16_t* table;        /* [atlas_size][65][2] */
    uint32_t lod_count, tree_size, atlas_size;
    float terrain_min, terrain_max;
} minmax_ctx;

static void tile_height_bounds(const minmax_ctx* mc, const bt_tile* tile, float* hmin, float* hmax) {
    *hmin = mc->terrain_min; *hmax = mc->terrain_max;
    if (!mc->table || !mc->tile_tree) return;
    coord c = {tile->side, tile->lod, {tile->x, tile->y}, {0.5f, 0.5f}};
    const uint32_t L = tile->lod < mc->lod_count - 1u ? tile->lod : mc->lod_count - 1u;
    change_lod(&c, L);
    const uint32_t ts = mc->tree_size;
    const bt_tile_tree_entry e = mc->tile_tree[((c.side * mc->lod_count + c.lod) * ts + c.xy[0] % ts) * ts + c.xy[1] % ts];
    if (e.atlas_index >= mc->atlas_size || e.atlas_lod > L) return;       /* nothing resident (INVALID) or a malformed entry */
    change_lod(&c, e.atlas_lod);
    const uint32_t depth = tile->lod - e.atlas_lod;                       /* the tile is 2^-depth of the atlas tile wide */
    uint32_t entry = 0;
    if (depth >= 3u) {
        uint32_t cx = bto_f32_to_u32(c.uv[0] * (float)BTO_MINMAX_CELLS), cy = bto_f32_to_u32(c.uv[1] * (float)BTO_MINMAX_CELLS);
        if (cx > BTO_MINMAX_CELLS - 1u) cx = BTO_MINMAX_CELLS - 1u;
        if (cy > BTO_MINMAX_CELLS - 1u) cy = BTO_MINMAX_CELLS - 1u;
        entry = 1u + cy * BTO_MINMAX_CELLS + cx;
    }
    const uint16_t* mm = mc->table + ((size_t)e.atlas_index * (1 + BTO_MINMAX_CELLS * BTO_MINMAX_CELLS) + entry) * 2;
    if (mm[1] == 0) return;                                               /* no valid texel in the region */
    *hmin = o_mix(mc->terrain_min, mc->terrain_max, (float)mm[0] / 65535.0f);
    *hmax = o_mix(mc->terrain_min, mc->terrain_max, (float)mm[1] / 65535.0f);
}

static int tile_outside_frustum(const bt_culling_uniform* cull, const bt_mesh_transform* mesh, int spherical,
                                float min_height, float max_height, const bt_tile* tile) {
    static const float UV[5][2] = {{0.0f, 0.0f}, {1.0f, 0.0f}, {0.0f, 1.0f}, {1.0f, 1.0f}, {0.5f, 0.5f}};
    float pts[10][3];
    for (int i = 0; i < 5; ++i) {
        coord c = {tile->side, tile->lod, {tile->x, tile->y}, {UV[i][0], UV[i][1]}};
        float lp[3], wp[3], wn[3];
        local_position(spherical, &c, lp);
        position_local_to_world(mesh, lp, wp);
        normal_local_to_world(mesh, spherical, lp, wn);
        for (int k = 0; k < 3; ++k) { pts[2 * i][k] = wp[k] + min_height * wn[k]; pts[2 * i + 1][k] = wp[k] + max_height * wn[k]; }
    }
    for (int p = 0; p < 5; ++p) {
        const float* pl = cull->planes[p];
        int all_out = 1;
        for (int i = 0; i < 10; ++i) {
            float dist = ((pl[0] * pts[i][0] + pl[1] * pts[i][1]) + pl[2] * pts[i][2]) + pl[3];
            if (!(dist < 0.0f)) { all_out = 0; break; }
        }
        if (all_out) return 1;
    }
    return 0;
}

void bto_select_tiles(const bt_view_config_uniform* vc, const bt_model_approximation* approx,
                      const bt_culling_uniform* culling, const bt_mesh_transform* mesh,
                      int spherical, bt_tile* final_tiles, uint32_t capacity,
                      bt_indirect* indirect, bto_select_result* result) {
    bto_select_tiles_ex(vc, approx, culling, mesh, spherical, 0, 0.0f, 0.0f, final_tiles, capacity, indirect, result);
}

void bto_select_tiles_ex(const bt_view_config_uniform* vc, const bt_model_approximation* approx,
                         const bt_culling_uniform* culling, const bt_mesh_transform* mesh,
                         int spherical, int frustum_cull, float min_height, float max_height,
                         bt_tile* final_tiles, uint32_t capacity, bt_indirect* indirect, bto_select_result* result) {
    bto_select_tiles_minmax(vc, approx, culling, mesh, spherical, frustum_cull, min_height, max_height, NULL, NULL, 0, 0,
                            final_tiles, capacity, indirect, result);
}

void bto_select_tiles_minmax(const bt_view_config_uniform* vc, const bt_model_approximation* approx,
                             const bt_culling_uniform* culling, const bt_mesh_transform* mesh,
                             int spherical, int frustum_cull, float min_height, float max_height,
                             const bt_tile_tree_entry* tile_tree, const uint16_t* minmax_table, uint32_t lod_count, uint32_t atlas_size,
                             bt_tile* final_tiles, uint32_t capacity, bt_indirect* indirect, bto_select_result* result) {
    minmax_ctx mc = {tile_tree, minmax_table, lod_count, vc->tree_size, atlas_size, min_height, max_height};
    /* The reference keeps parents and children in ONE buffer used as a double-ended stack
     * (refine_tiles.wgsl:9-11); executed with invocations in id order this is a FIFO queue:
     * children are read back in the order they were written.  We keep two growable arrays.   */
    size_t cap_a = 64, cap_b = 64;
    bt_tile* cur = (bt_tile*)malloc(cap_a * sizeof(bt_tile));
    bt_tile* nxt = (bt_tile*)malloc(cap_b * sizeof(bt_tile));
    uint32_t n_cur = spherical ? 6u : 1u; /* prepare_root */
    for (uint32_t i = 0; i < n_cur; ++i) { bt_tile t = {i, 0, 0, 0}; cur[i] = t; }
    uint32_t n_final = 0, overflow = 0, max_frontier = n_cur, dropped = 0;

    for (uint32_t pass = 0; pass <= vc->refinement_count; ++pass) { /* tiling_prepass.rs:251-260 */
        uint32_t n_nxt = 0;
        for (uint32_t i = 0; i < n_cur; ++i) {
            bt_tile tile = cur[i];
            if (frustum_cull) {
                float hmin, hmax;
                tile_height_bounds(&mc, &tile, &hmin, &hmax);      /* the terrain-wide range unless a min / max table is given */
                if (tile_outside_frustum(culling, mesh, spherical, hmin, hmax, &tile)) continue;
            }
            if (should_be_divided(vc, approx, culling, mesh, spherical, &tile)) {
                if ((size_t)n_nxt + 4 > cap_b) { cap_b *= 2; nxt = (bt_tile*)realloc(nxt, cap_b * sizeof(bt_tile)); }
                for (uint32_t c = 0; c < 4; ++c) { /* subdivide — refine_tiles.wgsl:24-31 */
                    bt_tile ch = {tile.side, tile.lod + 1u, (tile.x << 1) + (c & 1u), (tile.y << 1) + ((c >> 1) & 1u)};
                    nxt[n_nxt++] = ch;
                }
            } else {
                if (n_final < capacity) final_tiles[n_final] = tile; else overflow = 1;
                n_final++;
            }
        }
        if (pass == vc->refinement_count) { dropped = n_nxt / 4; break; } /* children of the last pass are never read */
        if (n_nxt > capacity) overflow = 1;
        bt_tile* tp = cur; cur = nxt; nxt = tp;
        size_t tc = cap_a; cap_a = cap_b; cap_b = tc;
        n_cur = n_nxt;
        if (n_cur > max_frontier) max_frontier = n_cur;
        if (n_cur == 0) { /* remaining passes are no-ops */ }
    }
    free(cur); free(nxt);
    if (n_final > capacity) overflow = 1;
    uint32_t emitted = n_final > capacity ? capacity : n_final;
    /* prepare_render — prepare_prepass.wgsl:38-43 */
    indirect->vertex_count = vc->vertices_per_tile * emitted;
    indirect->instance_count = 1; indirect->base_vertex = 0; indirect->base_instance = 0;
    result->final_count = emitted; result->dropped_parents = dropped;
    result->max_frontier = max_frontier; result->overflow = overflow;
}

/* ============================================================================================ */
/* vertex generation — vertex.wgsl:30-98, functions.wgsl:35-71,190-246, attachments.wgsl         */
/* ============================================================================================ */

/* bilinear R16 fetch with exact f32 weights; formula order follows the CPU twin
 * AttachmentData::sample (terrain_data/mod.rs:221-264): lerp along y first, then x, with
 * glam's lerp a + (b - a) * s.  Addressing follows the GPU (floor + clamp-to-edge), which is
 * identical to the twin whenever uv*size - 0.5 >= 0.                                          */
static float sample_r16_layer(const uint16_t* layer, uint32_t size, float u, float v) {
    float px = u * (float)size - 0.5f;
    float py = v * (float)size - 0.5f;
    float fx = floorf(px), fy = floorf(py);
    float rx = px - fx, ry = py - fy;
    int32_t ix = (int32_t)fx, iy = (int32_t)fy;
    int32_t m = (int32_t)size - 1;
    int32_t x0 = ix < 0 ? 0 : (ix > m ? m : ix), x1 = ix + 1 < 0 ? 0 : (ix + 1 > m ? m : ix + 1);
    int32_t y0 = iy < 0 ? 0 : (iy > m ? m : iy), y1 = iy + 1 < 0 ? 0 : (iy + 1 > m ? m : iy + 1);
    float v00 = (float)layer[(size_t)y0 * size + x0] / 65535.0f;
    float v01 = (float)layer[(size_t)y1 * size + x0] / 65535.0f; /* values[0][1]: x, y+1 */
    float v10 = (float)layer[(size_t)y0 * size + x1] / 65535.0f;
    float v11 = (float)layer[(size_t)y1 * size + x1] / 65535.0f;
    float a = v00 + (v01 - v00) * ry;
    float b = v10 + (v11 - v10) * ry;
    return a + (b - a) * rx;
}

float bto_sample_r16(const uint16_t* tile, uint32_t size, float u, float v) {
    return sample_r16_layer(tile, size, u, v);
}

typedef struct atlas_tile { uint32_t index; coord c; } atlas_tile;

/* lookup_tile (default variant) — functions.wgsl:232-246, 198-206 */
static atlas_tile lookup_tile(const bto_vertex_inputs* in, coord c, uint32_t blend_lod, uint32_t lod_offset) {
    change_lod(&c, blend_lod - lod_offset);
    uint32_t ts = in->view_config.tree_size;
    uint32_t tx = c.xy[0] % ts, ty = c.xy[1] % ts;
    uint32_t idx = ((c.side * in->terrain_config.lod_count + c.lod) * ts + tx) * ts + ty;
    bt_tile_tree_entry e = in->tile_tree[idx];
    change_lod(&c, e.atlas_lod);
    atlas_tile t; t.index = e.atlas_index; t.c = c;
    return t;
}

/* sample_height — attachments.wgsl:7-10,12-24,45-49 */
static float sample_height_uv(const bto_vertex_inputs* in, uint32_t atlas_index, float u, float v) {
    uint32_t T = in->texture_size;
    /* textureSampleLevel clamps array_index to [0, layers-1] (WGSL): an INVALID entry (0xFFFFFFFF) reads the last layer */
    if (in->atlas_size && atlas_index > in->atlas_size - 1u) atlas_index = in->atlas_size - 1u;
    const uint16_t* layer = in->height_atlas + (size_t)atlas_index * T * T;
    float s = sample_r16_layer(layer, T, u, v);
    return o_mix(in->terrain_config.min_height, in->terrain_config.max_height, s);
}

static float sample_height(const bto_vertex_inputs* in, const atlas_tile* t) {
    float u = t->c.uv[0] * in->attachments[0].scale + in->attachments[0].offset;
    float v = t->c.uv[1] * in->attachments[0].scale + in->attachments[0].offset;
    return sample_height_uv(in, t->index, u, v);
}

static inline void o_cross(const float a[3], const float b[3], float out[3]) {
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
}

/* sample_normal — attachments.wgsl:51-107 (non-FRAGMENT branch: textureSampleLevel, level 0) */
static void sample_normal(const bto_vertex_inputs* in, const atlas_tile* t, const float vertex_normal[3], float out[3]) {
    float u = t->c.uv[0] * in->attachments[0].scale + in->attachments[0].offset;
    float v = t->c.uv[1] * in->attachments[0].scale + in->attachments[0].offset;
    float side_length;
    if (in->spherical) side_length = (float)(3.14159265359 / 4.0) * in->terrain_config.scale;
    else side_length = in->terrain_config.scale;
    float pixels_per_side = in->attachments[0].size * o_tile_count(t->c.lod);
    float distance_between_samples = side_length / pixels_per_side;
    float offset = 0.5f / in->attachments[0].size;
    float left  = sample_height_uv(in, t->index, u + -offset, v + 0.0f);
    float up    = sample_height_uv(in, t->index, u + 0.0f, v + -offset);
    float right = sample_height_uv(in, t->index, u + offset, v + 0.0f);
    float down  = sample_height_uv(in, t->index, u + 0.0f, v + offset);
    float sv[3] = {left - right, down - up, distance_between_samples};
    float sn[3];
    o_normalize(sv, sn);
    float r[3];
    if (in->spherical) {
        static const float FACE_UP[6][3] = {{0, 1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, -1}, {-1, 0, 0}, {-1, 0, 0}};
        float n[3], tg[3], bt[3];
        o_normalize(vertex_normal, n);
        o_cross(FACE_UP[t->c.side], n, tg);
        o_cross(n, tg, bt);
        for (int k = 0; k < 3; ++k) r[k] = (tg[k] * sn[0] + bt[k] * sn[1]) + n[k] * sn[2];
    } else {
        /* TBN = mat3x3(1,0,0, 0,0,1, 0,1,0): columns (1,0,0),(0,0,1),(0,1,0) */
        r[0] = sn[0]; r[1] = sn[2]; r[2] = sn[1];
    }
    o_normalize(r, out);
}

/* ============================================================================================ */
/* configs[2]: "trilinear + anisotropic" sampling and per-vertex albedo (BT_FLAG_FILTER_TRILINEAR_ANISO,     */
/* BT_FLAG_ALBEDO).  The reference's vertex stage samples mip 0 only; its SAMPLER is linear / linear / linear */
/* with anisotropy_clamp 16 (src/render/terrain_bind_group.rs:120-126) and its fragment stage uses          */
/* textureSampleGrad (src/shaders/attachments.wgsl:16-20).  This is that sampler evaluated per vertex with   */
/* the footprint of one grid step, written down once here:      */
/*   P = footprint of one grid step in texels of mip 0, per axis;  N = min(ceil(Pmax / Pmin), 16) taps along */
/*   the major axis;  lod = clamp(log2(Pmax / N), 0, mips - 1);  every tap = mix(bilinear(floor(lod)),       */
/*   bilinear(floor(lod) + 1), fract(lod));  result = mean of the taps.  Mips are generate_mipmaps' chain.   */
/* ============================================================================================ */
static float srgb_lut[256];
static int srgb_lut_ready = 0;
static void srgb_lut_init(void) {   /* Rgba8UnormSrgb decode per texel BEFORE filtering (sRGB EOTF, IEC 61966-2-1) */
    if (srgb_lut_ready) return;
    for (int i = 0; i < 256; ++i) {
        double c = (double)i / 255.0;
        srgb_lut[i] = (float)(c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4));
    }
    srgb_lut_ready = 1;
}

static size_t mip_offset_texels(uint32_t T, uint32_t level) {   /* texels before `level` in the appended chain (levels 1..) */
    size_t off = 0;
    for (uint32_t j = 1; j < level; ++j) off += (size_t)(T >> j) * (T >> j);
    return off;
}

/* one bilinear tap at one mip level; R16: out[0]; Rgba8: four channels (rgb through the sRGB table) */
static void bilinear_level(const void* level_texels, int rgba, uint32_t size, float u, float v, float out[4]) {
    float px = u * (float)size - 0.5f, py = v * (float)size - 0.5f;
    float fx = floorf(px), fy = floorf(py);
    float rx = px - fx, ry = py - fy;
    int32_t ix = (int32_t)fx, iy = (int32_t)fy, m = (int32_t)size - 1;
    int32_t x0 = ix < 0 ? 0 : (ix > m ? m : ix), x1 = ix + 1 < 0 ? 0 : (ix + 1 > m ? m : ix + 1);
    int32_t y0 = iy < 0 ? 0 : (iy > m ? m : iy), y1 = iy + 1 < 0 ? 0 : (iy + 1 > m ? m : iy + 1);
    const int32_t xs[2] = {x0, x1}, ys[2] = {y0, y1};
    float t[2][2][4];   /* [x][y][channel] */
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
        size_t idx = (size_t)ys[b] * size + (size_t)xs[a];
        if (!rgba) {
            t[a][b][0] = (float)((const uint16_t*)level_texels)[idx] / 65535.0f;
            t[a][b][1] = t[a][b][2] = t[a][b][3] = 0.0f;
        } else {
            const uint8_t* q = (const uint8_t*)level_texels + idx * 4;
            t[a][b][0] = srgb_lut[q[0]]; t[a][b][1] = srgb_lut[q[1]]; t[a][b][2] = srgb_lut[q[2]];
            t[a][b][3] = (float)q[3] / 255.0f;
        }
    }
    for (int c = 0; c < (rgba ? 4 : 1); ++c) {
        float l0 = t[0][0][c] + (t[0][1][c] - t[0][0][c]) * ry;
        float l1 = t[1][0][c] + (t[1][1][c] - t[1][0][c]) * ry;
        out[c] = l0 + (l1 - l0) * rx;
    }
}

typedef struct filter_ctx {
    const void* mip0;        /* the atlas tile (layer) */
    const void* mips;        /* its appended levels 1.. (NULL when mip_level_count == 1) */
    int rgba;
    uint32_t T, mip_level_count;
    int trilinear;           /* 0: plain bilinear at mip 0 (ALBEDO without FILTER) */
} filter_ctx;

static void sample_filtered(const filter_ctx* fc, float u, float v, float step_u, float step_v, float out[4]) {
    const int nch = fc->rgba ? 4 : 1;
    if (!fc->trilinear) { bilinear_level(fc->mip0, fc->rgba, fc->T, u, v, out); return; }
    float Px = step_u * (float)fc->T, Py = step_v * (float)fc->T;
    float Pmax = fmaxf(Px, Py), Pmin = fminf(Px, Py);
    float ratio = ceilf(Pmax / Pmin);
    uint32_t N = (ratio >= 16.0f) ? 16u : (ratio >= 1.0f ? (uint32_t)ratio : 1u);   /* NaN (0 / 0) -> 1 */
    float lambda = bto_log2f(Pmax / (float)N);
    float d = fminf(fmaxf(lambda, 0.0f), (float)(fc->mip_level_count - 1u));          /* fmaxf(NaN, 0) = 0 */
    uint32_t lo = (uint32_t)floorf(d), hi = lo + 1u < fc->mip_level_count ? lo + 1u : lo;
    float w = d - (float)lo;
    const int x_major = Px >= Py;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (uint32_t i = 0; i < N; ++i) {
        float t = ((float)i + 0.5f) / (float)N - 0.5f;
        float uu = x_major ? u + t * step_u : u, vv = x_major ? v : v + t * step_v;
        float a[4], b[4];
        const uint8_t* base = (const uint8_t*)fc->mips;
        const size_t px = fc->rgba ? 4 : 2;
        bilinear_level(lo == 0 ? fc->mip0 : (const void*)(base + mip_offset_texels(fc->T, lo) * px), fc->rgba, fc->T >> lo, uu, vv, a);
        if (hi != lo) bilinear_level((const void*)(base + mip_offset_texels(fc->T, hi) * px), fc->rgba, fc->T >> hi, uu, vv, b);
        for (int c = 0; c < nch; ++c) acc[c] += (hi != lo) ? a[c] * (1.0f - w) + b[c] * w : a[c];
    }
    for (int c = 0; c < nch; ++c) out[c] = acc[c] / (float)N;
}

static size_t mip_chain_texels(uint32_t T, uint32_t levels) { return mip_offset_texels(T, levels); }

/* filter context of one resolved atlas tile of attachment `a` (0 = height R16, 1 = albedo Rgba8) */
static filter_ctx make_filter_ctx(const bto_vertex_inputs* in, const bto_filter_inputs* fi, int albedo, uint32_t atlas_index) {
    filter_ctx fc;
    uint32_t T = in->texture_size;
    if (in->atlas_size && atlas_index > in->atlas_size - 1u) atlas_index = in->atlas_size - 1u;
    const size_t chain = mip_chain_texels(T, fi->mip_level_count);
    fc.rgba = albedo;
    fc.T = T;
    fc.mip_level_count = fi->mip_level_count;
    fc.trilinear = (in->flags & BT_FLAG_FILTER_TRILINEAR_ANISO) != 0;
    if (!albedo) {
        fc.mip0 = in->height_atlas + (size_t)atlas_index * T * T;
        fc.mips = fi->height_mips ? (const void*)(fi->height_mips + (size_t)atlas_index * chain) : NULL;
    } else {
        fc.mip0 = fi->albedo_atlas + (size_t)atlas_index * T * T * 4;
        fc.mips = fi->albedo_mips ? (const void*)(fi->albedo_mips + (size_t)atlas_index * chain * 4) : NULL;
    }
    return fc;
}

/* the footprint of one grid step in the uv space of the resolved atlas tile (per axis; the grid is square) */
static float grid_step_uv(const bto_vertex_inputs* in, uint32_t tile_lod, uint32_t atlas_lod, uint32_t attachment) {
    float inv_g = 1.0f / in->view_config.grid_size;
    float rel = ldexpf(1.0f, (int32_t)atlas_lod - (int32_t)tile_lod);
    return (inv_g * rel) * in->attachments[attachment].scale;
}

static void gen_vertex_filtered(const bto_vertex_inputs* in, const bto_filter_inputs* fi, uint64_t vi, bt_vertex* out, uint32_t* albedo_out);

static void gen_vertex(const bto_vertex_inputs* in, uint64_t vi, bt_vertex* out) {
    const bt_view_config_uniform* vc = &in->view_config;
    const bt_model_approximation* ap = &in->approximation;
    int sph = in->spherical;
    /* vertex_info — vertex.wgsl:30-71 */
    uint32_t tile_index = (uint32_t)(vi / vc->vertices_per_tile);
    bt_tile tile = in->tiles[tile_index];
    /* compute_tile_uv — functions.wgsl:64-71 */
    uint32_t grid_index = (uint32_t)(vi % vc->vertices_per_tile);
    uint32_t r = grid_index % vc->vertices_per_row;
    uint32_t lo = 1u, hi = vc->vertices_per_row - 2u;
    uint32_t row_index = (r < lo ? lo : (r > hi ? hi : r)) - 1u;
    uint32_t column_index = grid_index / vc->vertices_per_row;
    coord c0 = {tile.side, tile.lod, {tile.x, tile.y},
                {(float)(column_index + (row_index & 1u)) / vc->grid_size, (float)(row_index >> 1) / vc->grid_size}};
    float lp[3], wp[3], wn[3], q[3];
    local_position(sph, &c0, lp);
    position_local_to_world(&in->mesh, lp, wp);
    normal_local_to_world(&in->mesh, sph, lp, wn);
    for (int k = 0; k < 3; ++k) q[k] = (wp[k] + ap->approximate_height * wn[k]) - in->view_world_position[k];
    float d = o_length(q);
    int high_precision = (in->flags & BT_FLAG_HIGH_PRECISION) && d < vc->precision_threshold_distance;

    if (high_precision) {
        float rel[3];
        relative_position(ap, c0, rel);
        for (int k = 0; k < 3; ++k) q[k] = rel[k] + ap->approximate_height * wn[k];
        d = o_length(q);
    }
    /* compute_morph — functions.wgsl:35-49 */
    coord c = c0;
    if (in->flags & BT_FLAG_MORPH) {
        float target_lod = bto_log2f(2.0f * vc->morph_distance / d);
        float ratio = (c0.lod == 0u) ? 0.0f
            : o_inverse_mix((float)c0.lod + vc->morph_range, (float)c0.lod, target_lod);
        for (int k = 0; k < 2; ++k) {
            float even = (float)(bto_f32_to_u32(c0.uv[k] * vc->grid_size) & ~1u) / vc->grid_size;
            c.uv[k] = o_mix(c0.uv[k], even, ratio);
        }
    }
    if (high_precision) {
        float rel[3];
        relative_position(ap, c, rel);
        for (int k = 0; k < 3; ++k) wp[k] = in->view_world_position[k] + rel[k];
        /* world_normal = approximate_world_normal */
    } else {
        local_position(sph, &c, lp);
        position_local_to_world(&in->mesh, lp, wp);
        normal_local_to_world(&in->mesh, sph, lp, wn);
    }
    /* compute_blend — functions.wgsl:51-62 */
    float target_lod = fminf(bto_log2f(vc->blend_distance / d), (float)in->terrain_config.lod_count - 0.00001f);
    uint32_t blend_lod = bto_f32_to_u32(target_lod);
    float blend_ratio = 0.0f;
    if ((in->flags & BT_FLAG_BLEND) && blend_lod != 0u)
        blend_ratio = o_inverse_mix((float)blend_lod + vc->blend_range, (float)blend_lod, target_lod);

    /* vertex — vertex.wgsl:85-98 */
    atlas_tile ta = lookup_tile(in, c, blend_lod, 0u);
    float height = sample_height(in, &ta);
    float nrm[3] = {wn[0], wn[1], wn[2]};
    if (in->flags & BT_FLAG_NORMALS) sample_normal(in, &ta, wn, nrm);
    if (blend_ratio > 0.0f) {
        atlas_tile tb = lookup_tile(in, c, blend_lod, 1u);
        height = o_mix(height, sample_height(in, &tb), blend_ratio);
        if (in->flags & BT_FLAG_NORMALS) { /* fragment.wgsl:103-107 */
            float n2[3];
            sample_normal(in, &tb, wn, n2);
            for (int k = 0; k < 3; ++k) nrm[k] = o_mix(nrm[k], n2[k], blend_ratio);
        }
    }
    /* vertex_output — vertex.wgsl:73-83 */
    for (int k = 0; k < 3; ++k) out->position[k] = wp[k] + height * wn[k];
    out->position[3] = 1.0f;
    out->normal[0] = nrm[0]; out->normal[1] = nrm[1]; out->normal[2] = nrm[2];
    out->uv[0] = c.uv[0]; out->uv[1] = c.uv[1];
    out->tile_index = tile_index;
}

/* gen_vertex with the configs[2] sampler: identical vertex_info / morph / blend / lookup; every height tap goes through
 * sample_filtered (FILTER flag) and the albedo attachment is sampled at the same point (ALBEDO flag) */
static void gen_vertex_filtered(const bto_vertex_inputs* in, const bto_filter_inputs* fi, uint64_t vi, bt_vertex* out, uint32_t* albedo_out) {
    const bt_view_config_uniform* vc = &in->view_config;
    const bt_model_approximation* ap = &in->approximation;
    int sph = in->spherical;
    uint32_t tile_index = (uint32_t)(vi / vc->vertices_per_tile);
    bt_tile tile = in->tiles[tile_index];
    uint32_t grid_index = (uint32_t)(vi % vc->vertices_per_tile);
    uint32_t r = grid_index % vc->vertices_per_row;
    uint32_t lo = 1u, hi = vc->vertices_per_row - 2u;
    uint32_t row_index = (r < lo ? lo : (r > hi ? hi : r)) - 1u;
    uint32_t column_index = grid_index / vc->vertices_per_row;
    coord c0 = {tile.side, tile.lod, {tile.x, tile.y},
                {(float)(column_index + (row_index & 1u)) / vc->grid_size, (float)(row_index >> 1) / vc->grid_size}};
    float lp[3], wp[3], wn[3], q[3];
    local_position(sph, &c0, lp);
    position_local_to_world(&in->mesh, lp, wp);
    normal_local_to_world(&in->mesh, sph, lp, wn);
    for (int k = 0; k < 3; ++k) q[k] = (wp[k] + ap->approximate_height * wn[k]) - in->view_world_position[k];
    float d = o_length(q);
    int high_precision = (in->flags & BT_FLAG_HIGH_PRECISION) && d < vc->precision_threshold_distance;
    if (high_precision) {
        float rel[3];
        relative_position(ap, c0, rel);
        for (int k = 0; k < 3; ++k) q[k] = rel[k] + ap->approximate_height * wn[k];
        d = o_length(q);
    }
    coord c = c0;
    if (in->flags & BT_FLAG_MORPH) {
        float target_lod = bto_log2f(2.0f * vc->morph_distance / d);
        float ratio = (c0.lod == 0u) ? 0.0f : o_inverse_mix((float)c0.lod + vc->morph_range, (float)c0.lod, target_lod);
        for (int k = 0; k < 2; ++k) {
            float even = (float)(bto_f32_to_u32(c0.uv[k] * vc->grid_size) & ~1u) / vc->grid_size;
            c.uv[k] = o_mix(c0.uv[k], even, ratio);
        }
    }
    if (high_precision) {
        float rel[3];
        relative_position(ap, c, rel);
        for (int k = 0; k < 3; ++k) wp[k] = in->view_world_position[k] + rel[k];
    } else {
        local_position(sph, &c, lp);
        position_local_to_world(&in->mesh, lp, wp);
        normal_local_to_world(&in->mesh, sph, lp, wn);
    }
    float target_lod = fminf(bto_log2f(vc->blend_distance / d), (float)in->terrain_config.lod_count - 0.00001f);
    uint32_t blend_lod = bto_f32_to_u32(target_lod);
    float blend_ratio = 0.0f;
    if ((in->flags & BT_FLAG_BLEND) && blend_lod != 0u)
        blend_ratio = o_inverse_mix((float)blend_lod + vc->blend_range, (float)blend_lod, target_lod);

    float height = 0.0f, nrm[3] = {wn[0], wn[1], wn[2]}, color[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (uint32_t pass = 0; pass < (blend_ratio > 0.0f ? 2u : 1u); ++pass) {
        atlas_tile t = lookup_tile(in, c, blend_lod, pass);
        const uint32_t aa = fi->albedo_attachment;
        float u = t.c.uv[0] * in->attachments[0].scale + in->attachments[0].offset;
        float v = t.c.uv[1] * in->attachments[0].scale + in->attachments[0].offset;
        float step = grid_step_uv(in, tile.lod, t.c.lod, 0);
        filter_ctx fh = make_filter_ctx(in, fi, 0, t.index);
        float s[4];
        sample_filtered(&fh, u, v, step, step, s);
        float h = o_mix(in->terrain_config.min_height, in->terrain_config.max_height, s[0]);
        float n2[3] = {wn[0], wn[1], wn[2]};
        if (in->flags & BT_FLAG_NORMALS) {   /* sample_normal with the filtered sampler for its four taps */
            float side_length = in->spherical ? (float)(3.14159265359 / 4.0) * in->terrain_config.scale : in->terrain_config.scale;
            float pixels_per_side = in->attachments[0].size * o_tile_count(t.c.lod);
            float dbs = side_length / pixels_per_side;
            float offset = 0.5f / in->attachments[0].size;
            float taps[4];
            const float du[4] = {-offset, 0.0f, offset, 0.0f}, dv[4] = {0.0f, -offset, 0.0f, offset};   /* left, up, right, down */
            for (int k = 0; k < 4; ++k) {
                sample_filtered(&fh, u + du[k], v + dv[k], step, step, s);
                taps[k] = o_mix(in->terrain_config.min_height, in->terrain_config.max_height, s[0]);
            }
            float sv[3] = {taps[0] - taps[2], taps[3] - taps[1], dbs}, sn[3], rr[3];
            o_normalize(sv, sn);
            if (in->spherical) {
                static const float FACE_UP[6][3] = {{0, 1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, -1}, {-1, 0, 0}, {-1, 0, 0}};
                float n[3], tg[3], bt[3];
                o_normalize(wn, n);
                o_cross(FACE_UP[t.c.side], n, tg);
                o_cross(n, tg, bt);
                for (int k = 0; k < 3; ++k) rr[k] = (tg[k] * sn[0] + bt[k] * sn[1]) + n[k] * sn[2];
            } else {
                rr[0] = sn[0]; rr[1] = sn[2]; rr[2] = sn[1];
            }
            o_normalize(rr, n2);
        }
        float col[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (in->flags & BT_FLAG_ALBEDO) {
            float ua = t.c.uv[0] * in->attachments[aa].scale + in->attachments[aa].offset;
            float va = t.c.uv[1] * in->attachments[aa].scale + in->attachments[aa].offset;
            float step_a = grid_step_uv(in, tile.lod, t.c.lod, aa);
            filter_ctx fa = make_filter_ctx(in, fi, 1, t.index);
            sample_filtered(&fa, ua, va, step_a, step_a, col);
        }
        if (pass == 0) {
            height = h;
            for (int k = 0; k < 3; ++k) nrm[k] = n2[k];
            for (int k = 0; k < 4; ++k) color[k] = col[k];
        } else {
            height = o_mix(height, h, blend_ratio);
            if (in->flags & BT_FLAG_NORMALS) for (int k = 0; k < 3; ++k) nrm[k] = o_mix(nrm[k], n2[k], blend_ratio);
            for (int k = 0; k < 4; ++k) color[k] = o_mix(color[k], col[k], blend_ratio);
        }
    }
    for (int k = 0; k < 3; ++k) out->position[k] = wp[k] + height * wn[k];
    out->position[3] = 1.0f;
    out->normal[0] = nrm[0]; out->normal[1] = nrm[1]; out->normal[2] = nrm[2];
    out->uv[0] = c.uv[0]; out->uv[1] = c.uv[1];
    out->tile_index = tile_index;
    if (albedo_out) {   /* pack4x8unorm of the linear colour */
        uint32_t packed = 0;
        for (int k = 0; k < 4; ++k) packed |= (uint32_t)floorf(0.5f + 255.0f * o_saturate(color[k])) << (8 * k);
        *albedo_out = packed;
    }
}

void bto_gen_vertices_filtered(const bto_vertex_inputs* in, const bto_filter_inputs* fi, uint64_t first, uint64_t count, bt_vertex* out,
                               uint32_t* albedo_out, int threads) {
    srgb_lut_init();
    (void)threads;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static, 4096) num_threads(threads)
#endif
    for (int64_t i = 0; i < (int64_t)count; ++i) gen_vertex_filtered(in, fi, first + (uint64_t)i, &out[i], albedo_out ? &albedo_out[i] : NULL);
}

void bto_gen_vertices(const bto_vertex_inputs* in, uint64_t first, uint64_t count, bt_vertex* out, int threads) {
    (void)threads;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static, 4096) num_threads(threads)
#endif
    for (int64_t i = 0; i < (int64_t)count; ++i) gen_vertex(in, first + (uint64_t)i, &out[i]);
}

int bto_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

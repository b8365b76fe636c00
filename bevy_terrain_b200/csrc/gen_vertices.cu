// gen_vertices.cu — per-tile heightmap-sampled vertex-grid generation (positions + normals).
//
// Replaces the vertex stage of the reference's indirect draw (src/shaders/render/vertex.wgsl:85-98,
// driven by DrawTerrainCommand, src/render/terrain_view_bind_group.rs:274-298) and fuses the
// height-derived normal of the fragment stage (sample_normal, src/shaders/attachments.wgsl:51-107,
// blended as src/shaders/render/fragment.wgsl:99-107).
//
// Design (B200):
//   * persistent CTAs (one per SM, warp-specialised: 19 compute warps + 1 TMA store warp) stride over work items
//     = (tile, block of CB column strips); image hand-over by mbarriers, no CTA-wide barrier in steady state
//   * the reference's triangle-strip index stream visits every grid vertex ~2x (8448 strip
//     vertices for 65x65 = 4225 unique ones at grid 64, compute_tile_uv functions.wgsl:64-71);
//     each unique vertex is computed ONCE and scattered to its 2-3 strip positions in a
//     shared-memory image of the output block — bit-identical to recomputation because the
//     vertex is a pure function of (tile, column, row)
//   * the image is written to HBM with TMA 1-D bulk stores (cp.async.bulk.global.shared::cta,
//     one per strip), double-buffered: the stores of item i drain while item i+1 is computed;
//     HBM sees only full-line, fully coalesced writes of the 40-B interleaved records
//   * the uniform block is staged into shared memory with one TMA bulk load + mbarrier
//   * height texels are gathered with read-only (ld.global.nc) 16-bit loads; lanes walk along
//     texel rows so one request touches 1-3 cache lines; exact-f32 bilinear weights (no texture
//     unit: its fixed-point weights are not reproducible)
#include "bt_device.cuh"
#include "bt_kernels.h"

namespace btd {

#ifndef BT_GEN_COMPUTE_WARPS
#define BT_GEN_COMPUTE_WARPS 19
#endif
#ifndef BT_GEN_ROTATION
#define BT_GEN_ROTATION 11
#endif
constexpr uint32_t GEN_ROTATION = BT_GEN_ROTATION;
constexpr int GEN_COMPUTE_WARPS = BT_GEN_COMPUTE_WARPS;  // compute warps per CTA (+ 1 store warp), one CTA per SM
constexpr int GEN_THREADS = (GEN_COMPUTE_WARPS + 1) * 32;
constexpr int GEN_CTAS_PER_SM = 1;
constexpr int GEN_BUFFERS = 2;         // output images per CTA: compute of item i+1 overlaps the TMA stores of item i
// Strips are laid out in the smem image with a stride that is 16 (mod 128) bytes: lanes of a warp write records of 32 different
// strips, and with that stride each quarter-warp of a 128-bit store (and each half-warp of a 64-bit one) covers all 32 banks
// once (the 40-byte records themselves are 8-byte aligned; the stride keeps every strip 16-byte aligned for the TMA).
__host__ __device__ constexpr uint32_t strip_stride_for(uint32_t strip_bytes) { return strip_bytes + (144u - strip_bytes % 128u) % 128u; }

struct GenParams {
    const ViewUniforms* uniforms;
    const bt_tile* tiles;
    const bt_indirect* indirect;
    const bt_tile_tree_entry* tile_tree;
    const uint16_t* height_atlas;
    bt_vertex* vertices;
    uint64_t vertex_capacity;
    GenStatus* status;
    uint32_t flags;
    uint32_t grid_size;        // g
    uint32_t strips_per_item;  // CB
    uint32_t items_per_tile;   // ceil(g / CB)
    uint32_t image_bytes;      // CB * strip_stride, per buffer
    uint32_t strip_stride;     // strip_stride_for(vertices_per_row * sizeof(bt_vertex))
    uint32_t magic_items;      // ceil(2^32 / items_per_tile): w / items_per_tile == umulhi(w, magic_items) (w < 2^32 / items_per_tile)
    uint32_t magic_cols;       // ceil(2^32 / (CB + 1)) for full-width items
    uint32_t magic_cols_ragged;  // the same for the narrower last item of a tile (g % CB strips), if any
};

struct Record {  // one bt_vertex as five 8-byte words
    float2 w[5];
};

// Loop-invariant values of one launch, hoisted into registers once per thread.  Every entry is the
// result of exactly the operations the per-vertex formula would perform, so hoisting is bit-neutral.
struct FrameConsts {
    float g, inv_g;            // grid_size and, when it is a power of two, its exact reciprocal
    bool  g_pow2;
    float Tf;                  // f32(texture_size)
    int   T, Tm1;
    uint32_t interior;         // texel (i, j) with i-1, j-1 < interior has its 4x4 neighbourhood inside the texture
    float min_h, max_h;
    float a_scale, a_offset;   // attachment_uv (attachments.wgsl:7-10)
    float normal_offset;       // 0.5 / attachments[0].size (attachments.wgsl:85)
    float side_length;         // attachments.wgsl:71,78
    float size0;               // attachments[0].size
    Vec3  planar_normal;       // normal_local_to_world for the planar model: a constant (functions.wgsl:14-24)
    float lod_cap;             // f32(lod_count) - 0.00001 (functions.wgsl:52)
    uint32_t tree_size, tree_mask, lod_count;  // tree_mask = tree_size - 1 when tree_size is a power of two, else 0
    uint32_t last_layer;       // atlas_size - 1: texture-array indices clamp to it
};

template <bool SPHERICAL>
__device__ __forceinline__ FrameConsts make_frame_consts(const ViewUniforms& U) {
    FrameConsts F;
    F.g = U.view_config.grid_size;
    const uint32_t gi = (uint32_t)F.g;
    F.g_pow2 = ((float)gi == F.g) && gi != 0u && (gi & (gi - 1u)) == 0u;
    F.inv_g = 1.0f / F.g;      // exact when g is a power of two: x * inv_g == x / g bit for bit
    F.T = (int)U.texture_size;
    F.Tm1 = F.T - 1;
    F.interior = F.T >= 4 ? (uint32_t)(F.T - 3) : 0u;   // 1 <= i <= T-3  <=>  i-1 >= 0 and i+2 <= T-1
    F.Tf = (float)U.texture_size;
    F.min_h = U.terrain_config.min_height;
    F.max_h = U.terrain_config.max_height;
    F.a_scale = U.attachments[0].scale;
    F.a_offset = U.attachments[0].offset;
    F.size0 = U.attachments[0].size;
    F.normal_offset = 0.5f / F.size0;
    F.side_length = SPHERICAL ? (float)(3.14159265359 / 4.0) * U.terrain_config.scale : U.terrain_config.scale;
    F.planar_normal = normal_local_to_world<false>(U.mesh, Vec3{0.0f, 0.0f, 0.0f});
    F.lod_cap = (float)U.terrain_config.lod_count - 0.00001f;
    F.tree_size = U.view_config.tree_size;
    F.tree_mask = (F.tree_size & (F.tree_size - 1u)) == 0u ? F.tree_size - 1u : 0u;
    F.lod_count = U.terrain_config.lod_count;
    F.last_layer = U.atlas_size ? U.atlas_size - 1u : 0u;
    return F;
}

// bilinear R16 fetch — exact f32 weights, texel centres at i+0.5, clamp-to-edge; operation order
// of the CPU twin AttachmentData::sample (src/terrain_data/mod.rs:221-264): lerp y, then x
__device__ __forceinline__ float sample_r16(const FrameConsts& F, const uint16_t* __restrict__ layer, float u, float v) {
    float px = u * F.Tf - 0.5f, py = v * F.Tf - 0.5f;
    float fx = floorf(px), fy = floorf(py);
    float rx = px - fx, ry = py - fy;
    int ix = (int)fx, iy = (int)fy;
    int x0 = min(max(ix, 0), F.Tm1), x1 = min(max(ix + 1, 0), F.Tm1);
    int y0 = min(max(iy, 0), F.Tm1), y1 = min(max(iy + 1, 0), F.Tm1);
    const uint16_t* r0 = layer + y0 * F.T;
    const uint16_t* r1 = layer + y1 * F.T;
    float v00 = unorm16_to_float(__ldg(r0 + x0));
    float v01 = unorm16_to_float(__ldg(r1 + x0));
    float v10 = unorm16_to_float(__ldg(r0 + x1));
    float v11 = unorm16_to_float(__ldg(r1 + x1));
    float a = v00 + (v01 - v00) * ry;
    float b = v10 + (v11 - v10) * ry;
    return a + (b - a) * rx;
}

struct AtlasTile {
    uint32_t index;
    Coord c;
};

// lookup_tile (default variant) — functions.wgsl:232-246,198-206
__device__ __forceinline__ AtlasTile lookup_tile(const FrameConsts& F, const bt_tile_tree_entry* __restrict__ tile_tree,
                                                 Coord c, uint32_t lod) {
    change_lod(c, lod);
    const uint32_t ts = F.tree_size;
    const uint32_t tx = F.tree_mask ? (c.x & F.tree_mask) : c.x % ts;   // xy % tree_size (functions.wgsl:199)
    const uint32_t ty = F.tree_mask ? (c.y & F.tree_mask) : c.y % ts;
    const uint32_t idx = ((c.side * F.lod_count + c.lod) * ts + tx) * ts + ty;
    const uint2 e = __ldg(reinterpret_cast<const uint2*>(tile_tree) + idx);
    change_lod(c, e.y);
    return AtlasTile{e.x, c};
}

// sample_height — attachments.wgsl:7-24,45-49
__device__ __forceinline__ float sample_height_uv(const FrameConsts& F, const uint16_t* __restrict__ layer, float u, float v) {
    const float s = sample_r16(F, layer, u, v);
    return mix(F.min_h, F.max_h, s);
}

// The five taps of sample_height + sample_normal (centre, left, up, right, down) land in a "plus" of at most
// 12 distinct texels: rows iy..iy+1 x columns ix-1..ix+2 and rows iy-1..iy+2 x columns ix..ix+1.  This routine
// loads and decodes each texel once and evaluates every tap with exactly the per-tap arithmetic of sample_r16
// (same operands, same order), so it is bit-identical to five independent taps.  `v + 0.0f == v` and
// `u + 0.0f == u` hold bit for bit because attachment_uv yields values >= +0.
struct FiveTaps { float centre, left, up, right, down; bool ok; };

__device__ __forceinline__ float lerp2(float v00, float v01, float v10, float v11, float rx, float ry) {
    const float a = v00 + (v01 - v00) * ry;
    const float b = v10 + (v11 - v10) * ry;
    return a + (b - a) * rx;
}

__device__ __forceinline__ FiveTaps sample_five(const FrameConsts& F, const uint16_t* __restrict__ layer, float u, float v) {
    const float o = F.normal_offset;
    // tap coordinates: (centre x, centre y), (left x, right x), (up y, down y) as packed pairs; each product is followed by a
    // scalar "- 0.5" (same two roundings per value as sample_r16)
    const float2 Tf2 = make_float2(F.Tf, F.Tf);
    const float2 pc = __fmul2_rn(make_float2(u, v), Tf2);
    const float2 plr = __fmul2_rn(__fadd2_rn(make_float2(u, u), make_float2(-o, o)), Tf2);
    const float2 pud = __fmul2_rn(__fadd2_rn(make_float2(v, v), make_float2(-o, o)), Tf2);
    const float pxc = pc.x - 0.5f, pyc = pc.y - 0.5f, pxl = plr.x - 0.5f, pxr = plr.y - 0.5f, pyu = pud.x - 0.5f, pyd = pud.y - 0.5f;
    const float fxc = floorf(pxc), fyc = floorf(pyc), fxl = floorf(pxl), fxr = floorf(pxr), fyu = floorf(pyu), fyd = floorf(pyd);
    const float2 rc = __fadd2_rn(make_float2(pxc, pyc), make_float2(-fxc, -fyc));
    const float2 rlr = __fadd2_rn(make_float2(pxl, pxr), make_float2(-fxl, -fxr));
    const float2 rud = __fadd2_rn(make_float2(pyu, pyd), make_float2(-fyu, -fyd));
    const float rxc = rc.x, ryc = rc.y, rxl = rlr.x, rxr = rlr.y, ryu = rud.x, ryd = rud.y;
    const int ix = (int)fxc, iy = (int)fyc;
    const uint32_t jl = (uint32_t)((int)fxl - ix + 1), jr = (uint32_t)((int)fxr - ix);  // expected jl in {0,1}, jr in {0,1}
    const uint32_t ju = (uint32_t)((int)fyu - iy + 1), jd = (uint32_t)((int)fyd - iy);
    FiveTaps out;
    out.ok = (jl | jr | ju | jd) <= 1u;   // offsets of 0.5/size texture units = ~0.5 texel: always true; guarded anyway
    // raw texel r<row><col>
    uint32_t r10, r11, r12, r13, r20, r21, r22, r23, r01, r02, r31, r32;
    if ((uint32_t)(ix - 1) < F.interior && (uint32_t)(iy - 1) < F.interior) {
        // the whole plus lies inside the texture (always the case for uv inside a tile with border_size >= 2):
        // one base address, row pointers by +-T, column offsets as immediates; no clamping
        const uint16_t* p1 = layer + (iy * F.T + ix);
        const uint16_t* p0 = p1 - F.T;
        const uint16_t* p2 = p1 + F.T;
        const uint16_t* p3 = p2 + F.T;
        r10 = __ldg(p1 - 1); r11 = __ldg(p1); r12 = __ldg(p1 + 1); r13 = __ldg(p1 + 2);
        r20 = __ldg(p2 - 1); r21 = __ldg(p2); r22 = __ldg(p2 + 1); r23 = __ldg(p2 + 2);
        r01 = __ldg(p0); r02 = __ldg(p0 + 1);
        r31 = __ldg(p3); r32 = __ldg(p3 + 1);
    } else {
        const int c0 = min(max(ix - 1, 0), F.Tm1), c1 = min(max(ix, 0), F.Tm1), c2 = min(max(ix + 1, 0), F.Tm1), c3 = min(max(ix + 2, 0), F.Tm1);
        const uint16_t* q0 = layer + min(max(iy - 1, 0), F.Tm1) * F.T;
        const uint16_t* q1 = layer + min(max(iy, 0), F.Tm1) * F.T;
        const uint16_t* q2 = layer + min(max(iy + 1, 0), F.Tm1) * F.T;
        const uint16_t* q3 = layer + min(max(iy + 2, 0), F.Tm1) * F.T;
        r10 = __ldg(q1 + c0); r11 = __ldg(q1 + c1); r12 = __ldg(q1 + c2); r13 = __ldg(q1 + c3);
        r20 = __ldg(q2 + c0); r21 = __ldg(q2 + c1); r22 = __ldg(q2 + c2); r23 = __ldg(q2 + c3);
        r01 = __ldg(q0 + c1); r02 = __ldg(q0 + c2);
        r31 = __ldg(q3 + c1); r32 = __ldg(q3 + c2);
    }
    // decode in pairs of adjacent columns of one row — the layout the packed y-lerps consume without re-pairing
    const float2 A01 = unorm16x2_to_float(r10, r11), A23 = unorm16x2_to_float(r12, r13);   // row 1: (t10, t11), (t12, t13)
    const float2 B01 = unorm16x2_to_float(r20, r21), B23 = unorm16x2_to_float(r22, r23);   // row 2
    const float2 E0 = unorm16x2_to_float(r01, r02), E3 = unorm16x2_to_float(r31, r32);     // rows 0 and 3: columns 1, 2
    // y-lerps of rows (1, 2) with the centre's ry:  L<col> = t1c + (t2c - t1c) * ryc  — lerp2's `a` / `b` for centre, left, right.
    // Packed: difference and product for two columns per instruction; the final additions stay scalar (a packed mul feeding a
    // packed add would be contracted into FFMA2 by ptxas, which the numerics contract forbids).
    const float2 ryc2 = make_float2(ryc, ryc);
    const float2 d01 = __fmul2_rn(__fadd2_rn(B01, make_float2(-A01.x, -A01.y)), ryc2);
    const float2 d23 = __fmul2_rn(__fadd2_rn(B23, make_float2(-A23.x, -A23.y)), ryc2);
    const float L0 = A01.x + d01.x, L1 = A01.y + d01.y, L2 = A23.x + d23.x, L3 = A23.y + d23.y;
    // x-stage: a + (b - a) * rx
    out.centre = L1 + (L2 - L1) * rxc;
    // left: columns (ix-1+jl, ix+jl); right: columns (ix+jr, ix+1+jr); rows iy, iy+1 (py == centre's)
    {
        const float la = jl ? L1 : L0, lb = jl ? L2 : L1, ra = jr ? L2 : L1, rb = jr ? L3 : L2;
        const float2 m = __fmul2_rn(__fadd2_rn(make_float2(lb, rb), make_float2(-la, -ra)), make_float2(rxl, rxr));
        out.left = la + m.x;
        out.right = ra + m.y;
    }
    // up: rows (iy-1+ju, iy+ju); down: rows (iy+jd, iy+1+jd); columns ix, ix+1 (px == centre's)
    {
        const float2 M1 = make_float2(A01.y, A23.x), M2 = make_float2(B01.y, B23.x);   // (t11, t12), (t21, t22)
        // (v00, v10) / (v01, v11): up = ju ? rows (1, 2) : rows (0, 1);  down = jd ? rows (2, 3) : rows (1, 2)
        const float2 ulo = ju ? M1 : E0, uhi = ju ? M2 : M1;
        const float2 wlo = jd ? M2 : M1, whi = jd ? E3 : M2;
        const float2 mu = __fmul2_rn(__fadd2_rn(uhi, make_float2(-ulo.x, -ulo.y)), make_float2(ryu, ryu));
        const float2 md = __fmul2_rn(__fadd2_rn(whi, make_float2(-wlo.x, -wlo.y)), make_float2(ryd, ryd));
        const float ua = ulo.x + mu.x, ub = ulo.y + mu.y, da = wlo.x + md.x, db = wlo.y + md.y;
        const float2 m = __fmul2_rn(__fadd2_rn(make_float2(ub, db), make_float2(-ua, -da)), make_float2(rxc, rxc));
        out.up = ua + m.x;
        out.down = da + m.y;
    }
    return out;
}

// sample_normal — attachments.wgsl:51-107 (textureSampleLevel branch); `pre` carries the shared taps
template <bool SPHERICAL>
__device__ __forceinline__ Vec3 sample_normal(const FrameConsts& F, const uint16_t* __restrict__ layer, const AtlasTile& t,
                                              float u, float v, Vec3 vertex_normal, const FiveTaps* pre = nullptr) {
    const float pixels_per_side = F.size0 * tile_count(t.c.lod);
    const float distance_between_samples = div_rn(F.side_length, pixels_per_side);
    const float offset = F.normal_offset;
    float left, up, right, down;
    if (pre && pre->ok) {
        // mix(min_h, max_h, s) = min_h * (1 - s) + max_h * s for two taps at a time; the sum of the two products stays scalar
        const float2 lo = make_float2(F.min_h, F.min_h), hi = make_float2(F.max_h, F.max_h), one = make_float2(1.0f, 1.0f);
        const float2 s_lr = make_float2(pre->left, pre->right), s_ud = make_float2(pre->up, pre->down);
        const float2 a_lr = __fmul2_rn(lo, __fadd2_rn(one, make_float2(-s_lr.x, -s_lr.y))), b_lr = __fmul2_rn(hi, s_lr);
        const float2 a_ud = __fmul2_rn(lo, __fadd2_rn(one, make_float2(-s_ud.x, -s_ud.y))), b_ud = __fmul2_rn(hi, s_ud);
        left = a_lr.x + b_lr.x;
        right = a_lr.y + b_lr.y;
        up = a_ud.x + b_ud.x;
        down = a_ud.y + b_ud.y;
    } else {
        left = sample_height_uv(F, layer, u + -offset, v + 0.0f);
        up = sample_height_uv(F, layer, u + 0.0f, v + -offset);
        right = sample_height_uv(F, layer, u + offset, v + 0.0f);
        down = sample_height_uv(F, layer, u + 0.0f, v + offset);
    }
    const Vec3 sn = normalize3(Vec3{left - right, down - up, distance_between_samples});
    Vec3 r;
    if (SPHERICAL) {
        Vec3 face_up;
        switch (t.c.side) {
            case 0: case 1: face_up = Vec3{0.0f, 1.0f, 0.0f}; break;
            case 2: case 3: face_up = Vec3{0.0f, 0.0f, -1.0f}; break;
            default: face_up = Vec3{-1.0f, 0.0f, 0.0f}; break;
        }
        const Vec3 n = normalize3(vertex_normal);
        const Vec3 tg = cross3(face_up, n);
        const Vec3 bt = cross3(n, tg);
        r = Vec3{(tg.x * sn.x + bt.x * sn.y) + n.x * sn.z, (tg.y * sn.x + bt.y * sn.y) + n.y * sn.z,
                 (tg.z * sn.x + bt.z * sn.y) + n.z * sn.z};
    } else {
        r = Vec3{sn.x, sn.z, sn.y};  // TBN columns (1,0,0),(0,0,1),(0,1,0)
    }
    return normalize3(r);
}

// vertex_info + vertex + vertex_output — vertex.wgsl:30-98
template <bool SPHERICAL>
__device__ __forceinline__ Record compute_vertex(const ViewUniforms& U, const FrameConsts& F, const bt_tile_tree_entry* __restrict__ tile_tree,
                                                 const uint16_t* __restrict__ atlas, uint32_t flags, const bt_tile& tile,
                                                 uint32_t tile_index, uint32_t gx, uint32_t gy) {
    const bt_view_config_uniform& vc = U.view_config;
    const bt_model_approximation& ap = U.approximation;
    const float g = F.g;
    // compute_tile_uv — functions.wgsl:64-71 (x / g == x * (1/g) exactly when g is a power of two)
    const float u0 = F.g_pow2 ? (float)gx * F.inv_g : (float)gx / g;
    const float v0 = F.g_pow2 ? (float)gy * F.inv_g : (float)gy / g;
    const Coord c0{tile.side, tile.lod, tile.x, tile.y, u0, v0};
    Vec3 lp = local_position<SPHERICAL>(c0);
    Vec3 wp = position_local_to_world(U.mesh, lp);
    Vec3 wn = SPHERICAL ? normal_local_to_world<SPHERICAL>(U.mesh, lp) : F.planar_normal;
    const float h0 = ap.approximate_height;
    const float* e = U.view_world_position;
    float d = length3(Vec3{(wp.x + h0 * wn.x) - e[0], (wp.y + h0 * wn.y) - e[1], (wp.z + h0 * wn.z) - e[2]});
    const bool high_precision = (flags & BT_FLAG_HIGH_PRECISION) && d < vc.precision_threshold_distance;
    if (high_precision) {
        const Vec3 rel = relative_position(ap, c0);
        d = length3(Vec3{rel.x + h0 * wn.x, rel.y + h0 * wn.y, rel.z + h0 * wn.z});
    }
    // compute_morph — functions.wgsl:35-49
    Coord c = c0;
    if (flags & BT_FLAG_MORPH) {
        const float target_lod = log2_det(div_rn(2.0f * vc.morph_distance, d));
        const float ratio = (c0.lod == 0u) ? 0.0f : inverse_mix_rn((float)c0.lod + vc.morph_range, (float)c0.lod, target_lod);
        const float even_u = F.g_pow2 ? (float)(f2u(c0.u * g) & ~1u) * F.inv_g : (float)(f2u(c0.u * g) & ~1u) / g;
        const float even_v = F.g_pow2 ? (float)(f2u(c0.v * g) & ~1u) * F.inv_g : (float)(f2u(c0.v * g) & ~1u) / g;
        c.u = mix(c0.u, even_u, ratio);
        c.v = mix(c0.v, even_v, ratio);
    }
    if (high_precision) {
        const Vec3 rel = relative_position(ap, c);
        wp = Vec3{e[0] + rel.x, e[1] + rel.y, e[2] + rel.z};
    } else if (flags & BT_FLAG_MORPH) {
        // without MORPH the coordinate is unchanged and the recomputation would reproduce lp/wp/wn bit for bit
        lp = local_position<SPHERICAL>(c);
        wp = position_local_to_world(U.mesh, lp);
        if (SPHERICAL) wn = normal_local_to_world<SPHERICAL>(U.mesh, lp);
    }
    // compute_blend — functions.wgsl:51-62
    const float target_lod = fminf(log2_det(div_rn(vc.blend_distance, d)), F.lod_cap);
    const uint32_t blend_lod = f2u(target_lod);
    float blend_ratio = 0.0f;
    if ((flags & BT_FLAG_BLEND) && blend_lod != 0u)
        blend_ratio = inverse_mix_rn((float)blend_lod + vc.blend_range, (float)blend_lod, target_lod);

    const int tile_texels = F.T * F.T;
    const bool normals = (flags & BT_FLAG_NORMALS) != 0u;
    const AtlasTile ta = lookup_tile(F, tile_tree, c, blend_lod);
    // An INVALID entry (0xFFFFFFFF: no tile loaded yet for this slot) samples like the reference does: WGSL clamps the
    // texture-array index to the last layer (textureSampleLevel array_index), so the frame keeps rendering.
    const uint16_t* layer_a = atlas + (size_t)min(ta.index, F.last_layer) * tile_texels;
    const float ua = ta.c.u * F.a_scale + F.a_offset, va = ta.c.v * F.a_scale + F.a_offset;  // attachment_uv
    float height;
    Vec3 nrm = wn;
    if (normals) {
        const FiveTaps taps = sample_five(F, layer_a, ua, va);
        height = mix(F.min_h, F.max_h, taps.centre);
        nrm = sample_normal<SPHERICAL>(F, layer_a, ta, ua, va, wn, &taps);
    } else {
        height = sample_height_uv(F, layer_a, ua, va);
    }
    if (blend_ratio > 0.0f) {
        const AtlasTile tb = lookup_tile(F, tile_tree, c, blend_lod - 1u);
        const uint16_t* layer_b = atlas + (size_t)min(tb.index, F.last_layer) * tile_texels;
        const float ub = tb.c.u * F.a_scale + F.a_offset, vb = tb.c.v * F.a_scale + F.a_offset;
        if (normals) {  // fragment.wgsl:103-107
            const FiveTaps taps = sample_five(F, layer_b, ub, vb);
            height = mix(height, mix(F.min_h, F.max_h, taps.centre), blend_ratio);
            const Vec3 n2 = sample_normal<SPHERICAL>(F, layer_b, tb, ub, vb, wn, &taps);
            nrm = Vec3{mix(nrm.x, n2.x, blend_ratio), mix(nrm.y, n2.y, blend_ratio), mix(nrm.z, n2.z, blend_ratio)};
        } else {
            height = mix(height, sample_height_uv(F, layer_b, ub, vb), blend_ratio);
        }
    }
    Record r;  // vertex_output — vertex.wgsl:73-83
    const float2 hn = __fmul2_rn(make_float2(height, height), make_float2(wn.x, wn.y));
    r.w[0] = make_float2(wp.x + hn.x, wp.y + hn.y);
    r.w[1] = make_float2(wp.z + height * wn.z, 1.0f);
    r.w[2] = make_float2(nrm.x, nrm.y);
    r.w[3] = make_float2(nrm.z, c.u);
    r.w[4] = make_float2(c.v, __uint_as_float(tile_index));
    return r;
}

// Record k of a strip starts at 40*k bytes: 16-byte aligned for even k, 8 (mod 16) for odd k — the parity is static at every
// call site, so each copy is two 128-bit stores and one 64-bit store.
template <int ODD>
__device__ __forceinline__ void store_record(unsigned char* image, uint32_t strip_stride, uint32_t strip, uint32_t k, const Record& r) {
    unsigned char* dst = image + strip * strip_stride + k * (uint32_t)sizeof(bt_vertex);
    if (ODD) {
        *reinterpret_cast<float2*>(dst) = r.w[0];
        *reinterpret_cast<float4*>(dst + 8) = make_float4(r.w[1].x, r.w[1].y, r.w[2].x, r.w[2].y);
        *reinterpret_cast<float4*>(dst + 24) = make_float4(r.w[3].x, r.w[3].y, r.w[4].x, r.w[4].y);
    } else {
        *reinterpret_cast<float4*>(dst) = make_float4(r.w[0].x, r.w[0].y, r.w[1].x, r.w[1].y);
        *reinterpret_cast<float4*>(dst + 16) = make_float4(r.w[2].x, r.w[2].y, r.w[3].x, r.w[3].y);
        *reinterpret_cast<float2*>(dst + 32) = r.w[4];
    }
}

struct GenSmemHeader {
    ViewUniforms U;
    uint64_t bar;                   // uniform-block TMA load
    uint64_t full[GEN_BUFFERS];     // image b completely written by the compute warps (count = compute warps)
    uint64_t empty[GEN_BUFFERS];    // image b completely read by the TMA stores (count = 1, the store warp)
};

// Warp-specialised pipeline, one CTA per SM:
//   * GEN_COMPUTE_WARPS warps compute vertices and scatter them into output image (it & 1); there is NO CTA-wide barrier
//     on the steady-state path: a warp that has finished its chunks of item i arrives on full[b] and starts on its
//     chunks of item i+1 in the other image, so the ragged last round of an item (1105 vertices = 34.5 warp-chunks) is
//     filled with the next item's work
//   * one store warp waits on full[b], issues the item's TMA bulk stores, waits until they have READ the image
//     (cp.async.bulk.wait_group.read) and arrives on empty[b], which the compute warps wait on before reusing image b
// FLAGS_CT >= 0: the flag word is a compile-time constant (the two common sets get their own instantiation, so the flag
// branches and the code of the disabled features disappear); FLAGS_CT < 0: generic kernel, flags read from the parameters.
template <bool SPHERICAL, int FLAGS_CT>
__global__ void __launch_bounds__(GEN_THREADS, 1) gen_vertices_kernel(const GenParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    GenSmemHeader& H = *reinterpret_cast<GenSmemHeader*>(smem_raw);
    unsigned char* images = smem_raw + ((sizeof(GenSmemHeader) + 127) & ~size_t(127));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&H.bar, 1);
        for (int b = 0; b < GEN_BUFFERS; ++b) { mbar_init(&H.full[b], GEN_COMPUTE_WARPS); mbar_init(&H.empty[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&H.bar, (uint32_t)sizeof(ViewUniforms));
        bulk_load(&H.U, P.uniforms, (uint32_t)sizeof(ViewUniforms), &H.bar);
    }
    mbar_wait_cta(&H.bar, 0);
    const ViewUniforms& U = H.U;

    const uint32_t g = P.grid_size;
    const uint32_t vpr = U.view_config.vertices_per_row, vpt = U.view_config.vertices_per_tile;
    const uint32_t strip_bytes = vpr * (uint32_t)sizeof(bt_vertex);
    const uint32_t strip_stride = P.strip_stride;
    uint32_t n_tiles = __ldg(&P.indirect->vertex_count) / vpt;
    const uint64_t cap_tiles = P.vertex_capacity / vpt;
    if (n_tiles > cap_tiles) {
        if (blockIdx.x == 0 && tid == 0) P.status->overflow = 1u;
        n_tiles = (uint32_t)cap_tiles;
    } else if (blockIdx.x == 0 && tid == 0) {
        P.status->overflow = 0u;
    }
    if (blockIdx.x == 0 && tid == 0) {
        const uint64_t n = (uint64_t)n_tiles * vpt;
        P.status->vertex_count_lo = (uint32_t)n;
        P.status->vertex_count_hi = (uint32_t)(n >> 32);
    }
    const uint32_t n_items = n_tiles * P.items_per_tile;

    if (warp == GEN_COMPUTE_WARPS) {
        // ---------------- store warp ----------------
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
                const uint32_t b = it % GEN_BUFFERS, use = it / GEN_BUFFERS;
                const uint32_t tile_index = P.items_per_tile == 1u ? w : __umulhi(w, P.magic_items);
                const uint32_t c0 = (w - tile_index * P.items_per_tile) * P.strips_per_item;
                const uint32_t n_strips = min(P.strips_per_item, g - c0);
                unsigned char* image = images + (size_t)b * P.image_bytes;
                mbar_wait_parked(&H.full[b], use & 1u, 256u);   // every compute warp has written (and proxy-fenced) its records
                fence_proxy_async();
                unsigned char* out = reinterpret_cast<unsigned char*>(P.vertices) +
                                     ((size_t)tile_index * vpt + (size_t)c0 * vpr) * sizeof(bt_vertex);
                for (uint32_t s = 0; s < n_strips; ++s)
                    bulk_store(out + (size_t)s * strip_bytes, image + (size_t)s * strip_stride, strip_bytes);
                bulk_commit();
                bulk_wait_read<0>();                      // the TMA has read the image: hand it back
                mbar_arrive(&H.empty[b]);
            }
            bulk_wait<0>();                               // all global writes complete before the CTA exits
        }
        return;
    }

    // ---------------- compute warps ----------------
    const FrameConsts F = make_frame_consts<SPHERICAL>(U);
    uint32_t it = 0, rot = 0;
    for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const uint32_t b = it % GEN_BUFFERS, use = it / GEN_BUFFERS;
        const uint32_t tile_index = P.items_per_tile == 1u ? w : __umulhi(w, P.magic_items);
        const uint32_t c0 = (w - tile_index * P.items_per_tile) * P.strips_per_item;
        const uint32_t n_strips = min(P.strips_per_item, g - c0);
        const uint32_t n_cols = n_strips + 1u, n_rows = g + 1u;
        unsigned char* image = images + (size_t)b * P.image_bytes;
        if (use > 0u) mbar_wait_parked(&H.empty[b], (use - 1u) & 1u, 64u);   // the stores of the previous user of image b have read it

        const uint4 tq = __ldg(reinterpret_cast<const uint4*>(P.tiles) + tile_index);
        const bt_tile tile{tq.x, tq.y, tq.z, tq.w};
        const uint32_t n_verts = n_cols * n_rows;
        const uint32_t n_chunks = (n_verts + 31u) >> 5;
        // floor(i / n_cols) == umulhi(i, magic) for i * n_cols < 2^32; the full-width magic comes from the host
        const uint32_t magic = n_strips == P.strips_per_item ? P.magic_cols : P.magic_cols_ragged;
        // chunk c of this item goes to warp (c + rotation) mod GEN_COMPUTE_WARPS; the rotation spreads the ragged tail
        uint32_t first = (uint32_t)warp + GEN_COMPUTE_WARPS - rot;   // chunk c of this item goes to warp (rot + c) mod GEN_COMPUTE_WARPS
        if (first >= GEN_COMPUTE_WARPS) first -= GEN_COMPUTE_WARPS;
        rot += GEN_ROTATION % GEN_COMPUTE_WARPS;                     // spreads the ragged tail (measured: DESIGN.md section 6)
        if (rot >= GEN_COMPUTE_WARPS) rot -= GEN_COMPUTE_WARPS;
        for (uint32_t c = first; c < n_chunks; c += GEN_COMPUTE_WARPS) {
            const uint32_t i = (c << 5) + (uint32_t)lane;
            if (i < n_verts) {
                const uint32_t gy = __umulhi(i, magic), lx = i - gy * n_cols;  // lanes walk along texel rows
                const Record r = compute_vertex<SPHERICAL>(U, F, P.tile_tree, P.height_atlas, FLAGS_CT >= 0 ? (uint32_t)FLAGS_CT : P.flags, tile, tile_index, c0 + lx, gy);
                // strip `lx`: this vertex is (column, row 2*gy) -> position 2*gy+1 (+ duplicated first index)
                if (lx < n_strips) {
                    store_record<1>(image, strip_stride, lx, 2u * gy + 1u, r);
                    if (gy == 0u) store_record<0>(image, strip_stride, lx, 0u, r);
                }
                // strip `lx-1`: this vertex is (column+1, row 2*gy+1) -> position 2*gy+2 (+ duplicated last index)
                if (lx >= 1u) {
                    store_record<0>(image, strip_stride, lx - 1u, 2u * gy + 2u, r);
                    if (gy == g) store_record<1>(image, strip_stride, lx - 1u, vpr - 1u, r);   // vpr = 2 * (g + 2) is even
                }
            }
        }
        fence_proxy_async();   // this lane's generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&H.full[b]);
    }
}

}  // namespace btd

extern "C" cudaError_t bt_launch_gen_vertices(const void* uniforms, const bt_tile* final_tiles, const bt_indirect* indirect,
                                              const bt_tile_tree_entry* tile_tree, const uint16_t* height_atlas,
                                              bt_vertex* vertices, uint64_t vertex_capacity, btd::GenStatus* status,
                                              int spherical, uint32_t flags, uint32_t grid_size, int sm_count,
                                              cudaStream_t stream) {
    using namespace btd;
    if (grid_size == 0) return cudaErrorInvalidValue;
    const uint32_t strip_stride = strip_stride_for(2u * (grid_size + 2u) * (uint32_t)sizeof(bt_vertex));
    const size_t header = (sizeof(GenSmemHeader) + 127) & ~size_t(127);
    const size_t budget = (200 * 1024) / GEN_CTAS_PER_SM - header - 1024;  // per CTA; leaves L1 some room out of the 227 KiB
    uint32_t cb = (uint32_t)(budget / ((size_t)GEN_BUFFERS * strip_stride));
    if (cb > 16u) cb = 16u;
    if (cb > grid_size) cb = grid_size;
    if (cb == 0) return cudaErrorInvalidValue;  // grid_size too large for one strip per CTA
    GenParams P;
    P.uniforms = static_cast<const ViewUniforms*>(uniforms);
    P.tiles = final_tiles;
    P.indirect = indirect;
    P.tile_tree = tile_tree;
    P.height_atlas = height_atlas;
    P.vertices = vertices;
    P.vertex_capacity = vertex_capacity;
    P.status = status;
    P.flags = flags;
    P.grid_size = grid_size;
    P.strips_per_item = cb;
    P.items_per_tile = (grid_size + cb - 1u) / cb;
    P.image_bytes = cb * strip_stride;
    P.strip_stride = strip_stride;
    P.magic_items = P.items_per_tile == 1u ? 0u : 0xFFFFFFFFu / P.items_per_tile + 1u;
    P.magic_cols = 0xFFFFFFFFu / (cb + 1u) + 1u;
    {
        const uint32_t last_strips = grid_size - (P.items_per_tile - 1u) * cb;   // 1..cb
        P.magic_cols_ragged = 0xFFFFFFFFu / (last_strips + 1u) + 1u;
    }
    const size_t smem = header + (size_t)GEN_BUFFERS * P.image_bytes;
    // kernels specialised for the common flag sets (measured: 6 % faster than the generic kernel for NORMALS, 3 % for the
    // reference default + NORMALS); everything else (HIGH_PRECISION, partial sets) takes the generic one
    constexpr int kNone = 0, kNormals = BT_FLAG_NORMALS, kMorphBlend = BT_FLAG_MORPH | BT_FLAG_BLEND, kDefault = kMorphBlend | BT_FLAG_NORMALS;
    const uint32_t vertex_flags = flags & (BT_FLAG_MORPH | BT_FLAG_BLEND | BT_FLAG_HIGH_PRECISION | BT_FLAG_NORMALS);
    void (*kernel)(const GenParams);
    switch (vertex_flags) {
        case kNone: kernel = spherical ? gen_vertices_kernel<true, kNone> : gen_vertices_kernel<false, kNone>; break;
        case kNormals: kernel = spherical ? gen_vertices_kernel<true, kNormals> : gen_vertices_kernel<false, kNormals>; break;
        case kMorphBlend: kernel = spherical ? gen_vertices_kernel<true, kMorphBlend> : gen_vertices_kernel<false, kMorphBlend>; break;
        case kDefault: kernel = spherical ? gen_vertices_kernel<true, kDefault> : gen_vertices_kernel<false, kDefault>; break;
        default: kernel = spherical ? gen_vertices_kernel<true, -1> : gen_vertices_kernel<false, -1>; break;
    }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (sm_count > 0 ? sm_count : 148) * GEN_CTAS_PER_SM;
    kernel<<<grid, GEN_THREADS, smem, stream>>>(P);
    return cudaGetLastError();
}

// split_kernel.cuh — split.wgsl:18-43 (included by pyramid.cu after Fmt<> / PyramidLayout).
//
// Each thread owns a strip of N pixel columns x SPLIT_ROWS rows (N * sizeof(texel) = 8 bytes: one 64-bit store per row).
//   * Per pixel column the x terms (tile coordinate, source coordinate, bilinear weight, clamped texel columns) are
//     computed once per thread.
//   * Per tile row the y terms (tile coordinate, source coordinate, weight, clamped source rows, inside_square) depend on
//     (task, y) only; the CTA computes them once into a shared-memory table and every thread reads one 16-byte entry per
//     row (broadcast) instead of redoing a division and ~40 instructions.
//   * Per source row the horizontal lerp  h(row) = mix(texel(row, x0), texel(row, x1), rx)  and the validity carrier
//     min(red0, red1) are cached in registers, so at the native ~1:1 ratio every output row fetches ONE new source row
//     (2 texels per pixel instead of 4) and vertically adjacent outputs share h(row).  value = mix(h(y0), h(y1), ry) is
//     exactly the shader's mix(mix(v00,v10,rx), mix(v01,v11,rx), ry): same operands, same order, bit-identical.
//   * pack: clamp(v,0,1) is the .SAT modifier of the final add; floor(0.5 + M*c) for 0.5 <= . < 2^23 is the mantissa of
//     the round-down sum with 2^23 (one full-rate FADD.RM instead of FRND + F2I); the texel bits are merged with PRMT.
// The code is branch-free per pixel (clamped coordinates make every load legal; predicates select at the end).
#pragma once

constexpr int SPLIT_ROWS = 16;
constexpr int SPLIT_TAB_ROWS = 128;   // rows of y terms a CTA can cache (covers every tile size >= 128 px; smaller tiles compute inline)

#ifndef BT_SPLIT_MIN_BLOCKS
#define BT_SPLIT_MIN_BLOCKS 4   // resident CTAs per SM (R16; Rgba8 one more): the kernel hides its load latency by occupancy.
                                // Measured at 65536^2: R16 3/4/5 blocks -> 13.9/12.0/12.1 ms, Rgba8 3/4/5 -> 29.1/25.4/24.2 ms
#endif
#ifndef BT_SPLIT_N_R16
#define BT_SPLIT_N_R16 4
#endif
#ifndef BT_SPLIT_N_RGBA8
#define BT_SPLIT_N_RGBA8 2
#endif

template <int FORMAT> struct SplitCfg;
template <> struct SplitCfg<BT_FORMAT_R16> { static constexpr int N = BT_SPLIT_N_R16, CH = 1, WORDS = BT_SPLIT_N_R16 / 2; };    // R16 packs only .x (preprocessing.wgsl:85-88)
template <> struct SplitCfg<BT_FORMAT_RGBA8> { static constexpr int N = BT_SPLIT_N_RGBA8, CH = 4, WORDS = BT_SPLIT_N_RGBA8; };

// base + off as ONE IMAD.WIDE.U32 (the compiler otherwise rebuilds row * width + column in 64 bits for every texel)
__device__ __forceinline__ const void* ptr_add_u32(const void* base, uint32_t off) {
    uint64_t r;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(off), "l"(reinterpret_cast<uint64_t>(base)));
    return reinterpret_cast<const void*>(r);
}

struct SplitRow { float ry; int ya, yb; uint32_t flags; };   // flags bit 0: inside_square (y part); bit 1: centre row

template <int FORMAT>
__global__ void __launch_bounds__(PYR_THREADS, FORMAT == BT_FORMAT_RGBA8 ? BT_SPLIT_MIN_BLOCKS + 1 : BT_SPLIT_MIN_BLOCKS)
split_kernel(PyramidLayout L, const PyramidTask* __restrict__ tasks, const DatasetDevice* __restrict__ datasets) {
    using F = Fmt<FORMAT>;
    using texel = typename F::texel;
    constexpr int N = SplitCfg<FORMAT>::N, CH = SplitCfg<FORMAT>::CH, WORDS = SplitCfg<FORMAT>::WORDS;
    constexpr float M = FORMAT == BT_FORMAT_R16 ? 65535.0f : 255.0f;
    __shared__ float4 row_tab[SPLIT_TAB_ROWS];
    const PyramidTask& task = tasks[blockIdx.y];
    const DatasetDevice d = datasets[task.dataset];
    const uint32_t packets_per_row = (L.T + N - 1) / N;
    const uint32_t row_groups = (L.T + SPLIT_ROWS - 1) / SPLIT_ROWS;
    const uint32_t total = packets_per_row * row_groups;
    const uint32_t first_pid = blockIdx.x * PYR_THREADS;
    const uint32_t pid = first_pid + threadIdx.x;
    const float inv_tile_scale = exp2i(-(int)(task.tile.lod & 31u));  // x / tile_count(lod) == x * 2^-lod exactly
    const float Cf = (float)L.C;
    // inverse_mix with top_left = 0, bottom_right = 1 is (x - 0) / (1 - 0) = x exactly: skip the division
    const bool full_x = d.top_left[0] == 0.0f && d.bottom_right[0] == 1.0f;
    const bool full_y = d.top_left[1] == 0.0f && d.bottom_right[1] == 1.0f;
    const float range_x = d.bottom_right[0] - d.top_left[0], range_y = d.bottom_right[1] - d.top_left[1];
    const float Wf = (float)d.width, Hf = (float)d.height;
    const int wm1 = (int)d.width - 1, hm1 = (int)d.height - 1;

    // ---- per-row terms: CTA-wide table (the CTA covers the row groups of its first..last packet) ----
    auto row_terms = [&](uint32_t y) -> SplitRow {
        SplitRow t{0.0f, 0, 0, 0u};
        if (y >= L.B && y < L.B + L.C) {                        // border rows stay zero (split.wgsl:19-21)
            const float tcy = div_rn((float)(y - L.B), Cf);     // == IEEE quotient (bt_device.cuh)
            float sy = ((float)task.tile.y + tcy) * inv_tile_scale;
            if (!full_y) sy = (sy - d.top_left[1]) / range_y;   // inverse_mix — preprocessing.wgsl:44-46
            const float py = sy * Hf - 0.5f;
            const float fy = floorf(py);
            t.ry = py - fy;
            const int iy = (int)fy;
            t.ya = min(max(iy, 0), hm1);                        // clamp-to-edge (ImageSampler::linear, preprocessor.rs:408-409)
            t.yb = min(max(iy + 1, 0), hm1);
            t.flags = 2u | (uint32_t)(tcy >= 0.0f && tcy <= 1.0f);   // inside_square, y part
        }
        return t;
    };
    const uint32_t yg_first = first_pid / packets_per_row;
    const uint32_t yg_last = (min(first_pid + PYR_THREADS, total) - 1u) / packets_per_row;
    const uint32_t tab_rows = (yg_last - yg_first + 1u) * SPLIT_ROWS;
    const bool use_tab = tab_rows <= (uint32_t)SPLIT_TAB_ROWS;       // CTA-uniform
    if (use_tab) {
        for (uint32_t t = threadIdx.x; t < tab_rows; t += PYR_THREADS) {
            const SplitRow s = row_terms(yg_first * SPLIT_ROWS + t);
            row_tab[t] = make_float4(s.ry, __int_as_float(s.ya), __int_as_float(s.yb), __uint_as_float(s.flags));
        }
        __syncthreads();
    }
    if (pid >= total) return;
    const uint32_t yg = pid / packets_per_row, x0 = (pid - yg * packets_per_row) * N;
    texel* tile = reinterpret_cast<texel*>(L.arena + (size_t)task.slot * L.tile_bytes);
    const bool aligned = ((L.T * sizeof(texel)) % (4 * WORDS) == 0) && (L.tile_bytes % (4 * WORDS) == 0) && x0 + N <= L.T;
    const bool fresh = (task.flags & kTaskFresh) != 0u;
    const texel* __restrict__ src = reinterpret_cast<const texel*>(d.pixels);

    // ---- per-column terms (computed for every lane, masked at the end) ----
    float rx[N];
    uint32_t c0[N], c1[N];
    uint32_t col_centre = 0, col_write = 0;  // bit i: pixel column lies in the tile centre / ... and passes inside_square in x
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint32_t x = x0 + i;
        const float tcx = div_rn((float)((int)x - (int)L.B), Cf);
        float sx = ((float)task.tile.x + tcx) * inv_tile_scale;
        if (!full_x) sx = (sx - d.top_left[0]) / range_x;
        const float px = sx * Wf - 0.5f;
        const float fx = floorf(px);
        rx[i] = px - fx;
        const int ix = (int)fx;
        c0[i] = (uint32_t)min(max(ix, 0), wm1) * (uint32_t)sizeof(texel);        // clamp-to-edge; byte offsets in the row
        c1[i] = (uint32_t)min(max(ix + 1, 0), wm1) * (uint32_t)sizeof(texel);
        if (x < L.T && x >= L.B && x < L.B + L.C) {
            col_centre |= 1u << i;
            if (tcx >= 0.0f && tcx <= 1.0f) col_write |= 1u << i;   // inside_square, x part
        }
    }

    // ---- cached horizontal lerps of two source rows ----
    // Values are held as packed pairs (sm_100 FMUL2 / FFMA2 / FADD2: two IEEE f32 results per instruction, element-wise
    // identical to the scalar operations).  R16: pair j = pixels (2j, 2j+1); Rgba8: pairs (2i, 2i+1) = channels (r,g), (b,a)
    // of pixel i.  A packed product never feeds a packed add (ptxas would contract the pair into FFMA2): sums are scalar.
    constexpr int V2 = N * CH / 2;
    static_assert((N * CH) % 2 == 0, "pairs");
    float2 rx2[V2], omrx2[V2];     // bilinear x weight (and 1 - weight) of the pixel each pair element belongs to
#pragma unroll
    for (int j = 0; j < V2; ++j) {
        const int i0 = (2 * j) / CH, i1 = (2 * j + 1) / CH;
        rx2[j] = make_float2(rx[i0], rx[i1]);
        omrx2[j] = make_float2(1.0f - rx[i0], 1.0f - rx[i1]);
    }
    float2 hA[V2], hB[V2];
    float mA[N], mB[N];            // min(red0, red1): > 0 iff both texels of the row are valid (textureGather(0, ..) != 0)
    int rowA = -1, rowB = -1;      // clamped source row held in A / B
    auto load_row = [&](float2 (&h)[V2], float (&m)[N], int y) {
        // row window of a sharded source (bt_preprocess.cu validates it); clamp so a bad window cannot fault
        const int ry_ = min(max(y - (int)d.row_begin, 0), (int)d.row_count - 1);
        const texel* r = src + (size_t)ry_ * d.width;
        float2 t0[V2], t1[V2];
        if constexpr (FORMAT == BT_FORMAT_R16) {
            uint32_t a[N], b[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                a[i] = __ldg(static_cast<const uint16_t*>(ptr_add_u32(r, c0[i])));
                b[i] = __ldg(static_cast<const uint16_t*>(ptr_add_u32(r, c1[i])));
            }
#pragma unroll
            for (int j = 0; j < V2; ++j) {
                t0[j] = unorm16x2_to_float(a[2 * j], a[2 * j + 1]);
                t1[j] = unorm16x2_to_float(b[2 * j], b[2 * j + 1]);
                m[2 * j] = fminf(t0[j].x, t1[j].x);
                m[2 * j + 1] = fminf(t0[j].y, t1[j].y);
            }
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const uint32_t a = __ldg(static_cast<const uint32_t*>(ptr_add_u32(r, c0[i])));
                const uint32_t b = __ldg(static_cast<const uint32_t*>(ptr_add_u32(r, c1[i])));
                t0[2 * i] = unorm8x2_to_float(a & 0xFFu, (a >> 8) & 0xFFu);
                t0[2 * i + 1] = unorm8x2_to_float((a >> 16) & 0xFFu, a >> 24);
                t1[2 * i] = unorm8x2_to_float(b & 0xFFu, (b >> 8) & 0xFFu);
                t1[2 * i + 1] = unorm8x2_to_float((b >> 16) & 0xFFu, b >> 24);
                m[i] = fminf(t0[2 * i].x, t1[2 * i].x);
            }
        }
#pragma unroll
        for (int j = 0; j < V2; ++j) {   // mix(t0, t1, rx) = t0 * (1 - rx) + t1 * rx
            const float2 p = __fmul2_rn(t0[j], omrx2[j]), q = __fmul2_rn(t1[j], rx2[j]);
            h[j] = make_float2(p.x + q.x, p.y + q.y);
        }
    };

    // never-true threshold for columns that are not written: w = min(mA, mB) > wthresh
    float wthresh[N];
#pragma unroll
    for (int i = 0; i < N; ++i) wthresh[i] = (col_write >> i & 1u) ? 0.0f : __int_as_float(0x7f800000);

    // one output row from the cached lerps: P holds source row ya, Q holds yb
    auto emit = [&](const float2 (&P)[V2], const float (&mP)[N], const float2 (&Q)[V2], const float (&mQ)[N],
                    const SplitRow& t, const texel* row, uint32_t (&pix)[N]) {
        const bool inside_y = (t.flags & 1u) != 0u;
        const float2 ry2 = make_float2(t.ry, t.ry), omry2 = make_float2(1.0f - t.ry, 1.0f - t.ry);
        const float2 M2 = make_float2(M, M), big2 = make_float2(8388608.0f, 8388608.0f);
        uint32_t bits[2 * V2];     // pack: floor(0.5 + M * saturate(v)) in the low mantissa bits (see unorm_bits)
#pragma unroll
        for (int j = 0; j < V2; ++j) {
            const float2 p = __fmul2_rn(P[j], omry2), q = __fmul2_rn(Q[j], ry2);        // mix(P, Q, ry)
            const float2 s = __fmul2_rn(make_float2(__saturatef(p.x + q.x), __saturatef(p.y + q.y)), M2);
            const float2 f = __fadd2_rd(make_float2(0.5f + s.x, 0.5f + s.y), big2);
            bits[2 * j] = __float_as_uint(f.x);
            bits[2 * j + 1] = __float_as_uint(f.y);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const bool w = inside_y && fminf(mP[i], mQ[i]) > wthresh[i];
            uint32_t texel_bits;
            if constexpr (FORMAT == BT_FORMAT_R16) texel_bits = bits[i];
            else texel_bits = __byte_perm(__byte_perm(bits[4 * i], bits[4 * i + 1], 0x0040), __byte_perm(bits[4 * i + 2], bits[4 * i + 3], 0x0040), 0x5410);
            pix[i] = w ? texel_bits : 0u;
        }
        if (!fresh) {   // textureLoad(atlas, coords): centre pixels not written keep earlier datasets' texels
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const bool w = inside_y && fminf(mP[i], mQ[i]) > wthresh[i];
                if (!w && (col_centre >> i & 1u)) {
                    if constexpr (FORMAT == BT_FORMAT_R16) pix[i] = row[x0 + i];
                    else pix[i] = *reinterpret_cast<const uint32_t*>(row + x0 + i);
                }
            }
        }
    };

    const uint32_t y_first = yg * SPLIT_ROWS;
    const uint32_t n_rows = min((uint32_t)SPLIT_ROWS, L.T - y_first);
    texel* row = tile + (size_t)y_first * L.T;
    const float4* tab = row_tab + (yg - yg_first) * SPLIT_ROWS;
#pragma unroll 1
    for (uint32_t r = 0; r < n_rows; ++r, row += L.T) {
        uint32_t pix[N];           // texel bits of the N pixels (low 16 bits for R16)
#pragma unroll
        for (int i = 0; i < N; ++i) pix[i] = 0u;
        SplitRow t;
        if (use_tab) {
            const float4 e = tab[r];
            t = SplitRow{e.x, __float_as_int(e.y), __float_as_int(e.z), __float_as_uint(e.w)};
        } else {
            t = row_terms(y_first + r);
        }
        if ((t.flags & 2u) != 0u && col_centre != 0u) {   // warp-uniform in y
            // roles instead of register copies: the buffer that already holds ya plays "row ya"
            if (rowA != t.ya && rowB == t.ya) {
                if (rowA != t.yb) { load_row(hA, mA, t.yb); rowA = t.yb; }
                emit(hB, mB, hA, mA, t, row, pix);
            } else {
                if (rowA != t.ya) { load_row(hA, mA, t.ya); rowA = t.ya; }
                if (rowB != t.yb) { load_row(hB, mB, t.yb); rowB = t.yb; }
                emit(hA, mA, hB, mB, t, row, pix);
            }
        }
        uint32_t words[WORDS];
        if constexpr (FORMAT == BT_FORMAT_R16) {
#pragma unroll
            for (int j = 0; j < WORDS; ++j) words[j] = __byte_perm(pix[2 * j], pix[2 * j + 1], 0x5410);
        } else {
#pragma unroll
            for (int j = 0; j < WORDS; ++j) words[j] = pix[j];
        }
        if (aligned) {
            if constexpr (WORDS == 1) *reinterpret_cast<uint32_t*>(row + x0) = words[0];
            else if constexpr (WORDS == 2) *reinterpret_cast<uint2*>(row + x0) = make_uint2(words[0], words[1]);
            else *reinterpret_cast<uint4*>(row + x0) = make_uint4(words[0], words[1], words[2], words[3]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (x0 + i < L.T) {
                    if constexpr (FORMAT == BT_FORMAT_R16) row[x0 + i] = (uint16_t)pix[i];
                    else *reinterpret_cast<uint32_t*>(row + x0 + i) = pix[i];
                }
            }
        }
    }
}

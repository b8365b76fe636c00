// split_staged.cuh — split.wgsl:18-43, second form (included by pyramid.cu after split_kernel.cuh).
//
// Same arithmetic, same operand order and the same per-thread register cache of horizontal lerps as split_kernel, but the
// source texels come from shared memory: the kernel of split_kernel.cuh hides its 2-byte / 4-byte gather loads (DRAM misses,
// ~800 cycles) by occupancy alone and the register file caps that at 50 %, which left 40 % of the issue slots empty.
//   * A CTA owns 512 pixel columns x `bands` bands of SPLIT_ROWS rows of ONE tile.  Column terms are computed once per CTA
//     lifetime (not once per strip), row terms once per CTA into a shared-memory table.
//   * For every band the source window  rows [ya(first centre row), yb(last centre row)] x columns [c0(first), c1(last)]
//     is fetched by ONE thread as 1-D TMA bulk copies (one per source row, completion on an mbarrier) into a stage buffer;
//     with two stages the window of band j+1 is in flight while band j is computed.  No register, no LSU slot and no 64-bit
//     address arithmetic is spent on the fetch; the consumers read 2 / 4-byte texels with LDS (2-way bank conflicts at most).
//   * Windows that do not fit the stage geometry the host chose (a source much denser than the tile grid), or sources whose
//     rows are not 16-byte aligned, make the CTA fall back to the global-load form of load_row: same results, old speed.
#pragma once

#ifndef BT_SPLIT_R16_BLOCKS
#define BT_SPLIT_R16_BLOCKS 5     // resident CTAs per SM the register allocation allows (launch bounds)
#endif
#ifndef BT_SPLIT_RGBA8_BLOCKS
#define BT_SPLIT_RGBA8_BLOCKS 3   // 71 registers; with one 46 KB stage three CTAs (24 warps) share an SM.  Measured at 65536^2 (whole
#endif                            // albedo pyramid): 2 blocks + 2 stages 22.7 ms, 3 blocks + 1 stage 22.1 ms, 4 blocks (spills) 22.1 ms
template <int FORMAT> struct StagedCfg;
template <> struct StagedCfg<BT_FORMAT_R16> { static constexpr int N = 4, THREADS = 128, MIN_BLOCKS = BT_SPLIT_R16_BLOCKS; };
template <> struct StagedCfg<BT_FORMAT_RGBA8> { static constexpr int N = 2, THREADS = 256, MIN_BLOCKS = BT_SPLIT_RGBA8_BLOCKS; };
constexpr int SPLIT_MAX_BANDS = 32;       // bands one CTA may own (512 rows)
constexpr int SPLIT_STAGE_COLS = 512;     // pixel columns one CTA owns = THREADS * N for both formats

struct SplitStageGeom {
    uint32_t rows;     // source rows one stage holds (0: no staging, every CTA uses global loads)
    uint32_t pitch;    // bytes per staged source row (multiple of 16)
    uint32_t stages;   // 1 or 2
    uint32_t bands;    // bands per CTA (1 .. SPLIT_MAX_BANDS)
};

__host__ __device__ inline uint32_t split_staged_smem_bytes(const SplitStageGeom& G) {
    return 64u + SPLIT_MAX_BANDS * 8u + G.bands * SPLIT_ROWS * 16u + G.stages * G.rows * G.pitch;
}

template <int FORMAT>
__global__ void __launch_bounds__(StagedCfg<FORMAT>::THREADS, StagedCfg<FORMAT>::MIN_BLOCKS)
split_staged_kernel(PyramidLayout L, const PyramidTask* __restrict__ tasks, const DatasetDevice* __restrict__ datasets, SplitStageGeom G) {
    using F = Fmt<FORMAT>;
    using texel = typename F::texel;
    constexpr int N = StagedCfg<FORMAT>::N, THREADS = StagedCfg<FORMAT>::THREADS, CH = SplitCfg<FORMAT>::CH;
    constexpr int WORDS = FORMAT == BT_FORMAT_R16 ? N / 2 : N;
    constexpr float M = FORMAT == BT_FORMAT_R16 ? 65535.0f : 255.0f;
    static_assert(N * THREADS == SPLIT_STAGE_COLS, "a CTA covers 512 columns");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                    // [2] one per stage
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem + 16);               // [0] min c0, [1] max c1 (byte offsets in a source row)
    int2* band_win = reinterpret_cast<int2*>(smem + 64);                   // [SPLIT_MAX_BANDS] source rows (lo, hi) of a band; hi < lo: none
    float4* row_tab = reinterpret_cast<float4*>(smem + 64 + SPLIT_MAX_BANDS * 8);
    unsigned char* stage_mem = smem + 64 + SPLIT_MAX_BANDS * 8 + G.bands * SPLIT_ROWS * 16;

    const PyramidTask& task = tasks[blockIdx.y];
    const DatasetDevice d = datasets[task.dataset];
    const uint32_t col_blocks = (L.T + SPLIT_STAGE_COLS - 1) / SPLIT_STAGE_COLS;
    const uint32_t cb = blockIdx.x % col_blocks, bg = blockIdx.x / col_blocks;
    const uint32_t y_begin = bg * G.bands * SPLIT_ROWS;
    const uint32_t n_rows_cta = min(G.bands * SPLIT_ROWS, L.T - y_begin);
    const uint32_t n_bands = (n_rows_cta + SPLIT_ROWS - 1) / SPLIT_ROWS;
    const uint32_t x0 = (cb * THREADS + threadIdx.x) * N;
    const float inv_tile_scale = exp2i(-(int)(task.tile.lod & 31u));
    const float Cf = (float)L.C;
    const bool full_x = d.top_left[0] == 0.0f && d.bottom_right[0] == 1.0f;
    const bool full_y = d.top_left[1] == 0.0f && d.bottom_right[1] == 1.0f;
    const float range_x = d.bottom_right[0] - d.top_left[0], range_y = d.bottom_right[1] - d.top_left[1];
    const float Wf = (float)d.width, Hf = (float)d.height;
    const int wm1 = (int)d.width - 1, hm1 = (int)d.height - 1;

    if (threadIdx.x == 0) {
        misc[0] = 0xFFFFFFFFu;
        misc[1] = 0u;
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- per-row terms of the CTA's rows (ya / yb already clamped into the rows present behind d.pixels) ----
    for (uint32_t t = threadIdx.x; t < n_rows_cta; t += THREADS) {
        const uint32_t y = y_begin + t;
        float ry = 0.0f;
        int ya = 0, yb = 0;
        uint32_t flags = 0u;
        if (y >= L.B && y < L.B + L.C) {                        // border rows stay zero (split.wgsl:19-21)
            const float tcy = div_rn((float)(y - L.B), Cf);
            float sy = ((float)task.tile.y + tcy) * inv_tile_scale;
            if (!full_y) sy = (sy - d.top_left[1]) / range_y;   // inverse_mix — preprocessing.wgsl:44-46
            const float py = sy * Hf - 0.5f;
            const float fy = floorf(py);
            ry = py - fy;
            const int iy = (int)fy;
            ya = min(max(min(max(iy, 0), hm1) - (int)d.row_begin, 0), (int)d.row_count - 1);      // clamp-to-edge, then the row window
            yb = min(max(min(max(iy + 1, 0), hm1) - (int)d.row_begin, 0), (int)d.row_count - 1);
            flags = 2u | (uint32_t)(tcy >= 0.0f && tcy <= 1.0f);
        }
        row_tab[t] = make_float4(ry, __int_as_float(ya), __int_as_float(yb), __uint_as_float(flags));
    }

    // ---- per-column terms ----
    float rx[N];
    uint32_t c0[N], c1[N];
    uint32_t col_centre = 0, col_write = 0;
    uint32_t cmin = 0xFFFFFFFFu, cmax = 0u;
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
        c0[i] = (uint32_t)min(max(ix, 0), wm1) * (uint32_t)sizeof(texel);
        c1[i] = (uint32_t)min(max(ix + 1, 0), wm1) * (uint32_t)sizeof(texel);
        if (x < L.T && x >= L.B && x < L.B + L.C) {
            col_centre |= 1u << i;
            if (tcx >= 0.0f && tcx <= 1.0f) col_write |= 1u << i;
            cmin = min(cmin, min(c0[i], c1[i]));
            cmax = max(cmax, max(c0[i], c1[i]));
        }
    }
    cmin = __reduce_min_sync(0xFFFFFFFFu, cmin);
    cmax = __reduce_max_sync(0xFFFFFFFFu, cmax);
    if ((threadIdx.x & 31u) == 0u && cmin != 0xFFFFFFFFu) {
        atomicMin(&misc[0], cmin);
        atomicMax(&misc[1], cmax);
    }
    __syncthreads();

    // ---- source window of every band; does the CTA fit the stage geometry? ----
    const uint32_t src_pitch = d.width * (uint32_t)sizeof(texel);
    const uint32_t cwin0 = misc[0] == 0xFFFFFFFFu ? 0u : (misc[0] & ~15u);
    const uint32_t cbytes = misc[0] == 0xFFFFFFFFu ? 0u : ((misc[1] + (uint32_t)sizeof(texel) - cwin0 + 15u) & ~15u);
    int fits = 1;
    // the staged form reads texel x1 at "offset of x0 + one texel": true everywhere but at the clamped image edges (such CTAs
    // - the first / last tile column of a dataset - take the global-load form)
#pragma unroll
    for (int i = 0; i < N; ++i) if ((col_centre >> i & 1u) != 0u && c1[i] != c0[i] + (uint32_t)sizeof(texel)) fits = 0;
    if (threadIdx.x < n_bands) {
        int lo = 0x7fffffff, hi = -1;
        const uint32_t r_end = min((threadIdx.x + 1u) * SPLIT_ROWS, n_rows_cta);
        for (uint32_t t = threadIdx.x * SPLIT_ROWS; t < r_end; ++t) {
            const float4 e = row_tab[t];
            if ((__float_as_uint(e.w) & 2u) != 0u) {
                lo = min(lo, min(__float_as_int(e.y), __float_as_int(e.z)));
                hi = max(hi, max(__float_as_int(e.y), __float_as_int(e.z)));
            }
        }
        if (cbytes == 0u) { lo = 0; hi = -1; }   // no centre column in this block: nothing is sampled
        band_win[threadIdx.x] = make_int2(lo, hi);
        if (hi >= lo && (uint32_t)(hi - lo + 1) > G.rows) fits = 0;
    }
    // the bulk copies need 16-byte aligned global addresses: the row pitch and the base pointer; the last row's copy may
    // run up to 15 bytes past cmax, so the window must end inside the row (rows are contiguous: only the image's last row matters)
    if (G.rows == 0u || cbytes > G.pitch || (src_pitch & 15u) != 0u || (reinterpret_cast<uint64_t>(d.pixels) & 15u) != 0u || cwin0 + cbytes > src_pitch) fits = 0;
    const bool staged = __syncthreads_and(fits) != 0;

    texel* tile = reinterpret_cast<texel*>(L.arena + (size_t)task.slot * L.tile_bytes);
    const bool aligned = ((L.T * sizeof(texel)) % (4 * WORDS) == 0) && (L.tile_bytes % (4 * WORDS) == 0) && x0 + N <= L.T;
    const bool fresh = (task.flags & kTaskFresh) != 0u;
    const unsigned char* __restrict__ src = static_cast<const unsigned char*>(d.pixels);

    auto issue = [&](uint32_t j) {   // thread 0: fetch the window of band j into stage j % stages
        const int2 w = band_win[j];
        uint64_t* bar = &bars[j % G.stages];
        if (w.y < w.x) { mbar_arrive(bar); return; }
        const uint32_t rows = (uint32_t)(w.y - w.x + 1);
        mbar_expect_tx(bar, rows * cbytes);
        unsigned char* dst = stage_mem + (size_t)(j % G.stages) * G.rows * G.pitch;
        const unsigned char* s = src + (size_t)w.x * src_pitch + cwin0;
        for (uint32_t r = 0; r < rows; ++r) bulk_load(dst + r * G.pitch, s + (size_t)r * src_pitch, cbytes, bar);
    };
    if (staged && threadIdx.x == 0) issue(0);

    constexpr int V2 = N * CH / 2;
    float2 rx2[V2], omrx2[V2];
#pragma unroll
    for (int j = 0; j < V2; ++j) {
        const int i0 = (2 * j) / CH, i1 = (2 * j + 1) / CH;
        rx2[j] = make_float2(rx[i0], rx[i1]);
        omrx2[j] = make_float2(1.0f - rx[i0], 1.0f - rx[i1]);
    }
    float wthresh[N];
#pragma unroll
    for (int i = 0; i < N; ++i) wthresh[i] = (col_write >> i & 1u) ? 0.0f : __int_as_float(0x7f800000);
    const uint32_t stage_off = (uint32_t)(stage_mem - smem);

    // The band loop, instantiated twice (CTA-uniform choice): STAGED reads the TMA-filled stage, the other form gathers from
    // global memory.  (One body with a run-time `if` around the loads made ptxas merge the two 16-bit load results through
    // PRMT + LOP3 pairs: 24 ALU instructions per source row instead of 8.)
    auto run = [&](auto staged_tag) {
        constexpr bool STAGED = decltype(staged_tag)::value;
        if constexpr (STAGED) {   // column byte offsets inside a staged row (columns that are not sampled read offset 0: inside the stage)
#pragma unroll
            for (int i = 0; i < N; ++i) c0[i] = (col_centre >> i & 1u) != 0u ? c0[i] - cwin0 : 0u;
        }
        bool adjacent[N];   // column i samples the texel right of column i - 1's (staged offsets: c0[i] == c0[i - 1] + one texel)
        adjacent[0] = false;
#pragma unroll
        for (int i = 1; i < N; ++i) adjacent[i] = c0[i] == c0[i - 1] + (uint32_t)sizeof(texel);
        float2 hA[V2], hB[V2];
        float mA[N], mB[N];            // product of the two texels' red channels: > 0 iff both are valid (textureGather(0, ..) != 0)
        int rowA = -1, rowB = -1;
        uint32_t stage_rows = stage_off;   // smem byte offset of the current band's stage, biased by -lo * pitch (wrapping arithmetic)
        auto load_row = [&](float2 (&h)[V2], float (&m)[N], int y) {
            float2 t0[V2], t1[V2];
            if constexpr (FORMAT == BT_FORMAT_R16) {
                uint32_t a[N], b[N];
                if constexpr (STAGED) {
                    const uint32_t r = stage_rows + (uint32_t)y * G.pitch;
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const unsigned char* p = smem + (r + c0[i]);
                        a[i] = *reinterpret_cast<const uint16_t*>(p);
                        b[i] = *reinterpret_cast<const uint16_t*>(p + 2);
                    }
                } else {
                    const unsigned char* r = src + (size_t)y * src_pitch;
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        a[i] = __ldg(static_cast<const uint16_t*>(ptr_add_u32(r, c0[i])));
                        b[i] = __ldg(static_cast<const uint16_t*>(ptr_add_u32(r, c1[i])));
                    }
                }
#pragma unroll
                for (int j = 0; j < V2; ++j) {
                    t0[j] = unorm16x2_to_float(a[2 * j], a[2 * j + 1]);
                    t1[j] = unorm16x2_to_float(b[2 * j], b[2 * j + 1]);
                    const float2 mm = __fmul2_rn(t0[j], t1[j]);
                    m[2 * j] = mm.x;
                    m[2 * j + 1] = mm.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    uint32_t a, b;
                    if constexpr (STAGED) {
                        const unsigned char* p = smem + (stage_rows + (uint32_t)y * G.pitch + c0[i]);
                        a = *reinterpret_cast<const uint32_t*>(p);
                        b = *reinterpret_cast<const uint32_t*>(p + 4);
                    } else {
                        const unsigned char* r = src + (size_t)y * src_pitch;
                        a = __ldg(static_cast<const uint32_t*>(ptr_add_u32(r, c0[i])));
                        b = __ldg(static_cast<const uint32_t*>(ptr_add_u32(r, c1[i])));
                    }
                    // the x0 texel of column i is the x1 texel of column i - 1 unless the source skips a column between them
                    // (every ~127 columns at the native ratio): reuse its decode (the same bits) instead of repeating it
                    if (STAGED && i > 0 && adjacent[i]) { t0[2 * i] = t1[2 * i - 2]; t0[2 * i + 1] = t1[2 * i - 1]; }
                    else unorm8x4_to_float(a, t0[2 * i], t0[2 * i + 1]);
                    unorm8x4_to_float(b, t1[2 * i], t1[2 * i + 1]);
                    m[i] = t0[2 * i].x * t1[2 * i].x;
                }
            }
#pragma unroll
            for (int j = 0; j < V2; ++j) {   // mix(t0, t1, rx) = t0 * (1 - rx) + t1 * rx
                const float2 p = __fmul2_rn(t0[j], omrx2[j]), q = __fmul2_rn(t1[j], rx2[j]);
                h[j] = make_float2(p.x + q.x, p.y + q.y);
            }
        };

        // One output row from the cached lerps (P holds source row ya, Q holds yb).  A pixel is written iff the four texels' red
        // channels are all non-zero and the column passes inside_square: `ind` is 1.0 then, else 0.0, and scales the pack
        // factor, so an unwritten pixel packs floor(0.5 + 0) = 0 without a select.
        auto emit = [&](const float2 (&P)[V2], const float (&mP)[N], const float2 (&Q)[V2], const float (&mQ)[N],
                        float ry, bool inside_y, const texel* row, uint32_t (&pix)[N]) {
            const float2 ry2 = make_float2(ry, ry), omry2 = make_float2(1.0f - ry, 1.0f - ry);
            const float2 M2 = make_float2(M, M), big2 = make_float2(8388608.0f, 8388608.0f);
            float ind[N];
            if constexpr (FORMAT == BT_FORMAT_R16) {
#pragma unroll
                for (int j = 0; j < N / 2; ++j) {
                    const float2 pm = __fmul2_rn(make_float2(mP[2 * j], mP[2 * j + 1]), make_float2(mQ[2 * j], mQ[2 * j + 1]));
                    ind[2 * j] = pm.x > wthresh[2 * j] ? 1.0f : 0.0f;
                    ind[2 * j + 1] = pm.y > wthresh[2 * j + 1] ? 1.0f : 0.0f;
                }
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) ind[i] = mP[i] * mQ[i] > wthresh[i] ? 1.0f : 0.0f;
            }
            uint32_t bits[2 * V2];
#pragma unroll
            for (int j = 0; j < V2; ++j) {
                const int i0 = (2 * j) / CH, i1 = (2 * j + 1) / CH;
                const float2 Mi = __fmul2_rn(make_float2(ind[i0], ind[i1]), M2);
                const float2 p = __fmul2_rn(P[j], omry2), q = __fmul2_rn(Q[j], ry2);        // mix(P, Q, ry)
                const float2 sv = __fmul2_rn(make_float2(__saturatef(p.x + q.x), __saturatef(p.y + q.y)), Mi);
                const float2 f = __fadd2_rd(make_float2(0.5f + sv.x, 0.5f + sv.y), big2);
                bits[2 * j] = __float_as_uint(f.x);
                bits[2 * j + 1] = __float_as_uint(f.y);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if constexpr (FORMAT == BT_FORMAT_R16) pix[i] = bits[i];
                else pix[i] = __byte_perm(__byte_perm(bits[4 * i], bits[4 * i + 1], 0x0040), __byte_perm(bits[4 * i + 2], bits[4 * i + 3], 0x0040), 0x5410);
            }
            if (!fresh) {   // textureLoad(atlas, coords): centre pixels not written keep earlier datasets' texels
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const bool w = inside_y && ind[i] != 0.0f;
                    if (!w) {
                        pix[i] = 0u;
                        if (col_centre >> i & 1u) {
                            if constexpr (FORMAT == BT_FORMAT_R16) pix[i] = row[x0 + i];
                            else pix[i] = *reinterpret_cast<const uint32_t*>(row + x0 + i);
                        }
                    }
                }
            }
        };

        texel* row = tile + (size_t)y_begin * L.T;
#pragma unroll 1
        for (uint32_t j = 0; j < n_bands; ++j) {
            if constexpr (STAGED) {
                if (G.stages == 2u && j + 1u < n_bands && threadIdx.x == 0) issue(j + 1u);   // stage (j+1)%2 was released by the barrier that ended band j-1
                mbar_wait(&bars[j % G.stages], (j / G.stages) & 1u);
                stage_rows = stage_off + (j % G.stages) * G.rows * G.pitch - (uint32_t)band_win[j].x * G.pitch;
            }
            const uint32_t rows_here = min((uint32_t)SPLIT_ROWS, n_rows_cta - j * SPLIT_ROWS);
            const float4* tab = row_tab + j * SPLIT_ROWS;
            if (x0 < L.T) {
#pragma unroll 1
                for (uint32_t r = 0; r < rows_here; ++r, row += L.T) {
                    uint32_t pix[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) pix[i] = 0u;
                    const float4 e = tab[r];
                    const int ya = __float_as_int(e.y), yb = __float_as_int(e.z);
                    const uint32_t flags = __float_as_uint(e.w);
                    // a centre row outside inside_square writes nothing: fresh tiles skip it (warp-uniform), others re-read the tile
                    if ((flags & 2u) != 0u && col_centre != 0u && ((flags & 1u) != 0u || !fresh)) {
                        if (rowA != ya && rowB == ya) {
                            if (rowA != yb) { load_row(hA, mA, yb); rowA = yb; }
                            emit(hB, mB, hA, mA, e.x, (flags & 1u) != 0u, row, pix);
                        } else {
                            if (rowA != ya) { load_row(hA, mA, ya); rowA = ya; }
                            if (rowB != yb) { load_row(hB, mB, yb); rowB = yb; }
                            emit(hA, mA, hB, mB, e.x, (flags & 1u) != 0u, row, pix);
                        }
                    }
                    uint32_t words[WORDS];
                    if constexpr (FORMAT == BT_FORMAT_R16) {
#pragma unroll
                        for (int k = 0; k < WORDS; ++k) words[k] = __byte_perm(pix[2 * k], pix[2 * k + 1], 0x5410);
                    } else {
#pragma unroll
                        for (int k = 0; k < WORDS; ++k) words[k] = pix[k];
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
            if constexpr (STAGED) {
                __syncthreads();   // every thread has read its last texel of this band's stage
                if (G.stages == 1u && j + 1u < n_bands && threadIdx.x == 0) issue(j + 1u);
            }
        }
    };
    if (staged) run(std::true_type{});
    else run(std::false_type{});
}

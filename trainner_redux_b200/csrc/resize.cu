// resize: separable resampling with ATen's index/weight rules, + fused clamp(0,1).
// Replaces traiNNer/data/degradations.py:1004-1021 (`resize_pt`): F.interpolate with
// bilinear/bicubic antialias=True, area (adaptive avg pool), nearest-exact, and the plain
// (non-aa, a=-0.75) bicubic that ends the "lanczos" mode.
//
// Every mode is expressed as, per output index o on an axis, a contiguous source window
// [lo, lo+n) with n normalised fp32 weights (ATen: aten/src/ATen/native/cpu/UpSampleKernel.cpp
// `_compute_indices_min_size_weights_aa`; restated in SURVEY.md §8a row a3 "R").  A tiny first
// launch writes the per-axis tables once; then a CTA owns an output tile: it stages its slice of
// the tables and the source rectangle it needs in shared memory with cp.async, runs the horizontal
// pass (result kept in smem), then the vertical pass — the same order ATen uses (W first, then H).
// HBM-bound: the source is read once per tile (+ halo, served by L2), the output written once.
#include <stdlib.h>
#include <string.h>

#include "otf_common.cuh"

namespace otf {


struct AxisSpec {
    int in_n, out_n;
    float scale;    // in/out, fp32 as ATen computes it (area_pixel_compute_scale with size given)
    float support;  // aa filter half-width in source pixels
    int max_taps;
};

// Keys cubic pieces, evaluated op by op in fp32 exactly as ATen's cubic_convolution1/2 (no FMA
// contraction: the CPU reference rounds after every operation, and at coordinates ~300 a fused
// evaluation moves weights by 1e-5).
__device__ __forceinline__ float cubic1(float x, float A) {
    // ((A + 2) * x - (A + 3)) * x * x + 1
    const float t = __fsub_rn(__fmul_rn(A + 2.0f, x), A + 3.0f);
    return __fadd_rn(__fmul_rn(__fmul_rn(t, x), x), 1.0f);
}
__device__ __forceinline__ float cubic2(float x, float A) {
    // ((A * x - 5 * A) * x + 8 * A) * x - 4 * A
    const float t = __fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x), 5.0f * A), x), 8.0f * A);
    return __fsub_rn(__fmul_rn(t, x), 4.0f * A);
}
__device__ __forceinline__ float aa_filter(int mode, float x) {
    x = fabsf(x);
    if (mode == OTF_RESIZE_BILINEAR_AA) return x < 1.0f ? __fsub_rn(1.0f, x) : 0.0f;
    if (x < 1.0f) return cubic1(x, -0.5f);
    if (x < 2.0f) return cubic2(x, -0.5f);
    return 0.0f;
}

// Fill (lo, n, w[0..max_taps)) for output index o. Weights for indices folded by edge clamping
// (non-aa bicubic) are accumulated onto the clamped source index.
__device__ void axis_weights(int mode, const AxisSpec& ax, int o, int* lo_out, int* n_out, float* w) {
    int lo = 0, n = 0;
    if (mode == OTF_RESIZE_BILINEAR_AA || mode == OTF_RESIZE_BICUBIC_AA) {
        // ATen `_compute_indices_min_size_weights_aa` with scalar_t = float. The C++ mixes in double
        // through its 0.5 / 1.0 literals; the promotions are reproduced literally.
        const float center = (float)((double)ax.scale * ((double)o + 0.5));
        const float invscale = ax.scale >= 1.0f ? (float)(1.0 / (double)ax.scale) : 1.0f;
        lo = max((int)((double)__fsub_rn(center, ax.support) + 0.5), 0);
        n = min((int)((double)__fadd_rn(center, ax.support) + 0.5), ax.in_n) - lo;
        n = clampi(n, 0, ax.max_taps);
        float total = 0.0f;
        for (int j = 0; j < n; ++j) {
            const float arg = (float)(((double)__fsub_rn((float)(j + lo), center) + 0.5) * (double)invscale);
            const float v = aa_filter(mode, arg);
            w[j] = v;
            total = __fadd_rn(total, v);
        }
        if (total != 0.0f)
            for (int j = 0; j < n; ++j) w[j] = __fdiv_rn(w[j], total);
    } else if (mode == OTF_RESIZE_AREA) {
        // adaptive_avg_pool: [floor(o*in/out), ceil((o+1)*in/out))
        lo = (int)(((int64_t)o * ax.in_n) / ax.out_n);
        const int hi = (int)((((int64_t)o + 1) * ax.in_n + ax.out_n - 1) / ax.out_n);
        n = hi - lo;
        const float inv = __fdiv_rn(1.0f, (float)n);
        for (int j = 0; j < n; ++j) w[j] = inv;
    } else if (mode == OTF_RESIZE_NEAREST_EXACT) {
        lo = min((int)floorf(__fmul_rn((float)o + 0.5f, ax.scale)), ax.in_n - 1);
        n = 1;
        w[0] = 1.0f;
    } else if (mode == OTF_RESIZE_NEAREST) {
        // legacy nearest (UpSample.h nearest_idx): identity / halving shortcuts, else floorf(o * scale)
        if (ax.out_n == ax.in_n) lo = o;
        else if (ax.out_n == 2 * ax.in_n) lo = o >> 1;
        else lo = min((int)floorf(__fmul_rn((float)o, ax.scale)), ax.in_n - 1);
        n = 1;
        w[0] = 1.0f;
    } else {  // OTF_RESIZE_BICUBIC: src = scale*(o+0.5)-0.5, 4 taps, A=-0.75, indices clamped
        // ATen's CPU build contracts `scale * (o + 0.5) - 0.5` into one fused multiply-add (checked against
        // the installed torch at 288->431: the unfused form is 1.4e-5 off, the fused one matches)
        const float src = fmaf(ax.scale, (float)o + 0.5f, -0.5f);
        const float fl = floorf(src);
        const float t = __fsub_rn(src, fl);
        const int i0 = (int)fl - 1;
        const float A = -0.75f;
        const float c[4] = {cubic2(__fadd_rn(t, 1.0f), A), cubic1(t, A), cubic1(__fsub_rn(1.0f, t), A),
                            cubic2(__fadd_rn(__fsub_rn(1.0f, t), 1.0f), A)};
        lo = clampi(i0, 0, ax.in_n - 1);
        const int hi = clampi(i0 + 3, 0, ax.in_n - 1);
        n = hi - lo + 1;
        for (int j = 0; j < n; ++j) w[j] = 0.0f;
        for (int k = 0; k < 4; ++k) {
            float* wk = &w[clampi(i0 + k, 0, ax.in_n - 1) - lo];
            *wk = __fadd_rn(*wk, c[k]);
        }
    }
    for (int j = n; j < ax.max_taps; ++j) w[j] = 0.0f;
    *lo_out = lo;
    *n_out = n;
}

// "lanczos" mode of resize_pt (degradations.py:961-1001) as ONE separable resampling: on a shrinking axis the
// reference convolves with a Lanczos-3 prefilter (reflect padding) and then takes the plain 4-tap bicubic sample;
// both are linear, so their composition is a single window of 4 + 2r weights per output (r = prefilter radius) —
// one launch and one pass over the image instead of three.  The prefilter taps follow :961-978 operation by
// operation: positions accumulated in double and stored as fp32, sinc(t) * sinc(t / 3) in fp32, normalised.
constexpr int kLanczosMaxR = 62;
__device__ void lanczos_axis_weights(const AxisSpec& ax, int o, int* lo_out, int* n_out, float* w) {
    AxisSpec bc = ax;
    bc.max_taps = 4;
    int lo4, n4;
    float c4[4];
    axis_weights(OTF_RESIZE_BICUBIC, bc, o, &lo4, &n4, c4);
    if (ax.out_n >= ax.in_n) {  // no prefilter on an axis that does not shrink
        for (int j = 0; j < ax.max_taps; ++j) w[j] = j < 4 ? c4[j] : 0.0f;
        *lo_out = lo4;
        *n_out = n4;
        return;
    }
    const double ratio = (double)ax.out_n / (double)ax.in_n;
    const int n = (int)ceil(3.0 / ratio + 1.0), r = n - 2, nt = 2 * n - 3;
    float f[2 * kLanczosMaxR + 1];
    float total = 0.0f;
    for (int j = 0; j < nt; ++j) {
        const int i = j < r ? r - j : j - r;             // |index| into the ramp
        double cur = 0.0;  // the reference's ramp accumulates (cur += ratio) in double and stores fp32: same running sum
        for (int t = 0; t < i; ++t) cur += ratio;
        float pos = (float)cur;
        if (j < r) pos = -pos;
        float v = 0.0f;
        if (-3.0f < pos && pos < 3.0f) {
            const float px = __fmul_rn(3.14159274101257324f, pos), p3 = __fmul_rn(3.14159274101257324f, __fdiv_rn(pos, 3.0f));
            const float s1 = pos != 0.0f ? __fdiv_rn(sinf(px), px) : 1.0f;
            const float s3 = __fdiv_rn(pos, 3.0f) != 0.0f ? __fdiv_rn(sinf(p3), p3) : 1.0f;
            v = __fmul_rn(s1, s3);
        }
        f[j] = v;
        total = __fadd_rn(total, v);
    }
    for (int j = 0; j < nt; ++j) f[j] = __fdiv_rn(f[j], total);
    const int LO = max(lo4 - r, 0), HI = min(lo4 + n4 - 1 + r, ax.in_n - 1);
    for (int j = 0; j < ax.max_taps; ++j) w[j] = 0.0f;
    for (int k = 0; k < n4; ++k)
        for (int j = 0; j < nt; ++j) {
            const int src = reflect_idx(lo4 + k + j - r, ax.in_n);
            w[src - LO] = fmaf(c4[k], f[j], w[src - LO]);
        }
    *lo_out = LO;
    *n_out = HI - LO + 1;
}

// ---- pass 0: per-launch weight tables ------------------------------------------------------
// table layout for an axis with out_n outputs and T = max_taps: int lo[out_n], int n[out_n],
// float w[out_n][T].  One thread per output index; the tables are a few KB and stay in L2.
//
// Shapes the register-blocked kernel (resize_rb_kernel below) takes get a second block behind the two tables: the
// windows of every group of GV output rows / GH output columns expanded into DENSE weight matrices (zeros where a
// tap does not reach), laid out exactly as a CTA wants them in shared memory, so that the resampler's prologue is a
// straight cp.async copy of its slice instead of a per-CTA scatter with dependent table lookups:
//   int   vbase[ngy]            first source row of row group g            (ngy = tiles_y * TH / GV, padded to whole tiles)
//   float wv[ngy][RV][GV]       wv[g][r][i] = weight of source row vbase[g] + r for output row g * GV + i
//   int   hbase[ngx]            first ibuf column of column group k's span, a multiple of 4, relative to its tile's c0
//   float whs[ngx][GH * 4SH + 4]  whs[k][i * 4SH + c] = weight of ibuf column hbase[k] + c for output column k * GH + i
//   int   tc0[tiles_x], tnq[tiles_x]   first source column of the tile (multiple of 4) and its count of column quads
struct RbPlan {
    int cfg;  // -1: shape not covered
    int GV, RV, GH, SH, TW, TH, pitch, tiles_x, tiles_y;
    int off_vbase, off_wv, off_hbase, off_whs, off_tc0, off_tnq, total_ints;  // in 4-byte words from the block's start
    int smem;
};

__global__ void __launch_bounds__(128) resize_tables_kernel(int mode, AxisSpec ay, AxisSpec ax, int* __restrict__ ty_lo,
                                                            int* __restrict__ tx_lo, RbPlan pl, int* __restrict__ dense) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ay.out_n) {
        float* w = reinterpret_cast<float*>(ty_lo + 2 * ay.out_n) + (size_t)i * ay.max_taps;
        if (mode == OTF_RESIZE_LANCZOS) lanczos_axis_weights(ay, i, ty_lo + i, ty_lo + ay.out_n + i, w);
        else axis_weights(mode, ay, i, ty_lo + i, ty_lo + ay.out_n + i, w);
    } else if (i - ay.out_n < ax.out_n) {
        const int o = i - ay.out_n;
        float* w = reinterpret_cast<float*>(tx_lo + 2 * ax.out_n) + (size_t)o * ax.max_taps;
        if (mode == OTF_RESIZE_LANCZOS) lanczos_axis_weights(ax, o, tx_lo + o, tx_lo + ax.out_n + o, w);
        else axis_weights(mode, ax, o, tx_lo + o, tx_lo + ax.out_n + o, w);
    }
    if (pl.cfg < 0) return;
    // dense block: thread = one (possibly padded) output index; it owns one column of its group's matrix
    constexpr int kMaxT = 16;  // rb shapes have max_taps <= 13
    float w[kMaxT], w2[kMaxT];
    auto window = [&](const AxisSpec& a, int o, int* lo, int* n, float* wp) {
        if (mode == OTF_RESIZE_LANCZOS) lanczos_axis_weights(a, o, lo, n, wp);
        else axis_weights(mode, a, o, lo, n, wp);
    };
    const int rows_pad = pl.tiles_y * pl.TH, cols_pad = pl.tiles_x * pl.TW;
    if (i < rows_pad) {
        const int t = i, g = t / pl.GV, ii = t - g * pl.GV;
        float* col = reinterpret_cast<float*>(dense + pl.off_wv) + (size_t)g * pl.RV * pl.GV + ii;
        for (int r = 0; r < pl.RV; ++r) col[r * pl.GV] = 0.0f;
        int lo0 = 0, n0 = 0;
        if (g * pl.GV < ay.out_n) window(ay, g * pl.GV, &lo0, &n0, w2);
        if (ii == 0) dense[pl.off_vbase + g] = lo0;
        if (t < ay.out_n) {
            int lo, n;
            window(ay, t, &lo, &n, w);
            for (int j = 0; j < n; ++j) {
                const int r = lo - lo0 + j;
                if (r >= 0 && r < pl.RV) col[r * pl.GV] = w[j];
            }
        }
    } else if (i - rows_pad < cols_pad) {
        const int o = i - rows_pad, k = o / pl.GH, ii = o - k * pl.GH, tile = o / pl.TW, SPAN = 4 * pl.SH, WHP = pl.GH * SPAN + 4;
        float* row = reinterpret_cast<float*>(dense + pl.off_whs) + (size_t)k * WHP + ii * SPAN;
        for (int c = 0; c < SPAN; ++c) row[c] = 0.0f;
        if (ii == 0)
            for (int c = 0; c < 4; ++c) row[pl.GH * SPAN + c] = 0.0f;  // (the pad words)
        int lo, n, lot = 0, lok = 0;
        window(ax, min(tile * pl.TW, ax.out_n - 1), &lot, &n, w2);
        const int c0 = lot & ~3;
        window(ax, min(k * pl.GH, ax.out_n - 1), &lok, &n, w2);
        const int hb = (lok - c0) & ~3;
        if (ii == 0) dense[pl.off_hbase + k] = hb;
        if (o < ax.out_n) {
            window(ax, o, &lo, &n, w);
            for (int j = 0; j < n; ++j) {
                const int c = lo - c0 - hb + j;
                if (c >= 0 && c < SPAN) row[c] = w[j];
            }
        }
        if (o == tile * pl.TW) {
            window(ax, min(tile * pl.TW + pl.TW, ax.out_n) - 1, &lo, &n, w2);
            dense[pl.off_tc0 + tile] = c0;
            dense[pl.off_tnq + tile] = min((min(lo + n, ax.in_n) - c0 + 3) >> 2, (pl.pitch - SPAN) >> 2);
        }
    }
}

// ---- fused Gaussian-noise epilogue (row g1: resize + noise + clamp in one launch) ----------------------------
// What otf_gaussian_noise_f32 would do to this launch's output (degradations.py:569-633 after :1004-1021), applied
// to the resampled pixel before it is stored.  The Philox positions are the noise kernel's (noise.cu: quad = four
// consecutive pixels of an output row), so the fused launch is BIT-IDENTICAL to resize followed by the noise kernel.
// Four lanes own four consecutive output columns (one quad) and four output rows each: lane j of the group draws the
// quad of row j (one Philox call, two Box-Muller pairs), a 4x4 exchange inside the group hands every lane the normal
// of its own column for each of the four rows — one Philox call per four pixels, as in the stand-alone kernel.
struct NoiseEpi {
    const float* sigma;          // fp32[B]
    const float* gray;           // fp32[B] or nullptr
    const uint64_t* offset_dev;  // device offset word or nullptr
    uint64_t seed, offset;
    int C, flags;
};

struct NoiseCtx {  // per-CTA constants (a CTA works inside one plane)
    float ca, cb;
    bool need_color, need_gray;
    uint64_t color_base, stream_color, stream_gray;
    int QW, flags;
};

__device__ __forceinline__ NoiseCtx noise_ctx(const NoiseEpi& ne, int plane, int OH, int OW) {
    NoiseCtx nc;
    const int b = plane / ne.C, c = plane - b * ne.C;
    const float sg = ne.sigma[b];
    const float g = ne.gray ? ne.gray[b] : 0.0f;
    const bool use_gray = ne.gray != nullptr;
    const float s255 = __fdiv_rn(sg, 255.0f);
    nc.ca = use_gray ? s255 * __fsub_rn(1.0f, g) : s255;  // the folded factors of gaussian_noise_kernel
    nc.cb = use_gray ? s255 * g : 0.0f;
    nc.need_color = !(use_gray && g == 1.0f);
    nc.need_gray = use_gray && g != 0.0f;
    nc.QW = (OW + 3) >> 2;
    const uint64_t off = ne.offset + (ne.offset_dev ? *ne.offset_dev : 0);
    nc.color_base = (uint64_t)b * ((uint64_t)ne.C * OH * nc.QW) + (uint64_t)c * OH * nc.QW;
    nc.stream_color = off * 8 + STREAM_COLOR;
    nc.stream_gray = off * 8 + STREAM_GRAY;
    nc.flags = ne.flags;
    return nc;
}

// The four noise values of output column x (x - (lane & 3) is a multiple of 4: the group's quad) at rows y[0..3].
// Must be called by all 32 lanes; rows >= OH are skipped (their value is unspecified).
__device__ __forceinline__ void noise_rows4(const Philox& ph, const NoiseCtx& nc, int OH, int x, const int (&y)[4], int lane,
                                            float (&nz)[4]) {
    const int j = lane & 3;
    const int my = j == 0 ? y[0] : j == 1 ? y[1] : j == 2 ? y[2] : y[3];
    const int xq = x >> 2;
    float q[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (my < OH) {
        float4 n = make_float4(0.f, 0.f, 0.f, 0.f), g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nc.need_color) n = normal4(ph, nc.color_base + (uint64_t)my * nc.QW + xq, nc.stream_color);
        if (nc.need_gray) g = normal4(ph, (uint64_t)my * nc.QW + xq, nc.stream_gray);
        q[0] = fmaf(g.x, nc.cb, n.x * nc.ca);  // gaussian_noise_kernel: noise = fmaf(ng, cb, nc * ca)
        q[1] = fmaf(g.y, nc.cb, n.y * nc.ca);
        q[2] = fmaf(g.z, nc.cb, n.z * nc.ca);
        q[3] = fmaf(g.w, nc.cb, n.w * nc.ca);
    }
    // 4x4 exchange: in step s lane j sends component (j + s) & 3 of its row to lane (j + s) & 3, i.e. lane j receives,
    // from lane i = (j - s) & 3, the component j of row i
    float r[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int comp = (j + s) & 3;
        const float send = comp == 0 ? q[0] : comp == 1 ? q[1] : comp == 2 ? q[2] : q[3];
        r[s] = __shfl_sync(0xffffffffu, send, (lane & ~3) | ((j - s) & 3));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int s = (j - i) & 3;
        nz[i] = s == 0 ? r[0] : s == 1 ? r[1] : s == 2 ? r[2] : r[3];
    }
}

__device__ __forceinline__ float noise_finish(float v, float nz, int flags) { return noise_tail(__fadd_rn(v, nz), flags); }

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_f32x4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// ---- pass 1: a CTA produces a (32*CW) x tile_h output tile ----------------------------------
// (1) its slice of the tables and the source rectangle it needs go to shared memory (cp.async:
// every byte in flight at once, no register staging); (2) horizontal pass smem -> smem: a thread owns
// one output column, keeps that column's NT taps in registers and walks down the rows (1 LDS + 1 FFMA
// per tap); (3) vertical pass smem -> global: a warp owns an output row, keeps the row's NT taps in
// registers (broadcast) and each lane produces CW columns; clamp fused.  Taps beyond a window's true
// length are zero and read zero-filled padding, so the unrolled loops need no predicates.
// smem: wx[TW][NT] wy[tile_h][NT] xlo[TW] ylo[tile_h] | src[rows_cap][SP] | tmp[rows_cap + NT][TW]
template <int NT, int CW, bool NOISE>
__global__ void __launch_bounds__(256) resize_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                     AxisSpec ay, AxisSpec ax, const int* __restrict__ ty_lo,
                                                     const int* __restrict__ tx_lo, int tile_h, int rows_cap,
                                                     int cols_cap, int clamp_out, int vec_ok, const __grid_constant__ NoiseEpi ne) {
    pdl_enter();
    constexpr int TW = 32 * CW;
    extern __shared__ __align__(16) float sm[];
    const int SP = (cols_cap + NT + 3) & ~3;  // src row pitch: >= NT zero columns behind every row, 16-byte rows
    float* wx = sm;
    float* wy = wx + TW * NT;
    int* xlo = reinterpret_cast<int*>(wy + tile_h * NT);
    int* ylo = xlo + TW;
    float* src = reinterpret_cast<float*>(ylo + tile_h);
    float* tmp = src + rows_cap * SP;

    const int plane = blockIdx.z;
    const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * tile_h;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tw = min(TW, ax.out_n - ox0), th = min(tile_h, ay.out_n - oy0);
    // windows are monotone in the output index: the rectangle is [first.lo, last.lo + last.n)
    const int row_lo = ty_lo[oy0], row_hi = ty_lo[oy0 + th - 1] + ty_lo[ay.out_n + oy0 + th - 1];
    int col_lo = tx_lo[ox0];
    const int col_hi = tx_lo[ox0 + tw - 1] + tx_lo[ax.out_n + ox0 + tw - 1];
    if (vec_ok) col_lo &= ~3;  // start the rectangle on a 16-byte boundary so whole quads move with one cp.async
    const int nrows = min(row_hi - row_lo, rows_cap), ncols = min(col_hi - col_lo, cols_cap);
    const float* ip = img + (size_t)plane * ay.in_n * ax.in_n + (size_t)row_lo * ax.in_n + col_lo;
    if (vec_ok) {
        const int nquad = (ncols + 3) >> 2;
        for (int r = wid; r < nrows; r += 8) {
            float* srow = src + r * SP;
            const float* grow = ip + (size_t)r * ax.in_n;
            for (int qd = lane; qd < nquad; qd += 32) {
                if (col_lo + 4 * qd + 3 < ax.in_n) {
                    cp_async_f32x4(srow + 4 * qd, grow + 4 * qd);
                } else {
                    for (int k = 0; k < 4; ++k) {
                        if (col_lo + 4 * qd + k < ax.in_n) cp_async_f32(srow + 4 * qd + k, grow + 4 * qd + k);
                        else srow[4 * qd + k] = 0.0f;
                    }
                }
            }
            for (int cidx = 4 * nquad + lane; cidx < SP; cidx += 32) srow[cidx] = 0.0f;
        }
    } else {
        for (int r = wid; r < nrows; r += 8) {
            float* srow = src + r * SP;
            for (int cidx = lane; cidx < ncols; cidx += 32) cp_async_f32(srow + cidx, ip + (size_t)r * ax.in_n + cidx);
            for (int cidx = ncols + lane; cidx < SP; cidx += 32) srow[cidx] = 0.0f;
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // table slices, zero-padded to NT taps (plain loads: tiny, L2 resident; overlap the copies)
    const float* gwx = reinterpret_cast<const float*>(tx_lo + 2 * ax.out_n);
    const float* gwy = reinterpret_cast<const float*>(ty_lo + 2 * ay.out_n);
    for (int i = tid; i < TW * NT; i += 256) {
        const int o = i / NT, j = i - o * NT;
        wx[i] = (o < tw && j < ax.max_taps) ? gwx[(size_t)(ox0 + o) * ax.max_taps + j] : 0.0f;
    }
    for (int i = tid; i < tile_h * NT; i += 256) {
        const int o = i / NT, j = i - o * NT;
        wy[i] = (o < th && j < ay.max_taps) ? gwy[(size_t)(oy0 + o) * ay.max_taps + j] : 0.0f;
    }
    for (int i = tid; i < TW; i += 256) xlo[i] = tx_lo[ox0 + min(i, tw - 1)] - col_lo;
    for (int i = tid; i < tile_h; i += 256) ylo[i] = ty_lo[oy0 + min(i, th - 1)] - row_lo;
    for (int i = tid; i < NT * TW; i += 256) tmp[nrows * TW + i] = 0.0f;  // zero rows behind the last one
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // horizontal pass: thread = (column, row phase)
    {
        constexpr int PH = 256 / TW;  // row phases: 8, 4 or 2
        const int col = tid % TW, ph = tid / TW;
        float w[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) w[j] = wx[col * NT + j];
        const float* sp = src + xlo[col];
        for (int r = ph; r < nrows; r += PH) {
            const float* rp = sp + r * SP;
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < NT; ++j) acc = fmaf(w[j], rp[j], acc);
            tmp[r * TW + col] = acc;
        }
    }
    __syncthreads();
    // vertical pass: warp = output row, lane = CW columns
    float* op = out + (size_t)plane * ay.out_n * ax.out_n;
    if (NOISE) {
        // four output rows per trip (t, t + 8, t + 16, t + 24: tile_h <= 32) so that the noise epilogue can share one
        // Philox call between the four lanes of a quad
        const Philox ph(ne.seed);
        const NoiseCtx nc = noise_ctx(ne, plane, ay.out_n, ax.out_n);
        float acc[4][CW];
        int yy[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = wid + 8 * i;
            yy[i] = t < th ? oy0 + t : ay.out_n;  // (out_n = "no such row")
#pragma unroll
            for (int k = 0; k < CW; ++k) acc[i][k] = 0.0f;
            if (t < th) {
                const float* tp = tmp + ylo[t] * TW + lane;
#pragma unroll
                for (int ii = 0; ii < NT; ++ii) {
                    const float wv = wy[t * NT + ii];
#pragma unroll
                    for (int k = 0; k < CW; ++k) acc[i][k] = fmaf(wv, tp[ii * TW + 32 * k], acc[i][k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < CW; ++k) {
            const int x = ox0 + lane + 32 * k;
            float nz[4];
            noise_rows4(ph, nc, ay.out_n, x, yy, lane, nz);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (yy[i] < ay.out_n && x < ax.out_n)
                    op[(size_t)yy[i] * ax.out_n + x] = noise_finish(clamp_out ? clamp01(acc[i][k]) : acc[i][k], nz[i], nc.flags);
        }
        return;
    }
    for (int t = wid; t < th; t += 8) {
        float w[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) w[i] = wy[t * NT + i];
        const float* tp = tmp + ylo[t] * TW + lane;
        float acc[CW];
#pragma unroll
        for (int k = 0; k < CW; ++k) acc[k] = 0.0f;
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int k = 0; k < CW; ++k) acc[k] = fmaf(w[i], tp[i * TW + 32 * k], acc[k]);
#pragma unroll
        for (int k = 0; k < CW; ++k) {
            const int x = ox0 + lane + 32 * k;
            if (x < ax.out_n) op[(size_t)(oy0 + t) * ax.out_n + x] = clamp_out ? clamp01(acc[k]) : acc[k];
        }
    }
}

// ---- pass 1 (v4): vertical pass first, straight from global memory ---------------------------------
// The tiled kernel above spends most of its issue slots on staging (source rectangle + table slices into
// shared memory, two barriers, per-element index math).  This kernel keeps only what the arithmetic needs:
//   (V) thread = a quad (VEC) or one (scalar) source column of the tile's column span and an output row:
//       NT coalesced row loads through L1 (a source row is re-read by the ~NT/scale output rows whose windows
//       cover it — L1 hits), 4*NT FMAs against the row's taps (shared-memory broadcast), one STS.128 into
//       vbuf[TH][pitch].  No source staging, no vertical halo in shared memory.
//   (H) thread = an output column with its NT taps in registers, walking down the TH rows of vbuf:
//       NT LDS + NT FFMA per output, clamp fused, coalesced 128 B stores.
// NT == 0 selects run-time tap counts (long windows of extreme down-scales) with the horizontal taps in smem.
// Rounding order differs from ATen's (W then H) by ~1e-7, inside the 1e-5 bar of the path.
template <int NT, bool VEC, bool NOISE>
__global__ void __launch_bounds__(256) resize_vh_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                        AxisSpec ay, AxisSpec ax, const int* __restrict__ ty_lo,
                                                        const int* __restrict__ tx_lo, int TW, int TH, int pitch,
                                                        int clamp_out, const __grid_constant__ NoiseEpi ne) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];
    const int nty = NT ? NT : ay.max_taps, ntx = NT ? NT : ax.max_taps;
    const int wxp = ntx | 1;  // odd pitch of the run-time horizontal tap table: conflict-free column reads
    float* vbuf = sm;                                   // [TH][pitch]
    float* wy = vbuf + (size_t)TH * pitch;              // [TH][nty]
    int* ylo = reinterpret_cast<int*>(wy + TH * nty);   // [TH]
    int* xlo = ylo + TH;                                // [TW]
    float* wxs = reinterpret_cast<float*>(xlo + TW);    // [TW][wxp]

    const int plane = blockIdx.z;
    const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
    const int tid = threadIdx.x;
    const int tw = min(TW, ax.out_n - ox0), th = min(TH, ay.out_n - oy0);
    int col_lo = tx_lo[ox0];
    const int col_hi = tx_lo[ox0 + tw - 1] + tx_lo[ax.out_n + ox0 + tw - 1];
    if (VEC) col_lo &= ~3;
    const int ncols = min(col_hi - col_lo, pitch - (NT ? NT : ax.max_taps));  // host sizes pitch from a bound on the span
    const int nunit = VEC ? (ncols + 3) >> 2 : ncols;   // work items per output row
    const int ufill = VEC ? 4 * nunit : nunit;
    // Prologue without per-element divisions (it used to be 40 % of the kernel's instructions): every loop below
    // maps threads onto a fixed 2-D shape.
    {   // pull the tile's source rectangle towards L1 now: the vertical pass then finds its rows there instead of
        // paying one DRAM round trip per output-row iteration (one 128-byte line per prefetch)
        const int row_lo = ty_lo[oy0], row_hi = min(ty_lo[oy0 + th - 1] + ty_lo[ay.out_n + oy0 + th - 1], ay.in_n);
        const float* base = img + (size_t)plane * ay.in_n * ax.in_n + col_lo;
        const int lpr = (ncols * 4 + 127) / 128 + 1;  // lines per row (unaligned start)
        for (int r = row_lo + (tid >> 3); r < row_hi; r += 32)
            for (int l = tid & 7; l < lpr; l += 8)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (size_t)r * ax.in_n + min(32 * l, ncols - 1)));
    }
    const float* gwy = reinterpret_cast<const float*>(ty_lo + 2 * ay.out_n);
    const float* gwx = reinterpret_cast<const float*>(tx_lo + 2 * ax.out_n);
    if (NT) {
        // vertical taps: thread = (row o, tap j); horizontal taps: thread = (column o, a run of taps)
        constexpr int NTc = NT ? NT : 1;
        for (int i = tid; i < th * NTc; i += 256) {
            const int o = i / NTc, j = i % NTc;
            wy[i] = j < ay.max_taps ? __ldg(gwy + (size_t)(oy0 + o) * ay.max_taps + j) : 0.0f;
        }
        const int o = tid & (TW - 1), part = tid / TW, parts = 256 / TW;  // TW is a power of two
        if (o < tw) {
            const float* gp = gwx + (size_t)(ox0 + o) * ax.max_taps;
            for (int j = part; j < NTc; j += parts) wxs[o * wxp + j] = j < ax.max_taps ? __ldg(gp + j) : 0.0f;
        }
    } else {
        for (int i = tid; i < th * nty; i += 256) wy[i] = __ldg(gwy + (size_t)oy0 * ay.max_taps + i);  // nty == max_taps: contiguous
        for (int i = tid; i < tw * ntx; i += 256) {
            const int o = i / ntx, j = i - o * ntx;
            wxs[o * wxp + j] = __ldg(gwx + (size_t)ox0 * ax.max_taps + i);
        }
    }
    for (int i = tid; i < th; i += 256) ylo[i] = ty_lo[oy0 + i];
    for (int i = tid; i < tw; i += 256) xlo[i] = tx_lo[ox0 + i] - col_lo;
    // columns behind the span: padded taps (weight 0) of the last output columns read them
    for (int r = tid >> 4; r < th; r += 16)
        for (int c = ufill + (tid & 15); c < pitch; c += 16) vbuf[r * pitch + c] = 0.0f;
    __syncthreads();
    // ---- vertical pass: global -> vbuf ----
    {
        const int upar = min(nunit, 256), rpar = 256 / upar;
        const int u0 = tid % upar, tph = tid / upar;
        const float* ip = img + (size_t)plane * ay.in_n * ax.in_n + col_lo;
        const int W = ax.in_n, Hm1 = ay.in_n - 1;
        if (tph < rpar) {
            for (int t = tph; t < th; t += rpar) {
                const int r0 = ylo[t];
                const float* wt = wy + t * nty;
                // rows past the image only carry zero weights: step the row pointer by W while inside, then hold it
                const bool inside = r0 + nty - 1 <= Hm1;
                for (int u = u0; u < nunit; u += upar) {
                    if (VEC) {
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float* cp = ip + (size_t)r0 * W + 4 * u;
                        if (NT) {
                            float4 v[NT ? NT : 1];
                            if (inside) {
#pragma unroll
                                for (int i = 0; i < NT; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(cp + i * W));
                            } else {
#pragma unroll
                                for (int i = 0; i < NT; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(cp + (min(r0 + i, Hm1) - r0) * W));
                            }
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                const float w = wt[i];
                                acc.x = fmaf(w, v[i].x, acc.x); acc.y = fmaf(w, v[i].y, acc.y);
                                acc.z = fmaf(w, v[i].z, acc.z); acc.w = fmaf(w, v[i].w, acc.w);
                            }
                        } else {
                            for (int i = 0; i < nty; ++i) {
                                const float4 v = __ldg(reinterpret_cast<const float4*>(cp + (min(r0 + i, Hm1) - r0) * W));
                                const float w = wt[i];
                                acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                                acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                            }
                        }
                        *reinterpret_cast<float4*>(vbuf + t * pitch + 4 * u) = acc;
                    } else {
                        float acc = 0.0f;
                        const float* cp = ip + (size_t)r0 * W + u;
                        if (NT) {
                            float v[NT ? NT : 1];
                            if (inside) {
#pragma unroll
                                for (int i = 0; i < NT; ++i) v[i] = __ldg(cp + i * W);
                            } else {
#pragma unroll
                                for (int i = 0; i < NT; ++i) v[i] = __ldg(cp + (min(r0 + i, Hm1) - r0) * W);
                            }
#pragma unroll
                            for (int i = 0; i < NT; ++i) acc = fmaf(wt[i], v[i], acc);
                        } else {
                            for (int i = 0; i < nty; ++i) acc = fmaf(wt[i], __ldg(cp + (min(r0 + i, Hm1) - r0) * W), acc);
                        }
                        vbuf[t * pitch + u] = acc;
                    }
                }
            }
        }
    }
    __syncthreads();
    // ---- horizontal pass: vbuf -> global ----
    if (NOISE) {
        // the same sums, four output rows per trip (rows ph + (4g + i) * PH) so that the noise epilogue can share one
        // Philox call between the four lanes of a quad; every lane takes part in the exchange, stores are predicated
        const int col = tid % TW, ph = tid / TW, PH = 256 / TW;
        const int x = ox0 + col, lane = tid & 31;
        const bool colok = col < tw;
        const Philox phx(ne.seed);
        const NoiseCtx nc = noise_ctx(ne, plane, ay.out_n, ax.out_n);
        float* op = out + (size_t)plane * ay.out_n * ax.out_n + x;
        const float* vp = vbuf + (colok ? xlo[col] : 0);
        float w[NT ? NT : 1];
        if (NT) {
#pragma unroll
            for (int j = 0; j < (NT ? NT : 1); ++j) w[j] = colok ? wxs[col * wxp + j] : 0.0f;
        }
        const float* wp = wxs + (colok ? col : 0) * wxp;
        for (int t0 = ph; t0 < th; t0 += 4 * PH) {
            float acc[4];
            int yy[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = t0 + i * PH;
                yy[i] = t < th ? oy0 + t : ay.out_n;
                acc[i] = 0.0f;
                if (t < th && colok) {
                    const float* rp = vp + t * pitch;
                    if (NT) {
#pragma unroll
                        for (int j = 0; j < (NT ? NT : 1); ++j) acc[i] = fmaf(w[j], rp[j], acc[i]);
                    } else {
                        for (int j = 0; j < ntx; ++j) acc[i] = fmaf(wp[j], rp[j], acc[i]);
                    }
                }
            }
            float nz[4];
            noise_rows4(phx, nc, ay.out_n, x, yy, lane, nz);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (yy[i] < ay.out_n && colok)
                    op[(size_t)yy[i] * ax.out_n] = noise_finish(clamp_out ? clamp01(acc[i]) : acc[i], nz[i], nc.flags);
        }
    } else {
        const int col = tid % TW, ph = tid / TW, PH = 256 / TW;
        const int x = ox0 + col;
        if (col < tw) {
            float* op = out + (size_t)plane * ay.out_n * ax.out_n + (size_t)oy0 * ax.out_n + x;
            const float* vp = vbuf + xlo[col];
            if (NT) {
                float w[NT ? NT : 1];
#pragma unroll
                for (int j = 0; j < NT; ++j) w[j] = wxs[col * wxp + j];
                for (int t = ph; t < th; t += PH) {
                    const float* rp = vp + t * pitch;
                    float acc = 0.0f;
#pragma unroll
                    for (int j = 0; j < NT; ++j) acc = fmaf(w[j], rp[j], acc);
                    op[(size_t)t * ax.out_n] = clamp_out ? clamp01(acc) : acc;
                }
            } else {
                const float* wp = wxs + col * wxp;
                for (int t = ph; t < th; t += PH) {
                    const float* rp = vp + t * pitch;
                    float acc = 0.0f;
                    for (int j = 0; j < ntx; ++j) acc = fmaf(wp[j], rp[j], acc);
                    op[(size_t)t * ax.out_n] = clamp_out ? clamp01(acc) : acc;
                }
            }
        }
    }
}

// ---- pass 1 (v6): register-blocked dense windows ---------------------------------------------------------------
// Both passes of the vertical-first kernel above pay one load per tap per output (NT LDG.128 per intermediate quad,
// NT LDS per output) plus the index arithmetic around them: ~130 warp instructions per 32 outputs at 256->192, issue-
// and latency-bound at a quarter of the HBM roof.  Here a task produces a BLOCK of outputs along the resampled axis
// from the union of their windows, with the block's windows expanded into a small dense weight matrix (zeros where a
// tap does not reach): every source value is loaded once per block and feeds all the block's outputs from registers.
//   (V) task = (source column quad, group of GV output rows): RV coalesced LDG.128 straight from global memory (all
//       independent: one round trip per task), RV broadcast LDS.128 of the group's weights [r][GV], GV*RV FMA4,
//       GV STS.128 into ibuf[TH][pitch];
//   (H) thread = a group of GH output columns with its dense GH x 4SH weights in REGISTERS, walking down the tile's
//       rows: SH LDS.128 + GH*4SH FMAs per GH outputs, then clamp (+ the Gaussian-noise epilogue: with GH = 4 the
//       thread's outputs are exactly one Philox quad of the stand-alone noise kernel) and one 16-byte store.
// The zero taps cost FMAs (~1.6x the minimum) — the FMA pipe is nowhere near busy here — and buy a 3x cut in load
// instructions and no per-tap index arithmetic at all.  Summation order inside a window is unchanged (taps in
// ascending source order, zeros in between are exact no-ops), so results equal the vertical-first kernel's bit for bit.
template <int GV, int RV, int GH, int SH, bool VEC, bool NOISE>
#ifndef OTF_RB_MINB
#define OTF_RB_MINB 4
#endif
__global__ void __launch_bounds__(128, OTF_RB_MINB) resize_rb_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                           AxisSpec ay, AxisSpec ax, const int* __restrict__ dense,
                                                           const __grid_constant__ RbPlan pl, int clamp_out, int vec_out,
                                                           const __grid_constant__ NoiseEpi ne) {
    pdl_enter();
    constexpr int SPAN = 4 * SH, WHP = GH * SPAN + 4;  // (+4: consecutive column groups start 4 banks apart)
    extern __shared__ __align__(16) float sm[];
    const int TW = pl.TW, TH = pl.TH, pitch = pl.pitch;
    const int NG = TH / GV, KP = TW / GH;
    float* ibuf = sm;                                   // [TH][pitch]  vertical-pass result
    float* wv = ibuf + (size_t)TH * pitch;              // [NG][RV][GV] dense vertical weights of each row group
    float* whs = wv + NG * RV * GV;                     // [KP][WHP]    dense horizontal weights of each column group
    int* hbase = reinterpret_cast<int*>(whs + KP * WHP);  // [KP] first ibuf column of the group's span (multiple of 4)

    const int plane = blockIdx.z;
    const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
    const int tid = threadIdx.x;
    const int tw = min(TW, ax.out_n - ox0), th = min(TH, ay.out_n - oy0);
    const int H = ay.in_n, W = ax.in_n;
    // this CTA's slices of the dense block -> shared memory, asynchronously (16-byte pieces; every slice starts on a
    // 16-byte boundary: TH / GV * RV * GV, TW / GH * WHP and TW / GH words are multiples of 4)
    {
        const float* gwv = reinterpret_cast<const float*>(dense + pl.off_wv) + (size_t)blockIdx.y * NG * RV * GV;
        const float* gwh = reinterpret_cast<const float*>(dense + pl.off_whs) + (size_t)blockIdx.x * KP * WHP;
        const float* ghb = reinterpret_cast<const float*>(dense + pl.off_hbase) + (size_t)blockIdx.x * KP;
        for (int i = tid; i < NG * RV * GV / 4; i += 128) cp_async_f32x4(wv + 4 * i, gwv + 4 * i);
        for (int i = tid; i < KP * WHP / 4; i += 128) cp_async_f32x4(whs + 4 * i, gwh + 4 * i);
        for (int i = tid; i < KP / 4; i += 128) cp_async_f32x4(reinterpret_cast<float*>(hbase) + 4 * i, ghb + 4 * i);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int c0 = __ldg(dense + pl.off_tc0 + blockIdx.x), nq = __ldg(dense + pl.off_tnq + blockIdx.x);
    const int* gvb = dense + pl.off_vbase + blockIdx.y * NG;
    // columns behind the span: zero-weight taps of the last column groups read them (SPAN words per row suffice)
    for (int i = tid; i < TH * SH; i += 128) {
        const int r = i / SH, c = i - r * SH;
        *reinterpret_cast<float4*>(ibuf + (size_t)r * pitch + 4 * (nq + c)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- vertical pass: global -> ibuf ----
    {
        const float* ip = img + (size_t)plane * H * W;
        const int ntask = nq * ((th + GV - 1) / GV);
        int task = tid;
        float4 v[RV];
        auto load = [&](int tk) {
            const int g = tk / nq, q = tk - g * nq;
            const int r0 = __ldg(gvb + g), col = c0 + 4 * q;
            const int last = H - 1 - r0;  // rows past the image only meet zero weights: hold the last row
            const float* p0 = ip + (size_t)r0 * W + col;
#pragma unroll
            for (int r = 0; r < RV; ++r) {
                const float* p = p0 + min(r, last) * W;
                if (VEC) {
                    v[r] = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    v[r].x = col < W ? __ldg(p) : 0.0f;
                    v[r].y = col + 1 < W ? __ldg(p + 1) : 0.0f;
                    v[r].z = col + 2 < W ? __ldg(p + 2) : 0.0f;
                    v[r].w = col + 3 < W ? __ldg(p + 3) : 0.0f;
                }
            }
        };
        if (task < ntask) load(task);  // in flight while the weights arrive
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        while (task < ntask) {
            const int g = task / nq, q = task - g * nq;
            const float* wg = wv + g * RV * GV;
            float4 acc[GV];
#pragma unroll
            for (int i = 0; i < GV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < RV; ++r) {
                float w[GV];
                if (GV == 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wg + r * GV);
                    w[0] = w4.x; w[1] = w4.y; w[2 % GV] = w4.z; w[3 % GV] = w4.w;
                } else {
                    const float2 w2 = *reinterpret_cast<const float2*>(wg + r * GV);
                    w[0] = w2.x; w[1] = w2.y;
                }
#pragma unroll
                for (int i = 0; i < GV; ++i) {
                    acc[i].x = fmaf(w[i], v[r].x, acc[i].x); acc[i].y = fmaf(w[i], v[r].y, acc[i].y);
                    acc[i].z = fmaf(w[i], v[r].z, acc[i].z); acc[i].w = fmaf(w[i], v[r].w, acc[i].w);
                }
            }
#pragma unroll
            for (int i = 0; i < GV; ++i) *reinterpret_cast<float4*>(ibuf + (size_t)(g * GV + i) * pitch + 4 * q) = acc[i];
            task += 128;
            if (task < ntask) load(task);
        }
    }
    __syncthreads();
    // ---- horizontal pass: ibuf -> global ----
    const int k = tid & (KP - 1), ph = tid / KP, PH = 128 / KP;
    const int x0 = ox0 + k * GH;
    if (k * GH >= tw) return;
    float wr[GH][SPAN];
#pragma unroll
    for (int i = 0; i < GH; ++i)
#pragma unroll
        for (int s4 = 0; s4 < SH; ++s4) {
            const float4 t4 = *reinterpret_cast<const float4*>(whs + k * WHP + i * SPAN + 4 * s4);
            wr[i][4 * s4] = t4.x; wr[i][4 * s4 + 1] = t4.y; wr[i][4 * s4 + 2] = t4.z; wr[i][4 * s4 + 3] = t4.w;
        }
    const float* rp0 = ibuf + hbase[k];
    float* op = out + (size_t)plane * ay.out_n * ax.out_n + x0;
    const Philox phx(NOISE ? ne.seed : 0);
    NoiseCtx nc;
    if (NOISE) nc = noise_ctx(ne, plane, ay.out_n, ax.out_n);
    for (int t = ph; t < th; t += PH) {
        float in[SPAN];
#pragma unroll
        for (int s4 = 0; s4 < SH; ++s4) {
            const float4 t4 = *reinterpret_cast<const float4*>(rp0 + (size_t)t * pitch + 4 * s4);
            in[4 * s4] = t4.x; in[4 * s4 + 1] = t4.y; in[4 * s4 + 2] = t4.z; in[4 * s4 + 3] = t4.w;
        }
        float acc[GH];
#pragma unroll
        for (int i = 0; i < GH; ++i) {
            float a = 0.0f;
#pragma unroll
            for (int c = 0; c < SPAN; ++c) a = fmaf(wr[i][c], in[c], a);
            acc[i] = clamp_out ? clamp01(a) : a;
        }
        const int y = oy0 + t;
        if (NOISE) {
            // the quad of gaussian_noise_kernel that holds these outputs: x0 is a multiple of GH, GH divides 4
            float4 n = make_float4(0.f, 0.f, 0.f, 0.f), gq = make_float4(0.f, 0.f, 0.f, 0.f);
            const int xq = x0 >> 2;
            if (nc.need_color) n = normal4(phx, nc.color_base + (uint64_t)y * nc.QW + xq, nc.stream_color);
            if (nc.need_gray) gq = normal4(phx, (uint64_t)y * nc.QW + xq, nc.stream_gray);
            const float nn[4] = {n.x, n.y, n.z, n.w}, gg[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
            for (int i = 0; i < GH; ++i) {
                const int comp = GH == 4 ? i : ((x0 & 3) + i);
                float nv = 0.0f, gv = 0.0f;
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    if (m == comp) { nv = nn[m]; gv = gg[m]; }
                acc[i] = noise_finish(acc[i], fmaf(gv, nc.cb, nv * nc.ca), nc.flags);
            }
        }
        float* orow = op + (size_t)y * ax.out_n;
        if (vec_out) {
            if (GH == 4) *reinterpret_cast<float4*>(orow) = make_float4(acc[0], acc[1], acc[2 % GH], acc[3 % GH]);
            else *reinterpret_cast<float2*>(orow) = make_float2(acc[0], acc[1]);
        } else {
#pragma unroll
            for (int i = 0; i < GH; ++i)
                if (x0 + i < ax.out_n) orow[i] = acc[i];
        }
    }
}

// Fallback for extreme down-scales (> 64 taps per output on an axis): one thread per output pixel,
// horizontal sums nested inside the vertical sum, everything straight from L1/L2.  Correct, not fast.
template <bool NOISE>
__global__ void __launch_bounds__(256) resize_generic_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                             AxisSpec ay, AxisSpec ax, const int* __restrict__ ty_lo,
                                                             const int* __restrict__ tx_lo, int clamp_out,
                                                             const __grid_constant__ NoiseEpi ne) {
    pdl_enter();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= ax.out_n || y >= ay.out_n) return;
    const float* wx = reinterpret_cast<const float*>(tx_lo + 2 * ax.out_n) + (size_t)x * ax.max_taps;
    const float* wy = reinterpret_cast<const float*>(ty_lo + 2 * ay.out_n) + (size_t)y * ay.max_taps;
    const int xl = tx_lo[x], xn = tx_lo[ax.out_n + x], yl = ty_lo[y], yn = ty_lo[ay.out_n + y];
    const float* ip = img + (size_t)blockIdx.z * ay.in_n * ax.in_n;
    float acc = 0.0f;
    for (int i = 0; i < yn; ++i) {
        const float* rp = ip + (size_t)(yl + i) * ax.in_n + xl;
        float h = 0.0f;
        for (int j = 0; j < xn; ++j) h = fmaf(wx[j], __ldg(rp + j), h);
        acc = fmaf(wy[i], h, acc);
    }
    float v = clamp_out ? clamp01(acc) : acc;
    if (NOISE) {  // (rare path: every thread draws its own quad and keeps one component)
        const Philox ph(ne.seed);
        const NoiseCtx nc = noise_ctx(ne, blockIdx.z, ay.out_n, ax.out_n);
        float4 n = make_float4(0.f, 0.f, 0.f, 0.f), g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nc.need_color) n = normal4(ph, nc.color_base + (uint64_t)y * nc.QW + (x >> 2), nc.stream_color);
        if (nc.need_gray) g = normal4(ph, (uint64_t)y * nc.QW + (x >> 2), nc.stream_gray);
        const int k = x & 3;
        const float nn = k == 0 ? n.x : k == 1 ? n.y : k == 2 ? n.z : n.w, gg = k == 0 ? g.x : k == 1 ? g.y : k == 2 ? g.z : g.w;
        v = noise_finish(v, fmaf(gg, nc.cb, nn * nc.ca), nc.flags);
    }
    out[(size_t)blockIdx.z * ay.out_n * ax.out_n + (size_t)y * ax.out_n + x] = v;
}

static AxisSpec make_axis(int mode, int in_n, int out_n) {
    AxisSpec a;
    a.in_n = in_n;
    a.out_n = out_n;
    a.scale = (float)in_n / (float)out_n;
    a.support = 0.0f;
    if (mode == OTF_RESIZE_BILINEAR_AA || mode == OTF_RESIZE_BICUBIC_AA) {
        const float interp = mode == OTF_RESIZE_BILINEAR_AA ? 2.0f : 4.0f;
        a.support = a.scale >= 1.0f ? (interp * 0.5f) * a.scale : interp * 0.5f;
        a.max_taps = (int)ceilf(a.support) * 2 + 1;
    } else if (mode == OTF_RESIZE_AREA) {
        a.max_taps = (in_n + out_n - 1) / out_n + 1;
    } else if (mode == OTF_RESIZE_NEAREST_EXACT || mode == OTF_RESIZE_NEAREST) {
        a.max_taps = 1;
    } else if (mode == OTF_RESIZE_LANCZOS && out_n < in_n) {
        const int n = (int)ceil(3.0 / ((double)out_n / (double)in_n) + 1.0);  // prefilter: 2n - 3 taps, radius n - 2
        a.max_taps = 4 + 2 * (n - 2);
    } else {
        a.max_taps = 4;
    }
    return a;
}
static bool lanczos_ok(int in_n, int out_n) {
    if (out_n >= in_n) return true;
    const int r = (int)ceil(3.0 / ((double)out_n / (double)in_n) + 1.0) - 2;
    return r <= kLanczosMaxR && r < in_n;  // reflect padding needs r < extent (F.pad raises otherwise)
}

static size_t table_ints(const AxisSpec& a) { return (size_t)a.out_n * (2 + a.max_taps); }

// source extent covered by nout consecutive outputs (an upper bound)
static int span(const AxisSpec& a, int nout) {
    int v = (int)ceilf(a.scale * (float)(nout - 1)) + a.max_taps + 2;
    return v > a.in_n ? a.in_n : v;
}

// Which register-blocked configuration (if any) takes this shape, its tile and the layout of its dense block.
// Window bound for G consecutive outputs: lo(o + G - 1) - lo(o) <= floor((G - 1) * scale) + 1 and a window has at most
// max_taps entries; the horizontal span additionally starts on a multiple of 4 (up to 3 words of slack).
static RbPlan rb_plan(const AxisSpec& ay, const AxisSpec& ax) {
    RbPlan p;
    memset(&p, 0, sizeof(p));
    p.cfg = -1;
    auto need = [](const AxisSpec& a, int G) { return (int)floorf((float)(G - 1) * a.scale + 1e-3f) + 2 + a.max_taps; };
    if (need(ay, 4) <= 10 && need(ax, 4) + 3 <= 12) { p.cfg = 0; p.GV = 4; p.RV = 10; p.GH = 4; p.SH = 3; }
    else if (need(ay, 4) <= 13 && need(ax, 4) + 3 <= 16) { p.cfg = 1; p.GV = 4; p.RV = 13; p.GH = 4; p.SH = 4; }
    else if (need(ay, 2) <= 12 && need(ax, 2) + 3 <= 16) { p.cfg = 2; p.GV = 2; p.RV = 12; p.GH = 2; p.SH = 4; }
    else return p;
    const int OH = ay.out_n, OW = ax.out_n, WHP = p.GH * 4 * p.SH + 4;
    p.TW = OW >= 96 ? 128 : OW >= 48 ? 64 : 32;
    p.TH = OH >= 24 ? 32 : 16;
    for (;;) {
        p.pitch = ((span(ax, p.TW) + 3 + 3) & ~3) + 4 * p.SH;
        p.smem = (p.TH * p.pitch + (p.TH / p.GV) * p.RV * p.GV + (p.TW / p.GH) * WHP + p.TW / p.GH) * 4;
        if (p.smem <= 44 * 1024 || p.TW == 32) break;
        p.TW /= 2;
    }
    if (p.smem > 100 * 1024) { p.cfg = -1; return p; }
    p.tiles_x = ceil_div(OW, p.TW);
    p.tiles_y = ceil_div(OH, p.TH);
    const int ngy = p.tiles_y * (p.TH / p.GV), ngx = p.tiles_x * (p.TW / p.GH);
    int off = 0;
    auto take = [&](int words) { const int o = off; off += (words + 3) & ~3; return o; };  // every array 16-byte aligned
    p.off_vbase = take(ngy);
    p.off_wv = take(ngy * p.RV * p.GV);
    p.off_hbase = take(ngx);
    p.off_whs = take(ngx * WHP);
    p.off_tc0 = take(p.tiles_x);
    p.off_tnq = take(p.tiles_x);
    p.total_ints = off;
    return p;
}
static size_t sparse_ints(const AxisSpec& ay, const AxisSpec& ax) { return (table_ints(ay) + table_ints(ax) + 3) & ~(size_t)3; }
static int launch_tables(int mode, const AxisSpec& ay, const AxisSpec& ax, int* ws, cudaStream_t st) {
    const RbPlan pl = rb_plan(ay, ax);
    int n = ay.out_n + ax.out_n;
    if (pl.cfg >= 0 && pl.tiles_y * pl.TH + pl.tiles_x * pl.TW > n) n = pl.tiles_y * pl.TH + pl.tiles_x * pl.TW;
    resize_tables_kernel<<<ceil_div(n, 128), 128, 0, st>>>(mode, ay, ax, ws, ws + table_ints(ay), pl, ws + sparse_ints(ay, ax));
    OTF_LAUNCH_CHECK("resize_tables_kernel");
    return OTF_OK;
}

}  // namespace otf

extern "C" int64_t otf_resize_workspace_bytes(int H, int W, int OH, int OW, int mode) {
    using namespace otf;
    if (H <= 0 || W <= 0 || OH <= 0 || OW <= 0 || mode < OTF_RESIZE_BILINEAR_AA || mode > OTF_RESIZE_LANCZOS) return -1;
    if (mode == OTF_RESIZE_LANCZOS && !(lanczos_ok(H, OH) && lanczos_ok(W, OW))) return -1;
    const AxisSpec ay = make_axis(mode, H, OH), ax = make_axis(mode, W, OW);
    const RbPlan pl = rb_plan(ay, ax);
    return (int64_t)(sparse_ints(ay, ax) + (pl.cfg >= 0 ? (size_t)pl.total_ints : 0)) * 4;
}

extern "C" int otf_resize_tables_f32(int H, int W, int OH, int OW, int mode, void* workspace_dev, int64_t workspace_bytes,
                                     void* stream) {
    using namespace otf;
    OTF_REQUIRE(workspace_dev, OTF_ERR_BAD_ARG, "resize_tables: null workspace");
    const int64_t need = otf_resize_workspace_bytes(H, W, OH, OW, mode);
    OTF_REQUIRE(need > 0, OTF_ERR_BAD_ARG, "resize_tables: mode %d or extents (%d, %d) -> (%d, %d) not supported", mode, H, W, OH, OW);
    OTF_REQUIRE(workspace_bytes >= need, OTF_ERR_WORKSPACE, "resize_tables: workspace too small");
    const AxisSpec ay = make_axis(mode, H, OH), ax = make_axis(mode, W, OW);
    return launch_tables(mode, ay, ax, (int*)workspace_dev, (cudaStream_t)stream);
}

namespace otf {
static int resize_impl(const float* img, int planes, int H, int W, float* out, int OH, int OW, int mode,
                       int clamp_out, void* workspace_dev, int64_t workspace_bytes, int tables_ready,
                       void* stream, const NoiseEpi* nep) {
    NoiseEpi ne;
    memset(&ne, 0, sizeof(ne));
    if (nep) ne = *nep;
    const bool noise = nep != nullptr;

    OTF_REQUIRE(img && out && img != out && workspace_dev, OTF_ERR_BAD_ARG, "resize: bad pointers");
    OTF_REQUIRE(planes > 0 && planes <= 65535 && H > 0 && W > 0 && OH > 0 && OW > 0, OTF_ERR_BAD_ARG, "resize: bad extents");
    OTF_REQUIRE(mode >= OTF_RESIZE_BILINEAR_AA && mode <= OTF_RESIZE_LANCZOS, OTF_ERR_BAD_ARG, "resize: unknown mode %d", mode);
    OTF_REQUIRE(mode != OTF_RESIZE_LANCZOS || (lanczos_ok(H, OH) && lanczos_ok(W, OW)), OTF_ERR_UNSUPPORTED,
                "resize: lanczos prefilter radius above %d (or not below the extent)", kLanczosMaxR);
    OTF_REQUIRE(workspace_bytes >= otf_resize_workspace_bytes(H, W, OH, OW, mode), OTF_ERR_WORKSPACE, "resize: workspace too small");
    const AxisSpec ay = make_axis(mode, H, OH), ax = make_axis(mode, W, OW);
    int* ty_lo = (int*)workspace_dev;
    int* tx_lo = ty_lo + table_ints(ay);
    cudaStream_t st = (cudaStream_t)stream;
    if (!tables_ready) {  // the tables depend on (H, W, OH, OW, mode) only: a caller may keep and reuse them
        if (int rc = launch_tables(mode, ay, ax, ty_lo, st)) return rc;
    }
    const int mt = ay.max_taps > ax.max_taps ? ay.max_taps : ax.max_taps;
    // Two tiled kernels: the vertical-first kernel wins when the image shrinks (long windows, output smaller than the
    // source: 256->102 bicubic 0.037 vs 0.072 ms), the staged horizontal-first kernel when it grows (0.052 vs 0.063 ms
    // at 256->384 bilinear).  OTF_RESIZE_IMPL=3|4 forces one of them for A/B runs.
    static const int forced = [] { const char* e = getenv("OTF_RESIZE_IMPL"); return e ? atoi(e) : 0; }();
    const int impl = forced ? forced : (ay.scale >= 1.0f && ax.scale >= 1.0f ? 4 : 3);  // (the fallbacks' choice; 6 = rb only where it fits)
    // The register-blocked kernel takes every shape whose block windows fit its dense matrices (moderate scales: the
    // chain's usual x0.4 .. x1.5 steps with bilinear / bicubic / area / nearest windows); the two tiled kernels below stay
    // for long windows (lanczos prefilters, extreme down-scales).
    if (forced == 0 || forced == 6) {
        const RbPlan pl = rb_plan(ay, ax);
        if (pl.cfg >= 0) {
            const int vec = (W % 4 == 0) && (((uintptr_t)img & 15) == 0);
            const int vec_out = (OW % pl.GH == 0) && (((uintptr_t)out & (4 * pl.GH - 1)) == 0);
            const int* dense = ty_lo + sparse_ints(ay, ax);
            const dim3 grid(pl.tiles_x, pl.tiles_y, planes);
#define OTF_RESIZE_RB3(GV_, RV_, GH_, SH_, V_, N_)                                                                        \
    do {                                                                                                                  \
        auto kfn = resize_rb_kernel<GV_, RV_, GH_, SH_, V_, N_>;                                                          \
        if (pl.smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem);         \
        launch_chain(kfn, dim3(grid), dim3(128), pl.smem, st, img, out, ay, ax, dense, pl, clamp_out, vec_out, ne);                             \
    } while (0)
#define OTF_RESIZE_RB(GV_, RV_, GH_, SH_)                                                                                 \
    do {                                                                                                                  \
        if (vec && noise) OTF_RESIZE_RB3(GV_, RV_, GH_, SH_, true, true);                                                 \
        else if (vec) OTF_RESIZE_RB3(GV_, RV_, GH_, SH_, true, false);                                                    \
        else if (noise) OTF_RESIZE_RB3(GV_, RV_, GH_, SH_, false, true);                                                  \
        else OTF_RESIZE_RB3(GV_, RV_, GH_, SH_, false, false);                                                            \
    } while (0)
            if (pl.cfg == 0) OTF_RESIZE_RB(4, 10, 4, 3);
            else if (pl.cfg == 1) OTF_RESIZE_RB(4, 13, 4, 4);
            else OTF_RESIZE_RB(2, 12, 2, 4);
            OTF_LAUNCH_CHECK("resize_rb_kernel");
            return OTF_OK;
        }
    }
    if (impl == 4) {
        const int vec = (W % 4 == 0) && (((uintptr_t)img & 15) == 0);
        const int NTv = mt <= 1 ? 1 : mt <= 2 ? 2 : mt <= 3 ? 3 : mt <= 4 ? 4 : mt <= 6 ? 6 : mt <= 8 ? 8 : mt <= 12 ? 12 : mt <= 16 ? 16 : 0;
        const int ntx = NTv ? NTv : ax.max_taps, nty = NTv ? NTv : ay.max_taps;
        int TW = OW >= 96 ? 128 : OW >= 48 ? 64 : 32, TH = 32, pitch = 0;
        size_t smem = 0;
        for (;;) {
            pitch = ((span(ax, TW) + (vec ? 3 : 0) + 3) & ~3) + ((ntx + 3) & ~3);
            smem = ((size_t)TH * pitch + (size_t)TH * nty + TH + TW + (size_t)TW * (ntx | 1)) * 4;
            if (smem <= 40 * 1024 || (TH == 4 && TW == 32)) break;
            if (TH > 8 || TW == 32) TH /= 2; else TW /= 2;
        }
        if (smem <= 200 * 1024) {
            const dim3 grid(ceil_div(OW, TW), ceil_div(OH, TH), planes);
#define OTF_RESIZE_VH2(NT_, V_, N_)                                                                                       \
    do {                                                                                                                  \
        auto kfn = resize_vh_kernel<NT_, V_, N_>;                                                                         \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
        launch_chain(kfn, dim3(grid), dim3(256), smem, st, img, out, ay, ax, ty_lo, tx_lo, TW, TH, pitch, clamp_out, ne);                       \
    } while (0)
#define OTF_RESIZE_VH(NT_)                                                                                                \
    do {                                                                                                                  \
        if (vec && noise) OTF_RESIZE_VH2(NT_, true, true);                                                                \
        else if (vec) OTF_RESIZE_VH2(NT_, true, false);                                                                   \
        else if (noise) OTF_RESIZE_VH2(NT_, false, true);                                                                 \
        else OTF_RESIZE_VH2(NT_, false, false);                                                                           \
    } while (0)
            switch (NTv) {
                case 1: OTF_RESIZE_VH(1); break;
                case 2: OTF_RESIZE_VH(2); break;
                case 3: OTF_RESIZE_VH(3); break;
                case 4: OTF_RESIZE_VH(4); break;
                case 6: OTF_RESIZE_VH(6); break;
                case 8: OTF_RESIZE_VH(8); break;
                case 12: OTF_RESIZE_VH(12); break;
                case 16: OTF_RESIZE_VH(16); break;
                default: OTF_RESIZE_VH(0); break;
            }
            OTF_LAUNCH_CHECK("resize_vh_kernel");
            return OTF_OK;
        }
    }
    if (mt > 64) {
        if (noise) launch_chain(resize_generic_kernel<true>, dim3(dim3(ceil_div(OW, 32), ceil_div(OH, 8), planes)), dim3(256), 0, st, img, out, ay, ax, ty_lo, tx_lo, clamp_out, ne);
        else launch_chain(resize_generic_kernel<false>, dim3(dim3(ceil_div(OW, 32), ceil_div(OH, 8), planes)), dim3(256), 0, st, img, out, ay, ax, ty_lo, tx_lo, clamp_out, ne);
        OTF_LAUNCH_CHECK("resize_generic_kernel");
        return OTF_OK;
    }
    const int NT = mt <= 1 ? 1 : mt <= 2 ? 2 : mt <= 4 ? 4 : mt <= 6 ? 6 : mt <= 8 ? 8 : mt <= 12 ? 12 : mt <= 16 ? 16
                   : mt <= 24 ? 24 : mt <= 32 ? 32 : 64;
    // tile width 32*CW: as wide as the output allows — per-CTA prologue latency (tables, fill, two barriers)
    // dominates small tiles, so a few padded columns cost less than more, smaller CTAs (measured: 2x at 256->320)
    int CW = OW >= 96 ? 4 : OW >= 48 ? 2 : 1;
    const int vec_ok = (W % 4 == 0) && (((uintptr_t)img & 15) == 0);
    // pick (CW, tile_h): prefer wide tiles, keep shared memory modest so several CTAs share an SM
    const size_t cap = 200 * 1024, want = 56 * 1024;
    int tile_h = 0, rows_cap = 0, cols_cap = 0;
    size_t smem = 0;
    for (;; CW /= 2) {
        const int TW = 32 * CW;
        cols_cap = span(ax, TW) + (vec_ok ? 6 : 0);  // slack for the 16-byte aligned rectangle start/end
        bool ok = false;
        for (tile_h = 32; tile_h >= 4; tile_h /= 2) {
            rows_cap = span(ay, tile_h);
            smem = ((size_t)TW * NT + (size_t)tile_h * NT + TW + tile_h + (size_t)rows_cap * ((cols_cap + NT + 3) & ~3) +
                    (size_t)(rows_cap + NT) * TW) * 4;
            if (smem <= want || (tile_h == 4 && CW == 1 && smem <= cap)) { ok = true; break; }
        }
        if (ok || CW == 1) break;
    }
    OTF_REQUIRE(smem <= cap, OTF_ERR_UNSUPPORTED, "resize: scale %f x %f too extreme for shared memory", (double)ay.scale,
                (double)ax.scale);
    if (tile_h < 4) tile_h = 4;
    const dim3 grid(ceil_div(OW, 32 * CW), ceil_div(OH, tile_h), planes);
#define OTF_RESIZE_LAUNCH(NT_, CW_)                                                                                       \
    do {                                                                                                                  \
        if (noise) OTF_RESIZE_LAUNCH2(NT_, CW_, true);                                                                    \
        else OTF_RESIZE_LAUNCH2(NT_, CW_, false);                                                                         \
    } while (0)
#define OTF_RESIZE_LAUNCH2(NT_, CW_, N_)                                                                                  \
    do {                                                                                                                  \
        auto kfn = resize_kernel<NT_, CW_, N_>;                                                                           \
        if (smem > 48 * 1024) {                                                                                           \
            cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
            if (e != cudaSuccess) return cuda_fail(e, "resize smem attribute");                                           \
        }                                                                                                                 \
        launch_chain(kfn, dim3(grid), dim3(256), smem, st, img, out, ay, ax, ty_lo, tx_lo, tile_h, rows_cap, cols_cap, clamp_out, vec_ok, ne);        \
    } while (0)
#define OTF_RESIZE_CW(NT_)                                  \
    do {                                                    \
        if (CW == 4) OTF_RESIZE_LAUNCH(NT_, 4);             \
        else if (CW == 2) OTF_RESIZE_LAUNCH(NT_, 2);        \
        else OTF_RESIZE_LAUNCH(NT_, 1);                     \
    } while (0)
    switch (NT) {
        case 1: OTF_RESIZE_CW(1); break;
        case 2: OTF_RESIZE_CW(2); break;
        case 4: OTF_RESIZE_CW(4); break;
        case 6: OTF_RESIZE_CW(6); break;
        case 8: OTF_RESIZE_CW(8); break;
        case 12: OTF_RESIZE_CW(12); break;
        case 16: OTF_RESIZE_CW(16); break;
        case 24: OTF_RESIZE_CW(24); break;
        case 32: OTF_RESIZE_CW(32); break;
        default: OTF_RESIZE_CW(64); break;
    }
    OTF_LAUNCH_CHECK("resize_kernel");
    return OTF_OK;
}
}  // namespace otf

extern "C" int otf_resize_f32(const float* img, int planes, int H, int W, float* out, int OH, int OW, int mode,
                              int clamp_out, void* workspace_dev, int64_t workspace_bytes, int tables_ready,
                              void* stream) {
    return otf::resize_impl(img, planes, H, W, out, OH, OW, mode, clamp_out, workspace_dev, workspace_bytes, tables_ready, stream,
                            nullptr);
}

extern "C" int otf_resize_gauss_f32(const float* img, int B, int C, int H, int W, float* out, int OH, int OW, int mode,
                                    int clamp_out, void* workspace_dev, int64_t workspace_bytes, int tables_ready,
                                    const float* sigma_dev, const float* gray_dev, uint64_t seed, uint64_t offset,
                                    const uint64_t* offset_dev, int noise_flags, void* stream) {
    using namespace otf;
    OTF_REQUIRE(sigma_dev, OTF_ERR_BAD_ARG, "resize_gauss: null sigma");
    OTF_REQUIRE(B > 0 && C > 0, OTF_ERR_BAD_ARG, "resize_gauss: bad extents");
    OTF_REQUIRE(!(noise_flags & (OTF_NOISE_ROUNDS | OTF_NOISE_FIELD_ONLY | OTF_NOISE_RAW_FIELD)), OTF_ERR_UNSUPPORTED,
                "resize_gauss: only the plain / clip tails are fused (rounds, field-only and injected fields take two launches)");
    NoiseEpi ne;
    memset(&ne, 0, sizeof(ne));
    ne.sigma = sigma_dev; ne.gray = gray_dev; ne.offset_dev = offset_dev; ne.seed = seed; ne.offset = offset; ne.C = C;
    ne.flags = noise_flags;
    return resize_impl(img, B * C, H, W, out, OH, OW, mode, clamp_out, workspace_dev, workspace_bytes, tables_ready, stream, &ne);
}

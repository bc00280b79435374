// resize: separable resampling with ATen's index/weight rules, + fused clamp(0,1).
// Replaces traiNNer/data/degradations.py:1004-1021 (`resize_pt`): F.interpolate with
// bilinear/bicubic antialias=True, area (adaptive avg pool), nearest-exact, and the plain
// (non-aa, a=-0.75) bicubic that ends the "lanczos" mode.
//
// Every mode is expressed as, per output index o on an axis, a contiguous source window
// [lo, lo+n) with n normalised fp32 weights (ATen: aten/src/ATen/native/cpu/UpSampleKernel.cpp
// `_compute_indices_min_size_weights_aa`; restated in SURVEY.md §8a row a3 "R").  A CTA owns an
// output tile; it builds the tile's column/row weight tables in shared memory, runs the
// horizontal pass over just the source rows the tile needs (result kept in smem), then the
// vertical pass — the same order ATen uses (W first, then H).  HBM-bound: source rows are read
// once per tile (+ vertical halo), output written once.
#include "otf_common.cuh"

namespace otf {

constexpr int RT_W = 32;  // output tile width  (one lane per column)

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

// smem layout: wx[RT_W][tx] | wy[tile_h][ty] | xlo[RT_W] xn[RT_W] ylo[tile_h] yn[tile_h] | tmp[rows_cap][RT_W]
__global__ void __launch_bounds__(256) resize_kernel(const float* __restrict__ img, float* __restrict__ out, int mode,
                                                     AxisSpec ay, AxisSpec ax, int tile_h, int rows_cap, int clamp_out) {
    extern __shared__ __align__(16) float sm[];
    float* wx = sm;
    float* wy = wx + RT_W * ax.max_taps;
    int* xlo = reinterpret_cast<int*>(wy + tile_h * ay.max_taps);
    int* xn = xlo + RT_W;
    int* ylo = xn + RT_W;
    int* yn = ylo + tile_h;
    float* tmp = reinterpret_cast<float*>(yn + tile_h);

    const int plane = blockIdx.z;
    const int ox0 = blockIdx.x * RT_W, oy0 = blockIdx.y * tile_h;
    const int tid = threadIdx.x;
    if (tid < RT_W) {
        const int o = min(ox0 + tid, ax.out_n - 1);
        axis_weights(mode, ax, o, &xlo[tid], &xn[tid], wx + tid * ax.max_taps);
    } else if (tid - RT_W < tile_h) {
        const int t = tid - RT_W;
        const int o = min(oy0 + t, ay.out_n - 1);
        axis_weights(mode, ay, o, &ylo[t], &yn[t], wy + t * ay.max_taps);
    }
    __syncthreads();
    // source rows this tile needs
    const int th = min(tile_h, ay.out_n - oy0);
    int row_lo = ylo[0], row_hi = ylo[0] + yn[0];
    for (int t = 1; t < th; ++t) {
        row_lo = min(row_lo, ylo[t]);
        row_hi = max(row_hi, ylo[t] + yn[t]);
    }
    const int nrows = min(row_hi - row_lo, rows_cap);
    const float* ip = img + (size_t)plane * ay.in_n * ax.in_n;
    const int lane = tid & 31, wid = tid >> 5;
    // horizontal pass: tmp[r][lane] = sum_j wx[lane][j] * in[row_lo + r][xlo[lane] + j]
    {
        const int lo = xlo[lane], n = xn[lane];
        const float* wl = wx + lane * ax.max_taps;
        for (int r = wid; r < nrows; r += 8) {
            const float* rp = ip + (size_t)(row_lo + r) * ax.in_n + lo;
            float acc = 0.0f;
            for (int j = 0; j < n; ++j) acc = fmaf(wl[j], __ldg(rp + j), acc);
            tmp[r * RT_W + lane] = acc;
        }
    }
    __syncthreads();
    // vertical pass
    const int x = ox0 + lane;
    if (x >= ax.out_n) return;
    for (int t = wid; t < th; t += 8) {
        const float* wl = wy + t * ay.max_taps;
        const float* tp = tmp + (ylo[t] - row_lo) * RT_W + lane;
        const int n = yn[t];
        float acc = 0.0f;
        for (int i = 0; i < n; ++i) acc = fmaf(wl[i], tp[i * RT_W], acc);
        if (clamp_out) acc = clamp01(acc);
        out[(size_t)plane * ay.out_n * ax.out_n + (size_t)(oy0 + t) * ax.out_n + x] = acc;
    }
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
    } else if (mode == OTF_RESIZE_NEAREST_EXACT) {
        a.max_taps = 1;
    } else {
        a.max_taps = 4;
    }
    return a;
}

}  // namespace otf

extern "C" int otf_resize_f32(const float* img, int planes, int H, int W, float* out, int OH, int OW, int mode,
                              int clamp_out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && img != out, OTF_ERR_BAD_ARG, "resize: bad pointers");
    OTF_REQUIRE(planes > 0 && planes <= 65535 && H > 0 && W > 0 && OH > 0 && OW > 0, OTF_ERR_BAD_ARG, "resize: bad extents");
    OTF_REQUIRE(mode >= OTF_RESIZE_BILINEAR_AA && mode <= OTF_RESIZE_BICUBIC, OTF_ERR_BAD_ARG, "resize: unknown mode %d", mode);
    const AxisSpec ay = make_axis(mode, H, OH), ax = make_axis(mode, W, OW);
    // pick the tallest tile (<= 32 rows) whose tables + row buffer fit in shared memory
    const size_t cap = 200 * 1024;
    int tile_h = 32, rows_cap = 0;
    size_t smem = 0;
    for (; tile_h >= 1; tile_h /= 2) {
        // rows spanned by tile_h consecutive outputs: windows advance by `scale` per output
        rows_cap = (int)ceilf(ay.scale * (float)(tile_h - 1)) + ay.max_taps + 2;
        if (rows_cap > H) rows_cap = H;
        smem = ((size_t)RT_W * ax.max_taps + (size_t)tile_h * ay.max_taps + 2 * RT_W + 2 * tile_h + (size_t)rows_cap * RT_W) * 4;
        if (smem <= cap) break;
    }
    OTF_REQUIRE(tile_h >= 1, OTF_ERR_UNSUPPORTED, "resize: scale %f too extreme for shared memory", (double)ay.scale);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(resize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "resize smem attribute");
    }
    const dim3 grid(ceil_div(OW, RT_W), ceil_div(OH, tile_h), planes);
    resize_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(img, out, mode, ay, ax, tile_h, rows_cap, clamp_out);
    OTF_LAUNCH_CHECK("resize_kernel");
    return OTF_OK;
}

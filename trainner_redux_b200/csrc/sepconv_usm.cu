// 1-D reflect-padded correlation (horizontal / vertical) and the fused USM sharpen.
// Replaces traiNNer/utils/img_process_util.py:35-55 (USMSharp: two 51x51 filter2d calls +
// ~8 elementwise launches) and the Lanczos prefilter of traiNNer/data/degradations.py:982-998.
//
// The USM Gaussian is an exact outer product (cv2.getGaussianKernel(r) x itself), so the
// 2601-tap 2-D correlation is evaluated as 51 + 51 taps.  HBM-bound: every pass streams the
// plane once; taps ride in the kernel parameter bank (uniform loads, no LSU traffic).
#include <stdlib.h>

#include "otf_common.cuh"

namespace otf {

constexpr int kMaxTaps = 1023;  // 4 KB of the kernel parameter bank
struct Taps {
    float w[kMaxTaps + 1];
    int n;
};

enum { EPI_NONE = 0, EPI_USM_MASK = 1, EPI_USM_BLEND = 2 };

// Horizontal pass: a CTA owns ROWS rows x TW columns; the row segment (+halo) sits in smem.
constexpr int H_TW = 256, H_ROWS = 4;
__global__ void __launch_bounds__(256) sepconv_h_generic_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W,
                                                        const __grid_constant__ Taps taps) {
    extern __shared__ float sm[];
    const int r = taps.n / 2, span = H_TW + 2 * r;
    const int plane = blockIdx.z, x0 = blockIdx.x * H_TW, y0 = blockIdx.y * H_ROWS;
    const float* ip = img + (size_t)plane * H * W;
    for (int row = 0; row < H_ROWS; ++row) {
        const int y = y0 + row;
        if (y >= H) break;
        for (int xx = threadIdx.x; xx < span; xx += blockDim.x) {
            const int gx = clampi(reflect_idx(x0 - r + xx, W), 0, W - 1);
            sm[row * span + xx] = __ldg(ip + (size_t)y * W + gx);
        }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= W) return;
    for (int row = 0; row < H_ROWS; ++row) {
        const int y = y0 + row;
        if (y >= H) break;
        const float* sp = sm + row * span + threadIdx.x;
        float acc = 0.0f;
        for (int j = 0; j < taps.n; ++j) acc = fmaf(taps.w[j], sp[j], acc);
        out[(size_t)plane * H * W + (size_t)y * W + x] = acc;
    }
}

// Vertical pass: a CTA owns a 32-wide x 64-tall tile; (64 + 2r) rows of it sit in smem.
// Epilogues fuse the USM elementwise work into the pass that produces the blur.
constexpr int V_TW = 32, V_TH = 64;
template <int EPI>
__global__ void __launch_bounds__(256) sepconv_v_generic_kernel(const float* __restrict__ tmp, float* __restrict__ out, int H, int W,
                                                        const __grid_constant__ Taps taps,
                                                        const float* __restrict__ img, float* __restrict__ aux,
                                                        float weight, float threshold) {
    extern __shared__ float sm[];
    const int r = taps.n / 2, rows = V_TH + 2 * r;
    const int plane = blockIdx.z, x0 = blockIdx.x * V_TW, y0 = blockIdx.y * V_TH;
    const float* tp = tmp + (size_t)plane * H * W;
    const int lx = threadIdx.x & 31, wy = threadIdx.x >> 5;  // 8 warps
    const int x = x0 + lx;
    const int gx = min(x, W - 1);
    for (int yy = wy; yy < rows; yy += 8) {
        const int gy = clampi(reflect_idx(y0 - r + yy, H), 0, H - 1);
        sm[yy * V_TW + lx] = __ldg(tp + (size_t)gy * W + gx);
    }
    __syncthreads();
    if (x >= W) return;
    for (int oy = wy; oy < V_TH; oy += 8) {
        const int y = y0 + oy;
        if (y >= H) break;
        const float* sp = sm + oy * V_TW + lx;
        float acc = 0.0f;
        for (int i = 0; i < taps.n; ++i) acc = fmaf(taps.w[i], sp[i * V_TW], acc);
        const size_t o = (size_t)plane * H * W + (size_t)y * W + x;
        if (EPI == EPI_NONE) {
            out[o] = acc;
        } else if (EPI == EPI_USM_MASK) {
            // img_process_util.py:47-53: residual, hard mask, clipped sharpen
            const float im = img[o];
            const float res = __fsub_rn(im, acc);
            out[o] = (__fmul_rn(fabsf(res), 255.0f) > threshold) ? 1.0f : 0.0f;           // mask
            aux[o] = clamp01(__fadd_rn(im, __fmul_rn(weight, res)));                        // sharp
        } else {
            // img_process_util.py:55: soft*sharp + (1-soft)*img   (acc = soft mask, aux = sharp)
            const float im = img[o];
            out[o] = __fadd_rn(__fmul_rn(acc, aux[o]), __fmul_rn(__fsub_rn(1.0f, acc), im));
        }
    }
}

static int fill_taps(Taps& t, const float* taps_host, int ntaps) {
    OTF_REQUIRE(taps_host, OTF_ERR_BAD_ARG, "sepconv: null taps");
    OTF_REQUIRE(ntaps > 0 && (ntaps % 2) == 1 && ntaps <= kMaxTaps, OTF_ERR_BAD_ARG, "sepconv: ntaps %d must be odd and <= %d", ntaps, kMaxTaps);
    for (int i = 0; i < ntaps; ++i) t.w[i] = taps_host[i];
    for (int i = ntaps; i <= kMaxTaps; ++i) t.w[i] = 0.0f;
    t.n = ntaps;
    return OTF_OK;
}

static int launch_h_generic(const float* img, int planes, int H, int W, const Taps& t, float* out, cudaStream_t st) {
    const dim3 grid(ceil_div(W, H_TW), ceil_div(H, H_ROWS), planes);
    const size_t smem = (size_t)H_ROWS * (H_TW + 2 * (t.n / 2)) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sepconv_h_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "sepconv_h smem attribute");
    }
    sepconv_h_generic_kernel<<<grid, 256, smem, st>>>(img, out, H, W, t);
    OTF_LAUNCH_CHECK("sepconv_h_generic_kernel");
    return OTF_OK;
}

template <int EPI>
static int launch_v_generic(const float* tmp, int planes, int H, int W, const Taps& t, float* out, const float* img, float* aux,
                    float weight, float threshold, cudaStream_t st) {
    const dim3 grid(ceil_div(W, V_TW), ceil_div(H, V_TH), planes);
    const size_t smem = (size_t)(V_TH + 2 * (t.n / 2)) * V_TW * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sepconv_v_generic_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "sepconv_v smem attribute");
    }
    sepconv_v_generic_kernel<EPI><<<grid, 256, smem, st>>>(tmp, out, H, W, t, img, aux, weight, threshold);
    OTF_LAUNCH_CHECK("sepconv_v_generic_kernel");
    return OTF_OK;
}

// ---- register-blocked passes for <= 63 taps (USM radius <= 31, every Lanczos prefilter in range) ----
// The taps sit centred in an NTP-slot array (NTP = 8/16/32/64, centre at NTP/2, zeros around), passed in the
// kernel parameter bank: with the tap loop fully unrolled every FFMA takes its weight as a constant-bank
// operand — no load instruction at all.  A thread produces 8 outputs from a (8 + NTP)-pixel window held in
// registers, so shared memory is read once per 8 x NTP FMAs (LDS.128 along rows, conflict-free LDS.32 down
// columns).  FMA-bound: 2*NTP flop per output per pass.
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// Packed FMA form (sm_100a FFMA2, fma.rn.f32x2): two adjacent outputs share the pixel, (out[2k+1], out[2k]) += p * (w[m-1], w[m])
// with p = win[2k + m], m = 0..NTP and w[-1] = w[NTP] = 0 — NTP + 1 packed instructions per output pair instead of 2 NTP
// scalar ones, every lane still the ascending fmaf chain of the scalar loop (bit-identical results).  The tap PAIRS ride
// in the kernel parameter bank as 64-bit constant operands.  ncu on the scalar version: issue slots 58-73 % busy with the
// FMA pipe at 34-37 % — issue-bound, which is what the packed form halves.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
    const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

template <int NTP>
struct TapsC {
    float2 p[NTP + 1];  // p[m] = (w[m-1], w[m]) over the NTP centred slots
};

template <int NTP>
__global__ void __launch_bounds__(256) sepconv_h_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W,
                                                        const __grid_constant__ TapsC<NTP> taps) {
    constexpr int CEN = NTP / 2, TW = 256, ROWS = 8, SWR = TW + NTP;  // smem col s <-> x = x0 - CEN + s
    constexpr int SW = ((SWR / 4) & 1) ? SWR : SWR + 4;                // odd number of 16-byte chunks per row
    static_assert(NTP % 4 == 0, "window in float4s");
    extern __shared__ __align__(16) float sm[];
    const int plane = blockIdx.z, x0 = blockIdx.x * TW, y0 = blockIdx.y * ROWS;
    const float* ip = img + (size_t)plane * H * W;
    for (int row = threadIdx.x >> 5; row < ROWS; row += 8) {
        const int y = min(y0 + row, H - 1);
        for (int s = threadIdx.x & 31; s < SWR; s += 32) {
            const int gx = clampi(reflect_idx(x0 - CEN + s, W), 0, W - 1);
            cp_async_f32(&sm[row * SW + s], ip + (size_t)y * W + gx);  // all copies in flight at once
        }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // a warp = 2 rows x 16 segments of 8 outputs; the 8 lanes of an LDS.128 phase are 2 rows x 4 segments, which with the
    // odd chunk pitch touch 8 distinct 16-byte chunks mod 8 (8 lanes of ONE row, 32 bytes apart, would conflict 2-way)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = 2 * (warp >> 1) + ((lane >> 2) & 1);
    const int seg = (16 * (warp & 1) + 4 * (lane >> 3) + (lane & 3)) * 8;
    const int y = y0 + row;
    if (y >= H || x0 + seg >= W) return;
    float win[8 + NTP];
    const float4* wp = reinterpret_cast<const float4*>(sm + row * SW + seg);
#pragma unroll
    for (int q = 0; q < (8 + NTP) / 4; ++q) {
        const float4 v = wp[q];
        win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
    }
    float2 a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = make_float2(0.0f, 0.0f);  // (out[2k+1], out[2k])
#pragma unroll
    for (int m = 0; m <= NTP; ++m)
#pragma unroll
        for (int k = 0; k < 4; ++k) ffma2(a[k], make_float2(win[2 * k + m], win[2 * k + m]), taps.p[m]);
    float* op = out + (size_t)plane * H * W + (size_t)y * W + x0 + seg;
    const float acc[8] = {a[0].y, a[0].x, a[1].y, a[1].x, a[2].y, a[2].x, a[3].y, a[3].x};
    if (x0 + seg + 8 <= W && (W & 3) == 0 && ((uintptr_t)out & 15) == 0) {
        reinterpret_cast<float4*>(op)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        reinterpret_cast<float4*>(op)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
        for (int ox = 0; ox < 8; ++ox)
            if (x0 + seg + ox < W) op[ox] = acc[ox];
    }
}

template <int NTP, int EPI>
__global__ void __launch_bounds__(256) sepconv_v_kernel(const float* __restrict__ tmp, float* __restrict__ out, int H, int W,
                                                        const __grid_constant__ TapsC<NTP> taps,
                                                        const float* __restrict__ img, float* __restrict__ aux,
                                                        float weight, float threshold) {
    constexpr int CEN = NTP / 2, TWV = 32, TH = 64, SH = TH + NTP;  // smem row s <-> y = y0 - CEN + s
    extern __shared__ __align__(16) float sm[];
    const int plane = blockIdx.z, x0 = blockIdx.x * TWV, y0 = blockIdx.y * TH;
    const float* tp = tmp + (size_t)plane * H * W;
    const int lx = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int x = x0 + lx, gx = min(x, W - 1);
    for (int s = wy; s < SH; s += 8) {
        const int gy = clampi(reflect_idx(y0 - CEN + s, H), 0, H - 1);
        cp_async_f32(&sm[s * TWV + lx], tp + (size_t)gy * W + gx);
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int ys = wy * 8;  // this warp's 8 output rows
    if (x >= W || y0 + ys >= H) return;
    // (requesting the epilogue's global operands before the tap loop was tried: 60 registers instead of 32 halve the
    // resident CTAs and the four passes take 0.231 ms instead of 0.185)
    float win[8 + NTP];
#pragma unroll
    for (int k = 0; k < 8 + NTP; ++k) win[k] = sm[(ys + k) * TWV + lx];
    float2 a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = make_float2(0.0f, 0.0f);  // (out[2k+1], out[2k]) down the column
#pragma unroll
    for (int m = 0; m <= NTP; ++m)
#pragma unroll
        for (int k = 0; k < 4; ++k) ffma2(a[k], make_float2(win[2 * k + m], win[2 * k + m]), taps.p[m]);
    const float acc[8] = {a[0].y, a[0].x, a[1].y, a[1].x, a[2].y, a[2].x, a[3].y, a[3].x};
#pragma unroll
    for (int oy = 0; oy < 8; ++oy) {
        const int y = y0 + ys + oy;
        if (y >= H) break;
        const size_t o = (size_t)plane * H * W + (size_t)y * W + x;
        if (EPI == EPI_NONE) {
            out[o] = acc[oy];
        } else if (EPI == EPI_USM_MASK) {
            // img_process_util.py:47-53: residual, hard mask, clipped sharpen
            const float im = img[o];
            const float res = __fsub_rn(im, acc[oy]);
            out[o] = (__fmul_rn(fabsf(res), 255.0f) > threshold) ? 1.0f : 0.0f;  // mask
            aux[o] = clamp01(__fadd_rn(im, __fmul_rn(weight, res)));               // sharp
        } else {
            // img_process_util.py:55: soft*sharp + (1-soft)*img   (acc = soft mask, aux = sharp)
            const float im = img[o];
            out[o] = __fadd_rn(__fmul_rn(acc[oy], aux[o]), __fmul_rn(__fsub_rn(1.0f, acc[oy]), im));
        }
    }
}

template <int NTP>
static TapsC<NTP> centre_taps(const Taps& t) {
    float w[NTP + 2];  // w[1 + slot]; w[0] = w[NTP + 1] = 0
    for (int i = 0; i < NTP + 2; ++i) w[i] = 0.0f;
    const int r = t.n / 2;
    for (int i = 0; i < t.n; ++i) w[1 + NTP / 2 - r + i] = t.w[i];
    TapsC<NTP> c;
    for (int m = 0; m <= NTP; ++m) c.p[m] = make_float2(w[m], w[m + 1]);  // (w[m-1], w[m])
    return c;
}

static int launch_h_generic(const float* img, int planes, int H, int W, const Taps& t, float* out, cudaStream_t st);
template <int EPI>
static int launch_v_generic(const float* tmp, int planes, int H, int W, const Taps& t, float* out, const float* img, float* aux,
                            float weight, float threshold, cudaStream_t st);

template <int NTP>
static int launch_h_fast(const float* img, int planes, int H, int W, const Taps& t, float* out, cudaStream_t st) {
    const dim3 grid(ceil_div(W, 256), ceil_div(H, 8), planes);
    sepconv_h_kernel<NTP><<<grid, 256, 8 * (256 + NTP + 4) * sizeof(float), st>>>(img, out, H, W, centre_taps<NTP>(t));
    OTF_LAUNCH_CHECK("sepconv_h_kernel");
    return OTF_OK;
}
static int launch_h(const float* img, int planes, int H, int W, const Taps& t, float* out, cudaStream_t st) {
    const int need = 2 * (t.n / 2) + 2;  // NTP must satisfy r <= NTP/2 - 1
    if (need <= 8) return launch_h_fast<8>(img, planes, H, W, t, out, st);
    if (need <= 16) return launch_h_fast<16>(img, planes, H, W, t, out, st);
    if (need <= 24) return launch_h_fast<24>(img, planes, H, W, t, out, st);
    if (need <= 32) return launch_h_fast<32>(img, planes, H, W, t, out, st);
    if (need <= 40) return launch_h_fast<40>(img, planes, H, W, t, out, st);
    if (need <= 52) return launch_h_fast<52>(img, planes, H, W, t, out, st);  // USMSharp's default radius 50 -> 51 taps
    if (need <= 64) return launch_h_fast<64>(img, planes, H, W, t, out, st);
    return launch_h_generic(img, planes, H, W, t, out, st);
}

template <int NTP, int EPI>
static int launch_v_fast(const float* tmp, int planes, int H, int W, const Taps& t, float* out, const float* img, float* aux,
                         float weight, float threshold, cudaStream_t st) {
    const dim3 grid(ceil_div(W, 32), ceil_div(H, 64), planes);
    sepconv_v_kernel<NTP, EPI><<<grid, 256, (64 + NTP) * 32 * sizeof(float), st>>>(tmp, out, H, W, centre_taps<NTP>(t), img, aux,
                                                                                   weight, threshold);
    OTF_LAUNCH_CHECK("sepconv_v_kernel");
    return OTF_OK;
}
template <int EPI>
static int launch_v(const float* tmp, int planes, int H, int W, const Taps& t, float* out, const float* img, float* aux,
                    float weight, float threshold, cudaStream_t st) {
    const int need = 2 * (t.n / 2) + 2;
    if (need <= 8) return launch_v_fast<8, EPI>(tmp, planes, H, W, t, out, img, aux, weight, threshold, st);
    if (need <= 16) return launch_v_fast<16, EPI>(tmp, planes, H, W, t, out, img, aux, weight, threshold, st);
    if (need <= 24) return launch_v_fast<24, EPI>(tmp, planes, H, W, t, out, img, aux, weight, threshold, st);
    if (need <= 32) return launch_v_fast<32, EPI>(tmp, planes, H, W, t, out, img, aux, weight, threshold, st);
    if (need <= 40) return launch_v_fast<40, EPI>(tmp, planes, H, W, t, out, img, aux, weight, threshold, st);
    if (need <= 52) return launch_v_fast<52, EPI>(tmp, planes, H, W, t, out, img, aux, weight, threshold, st);
    if (need <= 64) return launch_v_fast<64, EPI>(tmp, planes, H, W, t, out, img, aux, weight, threshold, st);
    return launch_v_generic<EPI>(tmp, planes, H, W, t, out, img, aux, weight, threshold, st);
}

}  // namespace otf

extern "C" int otf_sepconv_reflect_f32(const float* img, int planes, int H, int W, const float* taps_host, int ntaps,
                                       int axis, float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && img != out, OTF_ERR_BAD_ARG, "sepconv: bad pointers");
    OTF_REQUIRE(planes > 0 && planes <= 65535 && H > 0 && W > 0, OTF_ERR_BAD_ARG, "sepconv: bad extents");
    Taps t;
    if (int rc = fill_taps(t, taps_host, ntaps)) return rc;
    OTF_REQUIRE(ntaps / 2 < (axis == 0 ? H : W), OTF_ERR_BAD_ARG, "sepconv: reflect pad %d needs extent > pad", ntaps / 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (axis == 1) return launch_h(img, planes, H, W, t, out, st);
    return launch_v<EPI_NONE>(img, planes, H, W, t, out, nullptr, nullptr, 0.f, 0.f, st);
}

extern "C" int64_t otf_usm_workspace_bytes(int planes, int H, int W) {
    return (int64_t)planes * H * W * sizeof(float) * 3;  // tmp, mask, sharp
}

static int usm_chunk_planes(int planes, int H, int W) {
    // OTF_USM_CHUNK_PLANES=n runs the four passes n planes at a time (A/B switch).  Measured at 64 x 3 x 256^2
    // (profiles/r02_microbench_usm_chunks.txt): 37 planes 0.224 ms, 48: 0.219, 64: 0.206, 96: 0.200, whole batch 0.185 —
    // the L2 residency it buys is worth less than the partial waves and launch gaps of 4x more, smaller launches.
    static const int chunk_planes_env = getenv("OTF_USM_CHUNK_PLANES") ? atoi(getenv("OTF_USM_CHUNK_PLANES")) : 0;
    (void)H; (void)W;
    if (chunk_planes_env > 0 && chunk_planes_env < planes) return chunk_planes_env;
    return planes;
}

extern "C" int otf_usm_launch_count(int planes, int H, int W) {
    if (planes <= 0 || H <= 0 || W <= 0) return 0;
    const int chunk = usm_chunk_planes(planes, H, W);
    return 4 * ((planes + chunk - 1) / chunk);
}

extern "C" int otf_usm_sharp_f32(const float* img, int planes, int H, int W, const float* taps_host, int ntaps,
                                 float weight, float threshold, void* workspace_dev, int64_t workspace_bytes,
                                 float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && workspace_dev, OTF_ERR_BAD_ARG, "usm: null pointer");
    OTF_REQUIRE(planes > 0 && planes <= 65535 && H > 0 && W > 0, OTF_ERR_BAD_ARG, "usm: bad extents");
    OTF_REQUIRE(workspace_bytes >= otf_usm_workspace_bytes(planes, H, W), OTF_ERR_WORKSPACE, "usm: workspace too small");
    Taps t;
    if (int rc = fill_taps(t, taps_host, ntaps)) return rc;
    OTF_REQUIRE(ntaps / 2 < H && ntaps / 2 < W, OTF_ERR_BAD_ARG, "usm: reflect pad %d needs H,W > pad (got %dx%d)", ntaps / 2, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    // (optionally chunk by chunk, see usm_chunk_planes: the intermediates of every chunk sit at the same workspace addresses)
    const int chunk = usm_chunk_planes(planes, H, W);
    const size_t n = (size_t)chunk * H * W;
    float* tmp = (float*)workspace_dev;
    float* mask = tmp + n;
    float* sharp = mask + n;
    for (int p0 = 0; p0 < planes; p0 += chunk) {
        const int pc = planes - p0 < chunk ? planes - p0 : chunk;
        const float* im = img + (size_t)p0 * H * W;
        float* o = out + (size_t)p0 * H * W;
        int rc;
        if ((rc = launch_h(im, pc, H, W, t, tmp, st))) return rc;                                          // blur, rows
        if ((rc = launch_v<EPI_USM_MASK>(tmp, pc, H, W, t, mask, im, sharp, weight, threshold, st))) return rc;   // blur, cols + mask/sharp
        if ((rc = launch_h(mask, pc, H, W, t, tmp, st))) return rc;                                        // soft mask, rows
        if ((rc = launch_v<EPI_USM_BLEND>(tmp, pc, H, W, t, o, im, sharp, weight, threshold, st))) return rc;     // soft mask, cols + blend
    }
    return OTF_OK;
}

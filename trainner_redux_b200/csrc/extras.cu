// Fork extras: the tensor-math stages that this fork's feed_data runs around the
// classical primitives — traiNNer/models/paragon_otf_degradations.py:251-572 and the model's own copies
// (traiNNer/models/realesrgan_model.py:193-402).  All of them are HBM-bound pointwise / gather work:
//   warp_kernel          lens distortion, rolling shutter, chromatic aberration  (analytic grid + grid_sample)
//   taps_zero_kernel     motion blur (line kernel) and the 5x5 box of "oversharpen"   (conv2d, ZERO padding)
//   channel_gain_kernel  exposure, colour temperature, editing exposure          (x * g[c], clamp)
//   sensor_noise_kernel  x + N * std, clamp                                       (Philox or injected field)
// Legacy `nearest` resampling for the aliasing stage lives in resize.cu (OTF_RESIZE_NEAREST).
//
// The warp reproduces ATen's fp32 arithmetic operation by operation (checked bit for bit against the CPU
// build in oracle/paragon_oracle.py): torch.linspace = fma(step, i, -1) below the midpoint and
// fma(-step, n-1-i, 1) above it; grid_sample's unnormalise = fma(g + 1, size / 2, -0.5); bilinear sum =
// fma chain nw, ne, sw, se.  One ulp of a source coordinate at x >= 128 is 1.5e-5 of a pixel, so anything
// looser than that would show up against the 1e-5 bar on noisy images.
#include <stdlib.h>

#include "otf_common.cuh"

namespace otf {

__device__ __forceinline__ float linspace_pm1(int i, int n, float step) {
    // at::linspace(-1, 1, n) on fp32 (aten/src/ATen/native/cpu/RangeFactoriesKernel.cpp)
    return i < n / 2 ? __fmaf_rn(step, (float)i, -1.0f) : __fmaf_rn(-step, (float)(n - 1 - i), 1.0f);
}

// grid_sample source coordinate, align_corners=False (GridSamplerKernel.cpp ComputeLocation)
__device__ __forceinline__ float gs_unnormalise(float g, int size) {
    return __fmaf_rn(__fadd_rn(g, 1.0f), __fdiv_rn((float)size, 2.0f), -0.5f);
}
__device__ __forceinline__ float gs_reflect_clip(float x, int size) {
    const float low = -0.5f, twice_span = __fmul_rn((float)size, 2.0f);
    const float a = fabsf(__fsub_rn(x, low));
    // a < twice_span => trunc(fl(a / twice_span)) == 0 (the quotient cannot round up to 1): skip the IEEE division
    const float flips = a < twice_span ? 0.0f : truncf(__fdiv_rn(a, twice_span));
    const float extra = __fsub_rn(a, __fmul_rn(flips, twice_span));
    const float r = fminf(__fadd_rn(extra, low), __fadd_rn(__fsub_rn(twice_span, extra), low));
    return fminf(fmaxf(r, 0.0f), (float)(size - 1));
}

struct WarpParams {
    int mode;       // OTF_WARP_*
    float p0;       // lens strength | shutter slant
    float step_h, step_w;  // linspace steps 2/(H-1), 2/(W-1) as fp32
};

// The four taps of one bilinear sample with zeros padding, split from the per-channel gather: offsets (clamped into the
// plane so that every load can issue unconditionally), in-bounds flags and weights are per POSITION, and the C channels
// of a pixel share them — the loads of all channels are independent and in flight together.
struct BilinearTaps {
    int o00, o01, o10, o11;
    bool m00, m01, m10, m11;
    float nw, ne, sw, se;
    __device__ __forceinline__ BilinearTaps(int H, int W, float x, float y) {
        const float xw = floorf(x), yn = floorf(y);
        const float xe = __fadd_rn(xw, 1.0f), ys = __fadd_rn(yn, 1.0f);
        const float wxe = __fsub_rn(xe, x), wxw = __fsub_rn(x, xw), wys = __fsub_rn(ys, y), wyn = __fsub_rn(y, yn);
        nw = __fmul_rn(wxe, wys), ne = __fmul_rn(wxw, wys), sw = __fmul_rn(wxe, wyn), se = __fmul_rn(wxw, wyn);
        const int ix = (int)xw, iy = (int)yn;
        const bool x0 = ix >= 0 && ix < W, x1 = ix + 1 >= 0 && ix + 1 < W, y0 = iy >= 0 && iy < H, y1 = iy + 1 >= 0 && iy + 1 < H;
        const int cx0 = min(max(ix, 0), W - 1), cx1 = min(max(ix + 1, 0), W - 1);
        const int cy0 = min(max(iy, 0), H - 1) * W, cy1 = min(max(iy + 1, 0), H - 1) * W;
        o00 = cy0 + cx0, o01 = cy0 + cx1, o10 = cy1 + cx0, o11 = cy1 + cx1;
        m00 = x0 && y0, m01 = x1 && y0, m10 = x0 && y1, m11 = x1 && y1;
    }
    __device__ __forceinline__ float sample(const float* __restrict__ pl) const {
        const float va = __ldg(pl + o00), vb = __ldg(pl + o01), vc = __ldg(pl + o10), vd = __ldg(pl + o11);
        const float a = m00 ? va : 0.0f, b = m01 ? vb : 0.0f, c = m10 ? vc : 0.0f, d = m11 ? vd : 0.0f;
        float r = __fmul_rn(a, nw);
        r = __fmaf_rn(b, ne, r);
        r = __fmaf_rn(c, sw, r);
        return __fmaf_rn(d, se, r);
    }
};
__device__ __forceinline__ float bilinear_zero(const float* __restrict__ pl, int H, int W, float x, float y) {
    return BilinearTaps(H, W, x, y).sample(pl);
}

// thread = output pixel (j fastest); loops over the C channels of its sample.  One instantiation per mode: the
// three grids share little code, and a mode-specific kernel needs two thirds of the registers of the combined one.
template <int MODE>
__global__ void __launch_bounds__(256, 6) warp_kernel(const float* __restrict__ img, float* __restrict__ out, int C, int H, int W,
                                                   WarpParams p) {
    // 16 x 16 output tile per CTA: the lens grid samples the TRANSPOSED position (see below), so a flat row of threads
    // would read one source column (32 sectors per load); a square tile touches a square source patch either way
    const int j = blockIdx.x * 16 + (threadIdx.x & 15), i = blockIdx.y * 16 + (threadIdx.x >> 4), b = blockIdx.z;
    if (j >= W || i >= H) return;
    const float lh = linspace_pm1(i, H, p.step_h), lw = linspace_pm1(j, W, p.step_w);
    const size_t plane = (size_t)H * W;
    const float* ip = img + (size_t)b * C * plane;
    float* op = out + (size_t)b * C * plane + (size_t)i * W + j;
    if (MODE == OTF_WARP_CHROMA) {
        // affine_grid(align_corners=False): base = linspace * (n-1) / n, then base * scale (theta is diagonal);
        // R sampled at scale 1.001, G untouched, B at 0.999; zeros padding; clamp(0,1) on all three
        const float bx = __fdiv_rn(__fmul_rn(lw, (float)(W - 1)), (float)W), by = __fdiv_rn(__fmul_rn(lh, (float)(H - 1)), (float)H);
        for (int c = 0; c < C; ++c) {
            float v;
            if (C == 3 && c != 1) {
                const float s = c == 0 ? 1.001f : 0.999f;
                v = bilinear_zero(ip + c * plane, H, W, gs_unnormalise(__fmul_rn(bx, s), W), gs_unnormalise(__fmul_rn(by, s), H));
            } else {
                v = __ldg(ip + c * plane + (size_t)i * W + j);
            }
            op[c * plane] = clamp01(v);
        }
        return;
    }
    float gx, gy;
    if (MODE == OTF_WARP_LENS) {
        // paragon_otf_degradations.py:313-331 — note grid_x runs along the ROWS (meshgrid 'ij' of (H, W)) and is
        // stacked as the x coordinate: the reference samples the transposed position, reproduced as is
        const float r = __fsqrt_rn(__fadd_rn(__fmul_rn(lh, lh), __fmul_rn(lw, lw)));
        float rd = __fmul_rn(r, __fadd_rn(1.0f, __fmul_rn(p.p0, __fmul_rn(r, r))));
        float rr = r;
        if (r == 0.0f) { rd = 0.0f; rr = 1e-6f; }
        const float ratio = __fdiv_rn(rd, rr);
        gx = __fmul_rn(lh, ratio);
        gy = __fmul_rn(lw, ratio);
    } else {  // OTF_WARP_SHUTTER: :440-447 — x + slant * y, y
        gx = __fadd_rn(lw, __fmul_rn(p.p0, lh));
        gy = lh;
    }
    const float x = gs_reflect_clip(gs_unnormalise(gx, W), W), y = gs_reflect_clip(gs_unnormalise(gy, H), H);
    const BilinearTaps taps(H, W, x, y);
    if (C == 3) {
        const float v0 = taps.sample(ip), v1 = taps.sample(ip + plane), v2 = taps.sample(ip + 2 * plane);
        op[0] = v0, op[plane] = v1, op[2 * plane] = v2;
        return;
    }
    for (int c = 0; c < C; ++c) op[c * plane] = taps.sample(ip + c * plane);
}

// Lens distortion, coalesced on both sides.  The reference's grid samples the TRANSPOSED position (output (i, j) reads the
// source near column i, row j), so with a thread per output pixel in row-major order a warp reads one source COLUMN
// (up to 32 sectors per load) — the gather was bound by L1 sector throughput at 0.17 of the HBM roof.  Here a CTA owns a
// 32 x 32 output tile and walks it column-major for the gathers (lanes = consecutive output ROWS = consecutive source
// columns: 4 sectors per load), parks the results in a shared-memory tile, and writes the tile out row-major.  The per-pixel
// arithmetic is warp_kernel<OTF_WARP_LENS>'s, operation for operation.
// (one output per thread; 32 rows x 16 columns per 512-thread CTA at <= 40 registers, three CTAs = 48 warps resident: the
// gather is latency-bound — a 256-thread x 4-output version at 94 / 64 registers measured 0.087 / 0.064 ms, 1024 threads
// capped at 32 registers 0.079 ms)
constexpr int LENS_TI = 32, LENS_TJ = 16;
__global__ void __launch_bounds__(LENS_TI * LENS_TJ, 3) warp_lens_tile_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                                              int C, int H, int W, WarpParams p) {
    __shared__ float tile[3][LENS_TI][LENS_TJ + 1];
    const int i0 = blockIdx.y * LENS_TI, j0 = blockIdx.x * LENS_TJ, b = blockIdx.z;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const size_t plane = (size_t)H * W;
    const float* ip = img + (size_t)b * C * plane;
    float* op = out + (size_t)b * C * plane;
    float sx, sy;  // source coordinates of output (i0 + lane, j0 + wrp)
    {
        const int i = min(i0 + lane, H - 1), j = min(j0 + wrp, W - 1);
        const float lh = linspace_pm1(i, H, p.step_h), lw = linspace_pm1(j, W, p.step_w);
        const float r = __fsqrt_rn(__fadd_rn(__fmul_rn(lh, lh), __fmul_rn(lw, lw)));
        float rd = __fmul_rn(r, __fadd_rn(1.0f, __fmul_rn(p.p0, __fmul_rn(r, r))));
        float rr = r;
        if (r == 0.0f) { rd = 0.0f; rr = 1e-6f; }
        const float ratio = __fdiv_rn(rd, rr);
        sx = gs_reflect_clip(gs_unnormalise(__fmul_rn(lh, ratio), W), W);
        sy = gs_reflect_clip(gs_unnormalise(__fmul_rn(lw, ratio), H), H);
    }
    const BilinearTaps taps(H, W, sx, sy);
    const int ti = threadIdx.x / LENS_TJ, tj = threadIdx.x % LENS_TJ;  // store phase: 16 consecutive columns of two rows per warp
    const int oi = i0 + ti, oj = j0 + tj;
    const bool store = oi < H && oj < W;
    for (int c0 = 0; c0 < C; c0 += 3) {  // three channels per barrier pair
        const int nc = min(3, C - c0);
        if (nc == 3) {  // twelve independent loads in flight
            const float* q = ip + c0 * plane;
            const float v0 = taps.sample(q), v1 = taps.sample(q + plane), v2 = taps.sample(q + 2 * plane);
            tile[0][lane][wrp] = v0, tile[1][lane][wrp] = v1, tile[2][lane][wrp] = v2;  // [row i][col j]
        } else {
            for (int c = 0; c < nc; ++c) tile[c][lane][wrp] = taps.sample(ip + (c0 + c) * plane);
        }
        __syncthreads();
        if (store)
            for (int c = 0; c < nc; ++c) op[(c0 + c) * plane + (size_t)oi * W + oj] = tile[c][ti][tj];
        if (c0 + 3 < C) __syncthreads();
    }
}

// ---- small correlation with ZERO padding (F.conv2d(padding=K//2, groups=C)) ------------------------
struct TapList {
    int n;
    float w[OTF_MAX_TAPS];
    int8_t dy[OTF_MAX_TAPS], dx[OTF_MAX_TAPS];
};

__global__ void __launch_bounds__(256) taps_zero_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W,
                                                        int OH, int OW, const __grid_constant__ TapList taps, int epilogue,
                                                        float strength) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= OW || y >= OH) return;
    const float* ip = img + (size_t)blockIdx.z * H * W;
    float acc = 0.0f;
    for (int k = 0; k < taps.n; ++k) {
        const int yy = y + taps.dy[k], xx = x + taps.dx[k];
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc = fmaf(taps.w[k], __ldg(ip + (size_t)yy * W + xx), acc);
    }
    if (epilogue == OTF_TAPS_OVERSHARPEN) {
        // img + (img - blurred) * strength, clamp(0,1) — paragon_otf_degradations.py:480-482 (separate ATen ops)
        const float v = __ldg(ip + (size_t)y * W + x);
        acc = clamp01(__fadd_rn(v, __fmul_rn(__fsub_rn(v, acc), strength)));
    }
    out[(size_t)blockIdx.z * OH * OW + (size_t)y * OW + x] = acc;
}

// Tiled variant for K <= 31: the (tile + halo) source patch is staged in shared memory with the zero padding written
// in, so the tap loop carries no bounds checks; a thread owns 4 adjacent columns x 2 rows and every tap costs it
// 8 LDS + 8 FFMA (tap offsets and weights are warp-uniform reads of the parameter bank).
constexpr int kTapTW = 64, kTapTH = 32, kTapMaxR = 15;
__global__ void __launch_bounds__(256) taps_zero_tile_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W,
                                                             int OH, int OW, const __grid_constant__ TapList taps, int R,
                                                             int epilogue, float strength) {
    extern __shared__ float tile[];
    const int P = kTapTW + 2 * R + 1;  // odd pitch
    const int rows = kTapTH + 2 * R;
    const int ox0 = blockIdx.x * kTapTW, oy0 = blockIdx.y * kTapTH;
    const float* ip = img + (size_t)blockIdx.z * H * W;
    for (int idx = threadIdx.x; idx < rows * (kTapTW + 2 * R); idx += 256) {
        const int r = idx / (kTapTW + 2 * R), c = idx - r * (kTapTW + 2 * R);
        const int yy = oy0 - R + r, xx = ox0 - R + c;
        tile[r * P + c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(ip + (size_t)yy * W + xx) : 0.0f;
    }
    __syncthreads();
    const int tx = (threadIdx.x & 15) * 4, ty = threadIdx.x >> 4;  // rows ty and ty + 16
    float acc[2][4] = {};
    const float* base = tile + (ty + R) * P + tx + R;
    for (int k = 0; k < taps.n; ++k) {
        const float w = taps.w[k];
        const float* tp = base + taps.dy[k] * P + taps.dx[k];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            acc[0][c] = fmaf(w, tp[c], acc[0][c]);
            acc[1][c] = fmaf(w, tp[16 * P + c], acc[1][c]);
        }
    }
    float* op = out + (size_t)blockIdx.z * OH * OW;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int y = oy0 + ty + 16 * r;
        if (y >= OH) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int x = ox0 + tx + c;
            if (x >= OW) continue;
            float a = acc[r][c];
            if (epilogue == OTF_TAPS_OVERSHARPEN) {
                const float v = base[(16 * r) * P + c];  // centre pixel (OH == H, OW == W for odd K)
                a = clamp01(__fadd_rn(v, __fmul_rn(__fsub_rn(v, a), strength)));
            }
            op[(size_t)y * OW + x] = a;
        }
    }
}

// ---- "oversharpen": 5x5 box (zero padding) + clamp(img + (img - box) * strength) as one streaming pass ---------------
// The 25 equal taps are evaluated separably with running sums — 5-tap horizontal sums of each source row (from one
// float4 + two float2 loads), a rolling window of the last five of them down the column strip — instead of 25 LDS + 25
// FFMA per output through the generic tap list: the pass is a read and a write of the plane (HBM bound).  The sum
// w * (25 pixels) replaces the reference's sum of 25 products w * p (paragon_otf_degradations.py:472-479): <= 3e-7 apart.
constexpr int kBoxStrip = 16;
__global__ void __launch_bounds__(256) box5_sharpen_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W,
                                                           float w, float strength) {
    const int quad = blockIdx.x * 64 + (threadIdx.x & 63), strip = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int x = 4 * quad, y0 = strip * kBoxStrip;
    if (x >= W || y0 >= H) return;
    const float* ip = img + (size_t)blockIdx.z * H * W;
    float* op = out + (size_t)blockIdx.z * H * W;
    float h[5][4], c[3][4];  // horizontal sums of rows y-2..y+2, centre pixels of rows y..y+2 (y = the row being written)
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) h[i][k] = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) c[i][k] = 0.0f;
    const int y_end = min(y0 + kBoxStrip, H);
#pragma unroll 4
    for (int yy = y0 - 2; yy < y_end + 2; ++yy) {
        float p[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // columns x-2 .. x+5, zero outside the plane
        if (yy >= 0 && yy < H) {
            const float* rp = ip + (size_t)yy * W + x;
            const float4 m = __ldg(reinterpret_cast<const float4*>(rp));
            p[2] = m.x; p[3] = m.y; p[4] = m.z; p[5] = m.w;
            if (x >= 2) { const float2 l = __ldg(reinterpret_cast<const float2*>(rp - 2)); p[0] = l.x; p[1] = l.y; }
            if (x + 4 < W) { const float2 r = __ldg(reinterpret_cast<const float2*>(rp + 4)); p[6] = r.x; p[7] = r.y; }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) h[i][k] = h[i + 1][k];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) c[i][k] = c[i + 1][k];
        const float s23 = p[2] + p[3], s45 = p[4] + p[5];
        h[4][0] = (p[0] + p[1]) + s23 + p[4];
        h[4][1] = (p[1] + s23) + s45;
        h[4][2] = (s23 + s45) + p[6];
        h[4][3] = (p[3] + s45) + (p[6] + p[7]);
#pragma unroll
        for (int k = 0; k < 4; ++k) c[2][k] = p[2 + k];
        const int y = yy - 2;  // rows y-2 .. y+2 are in h[0..4], its centre pixels in c[0]
        if (y >= y0) {
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float blurred = __fmul_rn(((h[0][k] + h[1][k]) + (h[2][k] + h[3][k])) + h[4][k], w);
                const float v = c[0][k];
                o[k] = clamp01(__fadd_rn(v, __fmul_rn(__fsub_rn(v, blurred), strength)));
            }
            *reinterpret_cast<float4*>(op + (size_t)y * W + x) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---- x * gain[c], optional clamp --------------------------------------------------------------------
__global__ void __launch_bounds__(256) channel_gain_kernel(const float4* __restrict__ img, float4* __restrict__ out, int C,
                                                           int64_t quads_per_plane, float g0, float g1, float g2, int clamp_out) {
    const int plane = blockIdx.y, c = plane % C;
    const float g = c == 0 ? g0 : c == 1 ? g1 : g2;
    const float4* ip = img + (size_t)plane * quads_per_plane;
    float4* op = out + (size_t)plane * quads_per_plane;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < quads_per_plane; q += (int64_t)gridDim.x * 256) {
        float4 v = __ldg(ip + q);
        v.x = __fmul_rn(v.x, g); v.y = __fmul_rn(v.y, g); v.z = __fmul_rn(v.z, g); v.w = __fmul_rn(v.w, g);
        if (clamp_out) { v.x = clamp01(v.x); v.y = clamp01(v.y); v.z = clamp01(v.z); v.w = clamp01(v.w); }
        op[q] = v;
    }
}
__global__ void __launch_bounds__(256) channel_gain_scalar_kernel(const float* __restrict__ img, float* __restrict__ out, int C,
                                                                  int64_t hw, float g0, float g1, float g2, int clamp_out) {
    const int plane = blockIdx.y, c = plane % C;
    const float g = c == 0 ? g0 : c == 1 ? g1 : g2;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < hw; q += (int64_t)gridDim.x * 256) {
        const float v = __fmul_rn(__ldg(img + (size_t)plane * hw + q), g);
        out[(size_t)plane * hw + q] = clamp_out ? clamp01(v) : v;
    }
}

// ---- clamp(x + N * std, 0, 1) — paragon_otf_degradations.py:411-414 ---------------------------------
__global__ void __launch_bounds__(256) sensor_noise_kernel(const float* __restrict__ img, const float* __restrict__ noise,
                                                           float* __restrict__ out, int64_t n, float std, uint64_t seed,
                                                           uint64_t offset) {
    const Philox ph(seed);
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; 4 * q < n; q += (int64_t)gridDim.x * 256) {
        float nv[4];
        if (noise) {
            for (int k = 0; k < 4; ++k) nv[k] = 4 * q + k < n ? __ldg(noise + 4 * q + k) : 0.0f;
        } else {
            const float4 z = normal4(ph, (uint64_t)q, offset * 8 + STREAM_COLOR);
            nv[0] = z.x; nv[1] = z.y; nv[2] = z.z; nv[3] = z.w;
        }
        for (int k = 0; k < 4; ++k)
            if (4 * q + k < n) out[4 * q + k] = clamp01(__fadd_rn(__ldg(img + 4 * q + k), __fmul_rn(nv[k], std)));
    }
}

// ---- Bayer mosaic + bilinear demosaic — paragon_otf_degradations.py:526-552 ---------------------------
// The reference quantises to uint8 (truncation), keeps ONE channel per pixel (rows/cols even-even: channel 2,
// odd-odd: channel 0, mixed parity: channel 1) and calls cv2.demosaicing(COLOR_BAYER_BG2BGR) on the host, image by
// image.  OpenCV's bilinear demosaic is integer arithmetic: the two missing channels of a pixel are the rounded-up
// mean of their 2 (horizontal / vertical) or 4 (cross / diagonal) nearest mosaic samples, (a+b+1)>>1, (a+b+c+d+2)>>2;
// the outermost rows and columns copy their inner neighbours; images with fewer than 3 rows or columns come out
// zero.  Reproduced bit for bit (pinned against cv2 in tests/golden/paragon_goldens.npz).
__device__ __forceinline__ int bayer_at(const float* __restrict__ ip, size_t plane, int W, int y, int x) {
    const int c = (y & 1) ? ((x & 1) ? 0 : 1) : ((x & 1) ? 1 : 2);
    return (int)floorf(__fmul_rn(clamp01(__ldg(ip + c * plane + (size_t)y * W + x)), 255.0f));
}
__global__ void __launch_bounds__(256) demosaic_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6), b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W;
    const float* ip = img + (size_t)b * 3 * plane;
    float* op = out + (size_t)b * 3 * plane + (size_t)y * W + x;
    if (H < 3 || W < 3) {
        op[0] = op[plane] = op[2 * plane] = 0.0f;
        return;
    }
    const int yc = min(max(y, 1), H - 2), xc = min(max(x, 1), W - 2);  // border pixels copy their inner neighbour
    int v[3][3];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) v[dy][dx] = bayer_at(ip, plane, W, yc + dy - 1, xc + dx - 1);
    const int ctr = v[1][1];
    const int hor2 = (v[1][0] + v[1][2] + 1) >> 1, ver2 = (v[0][1] + v[2][1] + 1) >> 1;
    const int cross4 = (v[0][1] + v[2][1] + v[1][0] + v[1][2] + 2) >> 2;
    const int diag4 = (v[0][0] + v[0][2] + v[2][0] + v[2][2] + 2) >> 2;
    const bool ey = !(yc & 1), ex = !(xc & 1);
    const int c2 = ey ? (ex ? ctr : hor2) : (ex ? ver2 : diag4);   // channel 2 lives on (even, even)
    const int c0 = !ey ? (!ex ? ctr : hor2) : (!ex ? ver2 : diag4);  // channel 0 on (odd, odd)
    const int c1 = (ey != ex) ? ctr : cross4;                      // channel 1 on mixed parity
    op[0] = __fdiv_rn((float)c0, 255.0f);
    op[plane] = __fdiv_rn((float)c1, 255.0f);
    op[2 * plane] = __fdiv_rn((float)c2, 255.0f);
}

// floor(clamp(x,0,1) * 255) / 255 — the `(img * 255).astype("uint8")` in front of every codec round (:114-115)
__global__ void __launch_bounds__(256) trunc8_kernel(const float* __restrict__ img, float* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        out[i] = __fdiv_rn(floorf(__fmul_rn(clamp01(__ldg(img + i)), 255.0f)), 255.0f);
}

}  // namespace otf

extern "C" int otf_demosaic_f32(const float* img, int B, int H, int W, float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && img != out, OTF_ERR_BAD_ARG, "demosaic: bad pointers");
    OTF_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, OTF_ERR_BAD_ARG, "demosaic: bad extents");
    demosaic_kernel<<<dim3(ceil_div(W, 64), ceil_div(H, 4), B), 256, 0, (cudaStream_t)stream>>>(img, out, H, W);
    OTF_LAUNCH_CHECK("demosaic_kernel");
    return OTF_OK;
}

extern "C" int otf_trunc8_f32(const float* img, int64_t n, float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && n > 0, OTF_ERR_BAD_ARG, "trunc8: bad arguments");
    int grid = ceil_div(n, 256 * 4);
    if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
    trunc8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, out, n);
    OTF_LAUNCH_CHECK("trunc8_kernel");
    return OTF_OK;
}

extern "C" int otf_warp_f32(const float* img, int B, int C, int H, int W, int mode, float p0, float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && img != out, OTF_ERR_BAD_ARG, "warp: bad pointers");
    OTF_REQUIRE(B > 0 && B <= 65535 && C > 0 && H > 1 && W > 1, OTF_ERR_BAD_ARG, "warp: bad extents");
    OTF_REQUIRE(mode >= OTF_WARP_LENS && mode <= OTF_WARP_CHROMA, OTF_ERR_BAD_ARG, "warp: unknown mode %d", mode);
    WarpParams p;
    p.mode = mode;
    p.p0 = p0;
    p.step_h = 2.0f / (float)(H - 1);
    p.step_w = 2.0f / (float)(W - 1);
    const dim3 grid(ceil_div(W, 16), ceil_div(H, 16), B);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool lens_flat = getenv("OTF_LENS_FLAT") != nullptr;  // A/B switch: the thread-per-pixel kernel
    if (mode == OTF_WARP_LENS && !lens_flat)
        warp_lens_tile_kernel<<<dim3(ceil_div(W, LENS_TJ), ceil_div(H, LENS_TI), B), LENS_TI * LENS_TJ, 0, st>>>(img, out, C, H, W, p);
    else if (mode == OTF_WARP_LENS) warp_kernel<OTF_WARP_LENS><<<grid, 256, 0, st>>>(img, out, C, H, W, p);
    else if (mode == OTF_WARP_SHUTTER) warp_kernel<OTF_WARP_SHUTTER><<<grid, 256, 0, st>>>(img, out, C, H, W, p);
    else warp_kernel<OTF_WARP_CHROMA><<<grid, 256, 0, st>>>(img, out, C, H, W, p);
    OTF_LAUNCH_CHECK("warp_kernel");
    return OTF_OK;
}

extern "C" int otf_taps_zero_f32(const float* img, int planes, int H, int W, int K, const float* kernel_host, int epilogue,
                                 float strength, float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && kernel_host && img != out, OTF_ERR_BAD_ARG, "taps_zero: bad pointers");
    OTF_REQUIRE(planes > 0 && planes <= 65535 && H > 0 && W > 0 && K > 0 && K <= 127, OTF_ERR_BAD_ARG, "taps_zero: bad extents");
    OTF_REQUIRE(epilogue == OTF_TAPS_NONE || (epilogue == OTF_TAPS_OVERSHARPEN && (K & 1)), OTF_ERR_BAD_ARG, "taps_zero: epilogue");
    TapList t;
    t.n = 0;
    const int pad = K / 2;  // F.conv2d(padding=K//2): an even K grows the image by one row and column, as in the reference
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            const float w = kernel_host[i * K + j];
            if (w == 0.0f) continue;
            OTF_REQUIRE(t.n < OTF_MAX_TAPS, OTF_ERR_UNSUPPORTED, "taps_zero: more than %d non-zero taps", OTF_MAX_TAPS);
            t.w[t.n] = w;
            t.dy[t.n] = (int8_t)(i - pad);
            t.dx[t.n] = (int8_t)(j - pad);
            ++t.n;
        }
    const int OH = H + 2 * pad - K + 1, OW = W + 2 * pad - K + 1;
    bool box5 = epilogue == OTF_TAPS_OVERSHARPEN && K == 5 && t.n == 25 && (W & 3) == 0 && (((uintptr_t)img | (uintptr_t)out) & 15) == 0;
    for (int i = 1; box5 && i < 25; ++i) box5 = t.w[i] == t.w[0];
    if (box5) {  // the oversharpen stage as the reference configures it: 25 equal taps
        box5_sharpen_kernel<<<dim3(ceil_div(W / 4, 64), ceil_div(ceil_div(H, kBoxStrip), 4), planes), 256, 0, (cudaStream_t)stream>>>(
            img, out, H, W, t.w[0], strength);
        OTF_LAUNCH_CHECK("box5_sharpen_kernel");
        return OTF_OK;
    }
    if (pad <= kTapMaxR) {
        const size_t smem = (size_t)(kTapTH + 2 * pad) * (kTapTW + 2 * pad + 1) * sizeof(float);
        taps_zero_tile_kernel<<<dim3(ceil_div(OW, kTapTW), ceil_div(OH, kTapTH), planes), 256, smem, (cudaStream_t)stream>>>(
            img, out, H, W, OH, OW, t, pad, epilogue, strength);
        OTF_LAUNCH_CHECK("taps_zero_tile_kernel");
        return OTF_OK;
    }
    taps_zero_kernel<<<dim3(ceil_div(OW, 64), ceil_div(OH, 4), planes), 256, 0, (cudaStream_t)stream>>>(img, out, H, W, OH, OW, t,
                                                                                                     epilogue, strength);
    OTF_LAUNCH_CHECK("taps_zero_kernel");
    return OTF_OK;
}

extern "C" int otf_channel_gain_f32(const float* img, int B, int C, int64_t hw, float g0, float g1, float g2, int clamp01_out,
                                    float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out, OTF_ERR_BAD_ARG, "channel_gain: bad pointers");
    OTF_REQUIRE(B > 0 && C > 0 && hw > 0 && (int64_t)B * C <= 65535, OTF_ERR_BAD_ARG, "channel_gain: bad extents");
    const bool vec = hw % 4 == 0 && (((uintptr_t)img | (uintptr_t)out) & 15) == 0;
    const int64_t units = vec ? hw / 4 : hw;
    int gx = ceil_div(units, 256 * 4);
    if (gx < 1) gx = 1;
    if (gx > 4 * kNumSMs) gx = 4 * kNumSMs;
    if (vec)
        channel_gain_kernel<<<dim3(gx, B * C), 256, 0, (cudaStream_t)stream>>>((const float4*)img, (float4*)out, C, units, g0, g1, g2,
                                                                             clamp01_out);
    else
        channel_gain_scalar_kernel<<<dim3(gx, B * C), 256, 0, (cudaStream_t)stream>>>(img, out, C, hw, g0, g1, g2, clamp01_out);
    OTF_LAUNCH_CHECK("channel_gain_kernel");
    return OTF_OK;
}

extern "C" int otf_sensor_noise_f32(const float* img, int64_t n, float std, const float* noise_dev, uint64_t seed, uint64_t offset,
                                    float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && n > 0, OTF_ERR_BAD_ARG, "sensor_noise: bad arguments");
    int grid = ceil_div((n + 3) / 4, 256 * 2);
    if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
    sensor_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, noise_dev, out, n, std, seed, offset);
    OTF_LAUNCH_CHECK("sensor_noise_kernel");
    return OTF_OK;
}

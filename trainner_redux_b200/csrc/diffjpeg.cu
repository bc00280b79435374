// Fused DiffJPEG: one kernel for the whole of traiNNer/utils/diffjpeg.py:503-527
// (x255 -> RGB2YCbCr -> 4:2:0 -> 8x8 DCT -> quantise by table*factor[b] -> round -> dequantise ->
//  IDCT -> chroma x2 -> YCbCr2RGB -> clamp -> /255 -> crop), which the reference runs as ~60 ATen
// launches and ~25 image-sized HBM round trips.  Here the image is read once and written once.
//
// Mapping: one warp = one 16x16 MCU (4 Y blocks + Cb + Cr) held in the warp's shared memory; the separable 8x8 DCT runs
// as (block, row) and (block, column) tasks over the lanes (see diffjpeg_kernel below).
// Arithmetic order (documented for parity): 1-D DCT-II with the orthonormal matrix
// C[k][n] = 0.5*alpha_k*cos((2n+1)k*pi/16) evaluated by even/odd decomposition; quantisation divides by
// fl(table*factor) as the reference does (quotient within 1 ulp of IEEE), torch.round = rintf (half to even).
#include <string.h>

#include "otf_common.cuh"

namespace otf {

// Quantisation tables indexed [u][v] exactly as the reference stores them: the Annex-K luminance
// table TRANSPOSED (diffjpeg.py:18-31) and the chroma table (:32-37, symmetric).
__device__ const float c_ytab[8][8] = {
    {16, 12, 14, 14, 18, 24, 49, 72},     {11, 12, 13, 17, 22, 35, 64, 92},   {10, 14, 16, 22, 37, 55, 78, 95},
    {16, 19, 24, 29, 56, 64, 87, 98},     {24, 26, 40, 51, 68, 81, 103, 112}, {40, 58, 57, 87, 109, 104, 121, 100},
    {51, 60, 69, 80, 103, 113, 120, 103}, {61, 55, 56, 62, 77, 92, 101, 99},
};
__device__ const float c_ctab[8][8] = {
    {17, 18, 24, 47, 99, 99, 99, 99}, {18, 21, 26, 66, 99, 99, 99, 99}, {24, 26, 56, 99, 99, 99, 99, 99},
    {47, 66, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99},
    {99, 99, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99},
};

// 8-point orthonormal DCT-II, C[k][n] = 0.5*alpha_k*cos((2n+1)k*pi/16), by even/odd decomposition (the reference's
// 0.25*alpha_u*alpha_v*cos*cos tensor of diffjpeg.py:155-164 is the outer product of two of these): the mirror
// symmetry C[k][7-n] = (-1)^k C[k][n] halves the sums, the even half splits once more — 36 FP32 operations instead of
// the 64 FMAs of the dense matrix.  Rounding differs from the dense order by a few ulp of the coefficients (~1e-4 on a
// 0..2040 scale), i.e. by as much as the dense order differs from the reference's 64-term tensordot.
#define OTF_C4 3.535533845e-01f
#define OTF_C1 4.903926253e-01f
#define OTF_C2 4.619397521e-01f
#define OTF_C3 4.157347977e-01f
#define OTF_C5 2.777851224e-01f
#define OTF_C6 1.913417131e-01f
#define OTF_C7 9.754516184e-02f
// y[k] = sum_n C[k][n] x[n]
__device__ __forceinline__ void dct8(float (&x)[8]) {
    const float s0 = x[0] + x[7], s1 = x[1] + x[6], s2 = x[2] + x[5], s3 = x[3] + x[4];
    const float d0 = x[0] - x[7], d1 = x[1] - x[6], d2 = x[2] - x[5], d3 = x[3] - x[4];
    const float e0 = s0 + s3, e1 = s1 + s2, f0 = s0 - s3, f1 = s1 - s2;
    x[0] = OTF_C4 * (e0 + e1);
    x[4] = OTF_C4 * (e0 - e1);
    x[2] = fmaf(OTF_C6, f1, OTF_C2 * f0);
    x[6] = fmaf(-OTF_C2, f1, OTF_C6 * f0);
    x[1] = fmaf(OTF_C7, d3, fmaf(OTF_C5, d2, fmaf(OTF_C3, d1, OTF_C1 * d0)));
    x[3] = fmaf(-OTF_C5, d3, fmaf(-OTF_C1, d2, fmaf(-OTF_C7, d1, OTF_C3 * d0)));
    x[5] = fmaf(OTF_C3, d3, fmaf(OTF_C7, d2, fmaf(-OTF_C1, d1, OTF_C5 * d0)));
    x[7] = fmaf(-OTF_C1, d3, fmaf(OTF_C3, d2, fmaf(-OTF_C5, d1, OTF_C7 * d0)));
}
// x[n] = sum_k C[k][n] y[k]
__device__ __forceinline__ void idct8(float (&y)[8]) {
    const float a0 = OTF_C4 * y[0], a4 = OTF_C4 * y[4];
    const float p0 = a0 + a4, p1 = a0 - a4;
    const float q0 = fmaf(OTF_C6, y[6], OTF_C2 * y[2]), q1 = fmaf(-OTF_C2, y[6], OTF_C6 * y[2]);
    const float e0 = p0 + q0, e3 = p0 - q0, e1 = p1 + q1, e2 = p1 - q1;
    const float o0 = fmaf(OTF_C7, y[7], fmaf(OTF_C5, y[5], fmaf(OTF_C3, y[3], OTF_C1 * y[1])));
    const float o1 = fmaf(-OTF_C5, y[7], fmaf(-OTF_C1, y[5], fmaf(-OTF_C7, y[3], OTF_C3 * y[1])));
    const float o2 = fmaf(OTF_C3, y[7], fmaf(OTF_C7, y[5], fmaf(-OTF_C1, y[3], OTF_C5 * y[1])));
    const float o3 = fmaf(-OTF_C1, y[7], fmaf(OTF_C3, y[5], fmaf(-OTF_C5, y[3], OTF_C7 * y[1])));
    y[0] = e0 + o0; y[7] = e0 - o0;
    y[1] = e1 + o1; y[6] = e1 - o1;
    y[2] = e2 + o2; y[5] = e2 - o2;
    y[3] = e3 + o3; y[4] = e3 - o3;
}

// (div255 — x / 255 with IEEE rounding in three FP32-pipe operations — lives in otf_common.cuh: the libjpeg round uses it too)
// clamp(round(x * 255), 0, 255) / 255 for x already in [0, 1] (quantise8 of otf_common.cuh with the cheap division)
__device__ __forceinline__ float quantise8_unit(float x) { return div255(fminf(fmaxf(rintf(__fmul_rn(x, 255.0f)), 0.0f), 255.0f)); }

// Quantise + dequantise an 8x8 block held as: lane v (within its group of 8) owns coefficients D[u][v], u = register
// index; t[u] = fl(table[u][v] * factor) (diffjpeg.py:207-212 divides by exactly this product).  The quotient is
// d * rcp(t) corrected once with the exact residual: within 1 ulp of the IEEE quotient (correctly rounded in all but
// rare cases, exact whenever the quotient is representable, e.g. the half-integers torch.round ties on).
__device__ __forceinline__ void quant_dequant(float (&d)[8], const float (&t)[8], bool differentiable) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        float y;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t[u]));
        const float q0 = d[u] * y;
        const float x = fmaf(fmaf(-q0, t[u], d[u]), y, q0);
        float q = rintf(x);  // torch.round
        if (differentiable) {
            const float e = __fsub_rn(x, q);
            q = __fadd_rn(q, __fmul_rn(__fmul_rn(e, e), e));  // :40-42
        }
        d[u] = __fmul_rn(q, t[u]);  // :300-306
    }
}

// diffjpeg.py:57-61 evaluated on fp32 0-d tensors
__device__ __forceinline__ float quality_to_factor_dev(float v) {
    const float f = v < 50.0f ? __fdiv_rn(5000.0f, v) : __fsub_rn(200.0f, __fmul_rn(v, 2.0f));
    return __fdiv_rn(f, 100.0f);
}

// Optional fused tail (the chain's last launch): instead of the full image the kernel stores only the LQ crop window
// (traiNNer/data/transforms.py:133-135 + .contiguous(), realesrgan_model.py:627) into a dense (B,3,p,p) tensor, and the CTAs
// behind the codec's copy the GT crop window (transforms.py:124) — clamp/round, both crops and the JPEG in ONE launch.
struct CropTail {
    float* lq_out;           // nullptr: no fused crop, `out` receives the full image
    const float* gt;
    float* gt_out;
    const int32_t* tl_dev;   // device (top, left) override or nullptr
    int top, left, p, scale, Hg, Wg, planes, jpeg_ctas, vec_gt;
};

// One warp per 16x16 MCU, its six 8x8 blocks (Y00 Y01 Y10 Y11 Cb Cr) resident in the warp's shared memory (pitch 12 words,
// blocks 104 apart: a row is two 128-bit accesses, the column pass meets 32 distinct banks), the three passes of the
// separable transform as 48 (block, row) / (block, column) TASKS spread over the lanes:
//   1  lane = (row of the MCU, left / right 8 columns): 16-byte loads, x255, RGB -> YCbCr, 2x2 chroma mean (one shuffle
//      with the lane of the row below), rows into the blocks;
//   2  (block, row):    dct8 along the row;
//   3  (block, column): dct8 down the column, quantise / dequantise with fl(table * factor), idct8 — the eight
//      coefficients of a column never leave the lane;
//   4  (block, row):    idct8 along the row, level shift restored;
//   5  lane = its 8 pixels again: chroma nearest x2, YCbCr -> RGB, clamp, /255, optional 8-bit lattice, 16-byte stores
//      (or the LQ crop window only, see CropTail).
// Round 1-2 kept a lane's pixels, two MCUs' luma rows and the table columns in registers and transposed through shared
// memory (96 registers, 20 warps per SM, 35 % issue utilisation at 64 x 3 x 192^2: every warp one long serial chain);
// this layout needs 48 registers (40 warps per SM), runs at 56 % and 0.0169 ms instead of 0.0216 ms — the structure of the
// integer codec of the fork's JPEG round (libjpeg.cu), which reached 71 %.  Every value goes through the same
// operations in the same order as before: results are bit-identical to the round-2 kernel.
constexpr int kDP = 12, kDBlk = 8 * kDP + 8;
__device__ __forceinline__ void row_load8(const float* p, float (&d)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ void row_store8(float* p, const float (&d)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(d[4], d[5], d[6], d[7]);
}
#ifndef OTF_JPEG_MINB
#define OTF_JPEG_MINB 10
#endif
// NM = MCUs per warp (horizontally adjacent).  Shipped with NM = 1.  NM = 2 makes a pass 96 tasks = three FULL rounds of 32
// lanes (a single MCU's 48 tasks take two rounds, the second half empty: -25 % instructions in the three passes) but
// measured SLOWER at 64 x 3 x 192^2 — 0.0184 ms (64 registers) / 0.0235 ms (48 registers, spills) against 0.0169 ms: half
// as many, twice as long warps hide less latency than the idle lanes cost.
template <bool DIFF, int NM>
__global__ void __launch_bounds__(128, OTF_JPEG_MINB) diffjpeg_kernel(const float* __restrict__ img, float* __restrict__ out, int B, int H,
                                                       int W, int mcu_x, int mcu_y, const float* __restrict__ factor_dev,
                                                       float factor_scalar, int clamp_in,
                                                       int round8_out, int vec_ok, int factor_is_quality,
                                                       const __grid_constant__ CropTail ct) {
    pdl_enter();
    constexpr int kWarpFloats = NM * 6 * kDBlk + 128;  // blocks Y00 Y01 Y10 Y11 Cb Cr of each MCU + fl(table * factor) [2][8][8]
    constexpr int kTasks = NM * 48, kRounds = (kTasks + 31) / 32;
    extern __shared__ __align__(16) float s_mcu[];  // [warps of the CTA][kWarpFloats]
    int top = ct.top, left = ct.left;
    if (ct.lq_out && ct.tl_dev) {  // per-step offsets of a captured chain, clamped so that a bad upload cannot leave the image
        top = clampi(ct.tl_dev[0], 0, H - ct.p);
        left = clampi(ct.tl_dev[1], 0, W - ct.p);
    }
    if (ct.lq_out && (int)blockIdx.x >= ct.jpeg_ctas) {  // the GT crop rides in the same launch (no such CTAs when gt_out is NULL)
        const int64_t q0 = (int64_t)(blockIdx.x - ct.jpeg_ctas) * blockDim.x + threadIdx.x;
        const int64_t qs = (int64_t)(gridDim.x - ct.jpeg_ctas) * blockDim.x;
        if (ct.vec_gt) copy_window<true>(ct.gt, ct.Hg, ct.Wg, top * ct.scale, left * ct.scale, ct.p * ct.scale, ct.gt_out, ct.planes, q0, qs);
        else copy_window<false>(ct.gt, ct.Hg, ct.Wg, top * ct.scale, left * ct.scale, ct.p * ct.scale, ct.gt_out, ct.planes, q0, qs);
        return;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wid;
    const int groups_x = (mcu_x + NM - 1) / NM, per_img = groups_x * mcu_y;
    if (warp >= (int64_t)B * per_img) return;  // warp-uniform
    const int b = (int)(warp / per_img), m = (int)(warp - (int64_t)b * per_img);
    const int my = m / groups_x, mx0 = NM * (m - my * groups_x);
    float* blk = s_mcu + wid * kWarpFloats;
    float* tab = blk + NM * 6 * kDBlk;
    float factor = factor_dev ? factor_dev[b] : factor_scalar;
    if (factor_is_quality) factor = quality_to_factor_dev(factor);  // diffjpeg.py:57-61 fused (no extra launch)
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // fl(table[u][v] * factor), both tables (diffjpeg.py:207-212 divides by exactly this product)
        const int idx = lane + 32 * j;
        const float* t = idx < 64 ? &c_ytab[0][0] + idx : &c_ctab[0][0] + (idx - 64);
        tab[idx] = __fmul_rn(__ldg(t), factor);
    }
    const size_t hw = (size_t)H * W;
    const int ly = lane >> 1, half = lane & 1;  // row of the MCU, which 8 of its 16 columns
    const int y = my * 16 + ly;
    const bool row_ok = y < H;
    const bool rows_in = vec_ok && (my * 16 + 16 <= H);
    // ---- phase 1: 8 px x 3 channels per lane and MCU (zero padding outside the image: diffjpeg.py:515-522), x255,
    //      RGB -> YCbCr (:70-91), chroma 2x2 mean (:112-125; the +128 shift and the -128 level shift cancel) ----
#pragma unroll
    for (int p = 0; p < NM; ++p) {
        const int x0 = (mx0 + p) * 16 + half * 8;
        const bool interior = rows_in && ((mx0 + p) * 16 + 16 <= W);  // warp-uniform
        float* bp = blk + p * 6 * kDBlk;
        float px[3][8];
        const float* ip0 = img + (size_t)b * 3 * hw + (size_t)y * W + x0;
        if (interior) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(ip0 + c * hw)), d = __ldg(reinterpret_cast<const float4*>(ip0 + c * hw) + 1);
                px[c][0] = a.x; px[c][1] = a.y; px[c][2] = a.z; px[c][3] = a.w; px[c][4] = d.x; px[c][5] = d.y; px[c][6] = d.z; px[c][7] = d.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int k = 0; k < 8; ++k) px[c][k] = (row_ok && x0 + k < W) ? __ldg(ip0 + c * hw + k) : 0.0f;
        }
        float yrow[8], cb[8], cr[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float R = px[0][k], G = px[1][k], Bc = px[2][k];
            if (clamp_in) { R = clamp01(R); G = clamp01(G); Bc = clamp01(Bc); }
            R = __fmul_rn(R, 255.0f); G = __fmul_rn(G, 255.0f); Bc = __fmul_rn(Bc, 255.0f);
            yrow[k] = fmaf(Bc, 0.114f, fmaf(G, 0.587f, R * 0.299f)) - 128.0f;
            cb[k] = fmaf(Bc, 0.5f, fmaf(G, -0.331264f, R * -0.168736f));
            cr[k] = fmaf(Bc, -0.081312f, fmaf(G, -0.418688f, R * 0.5f));
        }
        row_store8(bp + ((ly >> 3) * 2 + half) * kDBlk + (ly & 7) * kDP, yrow);
        float cbs[4], crs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // the row below / above sits in lane ^ 2
            float s0 = cb[2 * k] + cb[2 * k + 1], s1 = cr[2 * k] + cr[2 * k + 1];
            s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            cbs[k] = s0 * 0.25f;
            crs[k] = s1 * 0.25f;
        }
        if (!(ly & 1)) {  // chroma row ly / 2, columns half * 4 .. + 3
            *reinterpret_cast<float4*>(bp + 4 * kDBlk + (ly >> 1) * kDP + half * 4) = make_float4(cbs[0], cbs[1], cbs[2], cbs[3]);
            *reinterpret_cast<float4*>(bp + 5 * kDBlk + (ly >> 1) * kDP + half * 4) = make_float4(crs[0], crs[1], crs[2], crs[3]);
        }
    }
    __syncwarp();
    // ---- phase 2: 1-D DCT along the rows (column index -> v), (block, row) tasks ----
#pragma unroll
    for (int rnd = 0; rnd < kRounds; ++rnd) {
        const int task = lane + 32 * rnd;
        if (kTasks % 32 == 0 || task < kTasks) {
            float* p = blk + (task >> 3) * kDBlk + (task & 7) * kDP;
            float f[8];
            row_load8(p, f);
            dct8(f);
            row_store8(p, f);
        }
    }
    __syncwarp();
    // ---- phase 3: per (block, column v): DCT along the rows' index (-> u), quantise, dequantise, IDCT (u -> row) ----
#pragma unroll
    for (int rnd = 0; rnd < kRounds; ++rnd) {
        const int task = lane + 32 * rnd;
        if (kTasks % 32 == 0 || task < kTasks) {
            const int bi = task >> 3, c = task & 7;
            float* p = blk + bi * kDBlk + c;
            const float* tp = tab + ((bi % 6) >= 4 ? 64 : 0) + c;
            float f[8], t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { f[k] = p[k * kDP]; t[k] = tp[8 * k]; }
            dct8(f);
            quant_dequant(f, t, DIFF);
            idct8(f);
#pragma unroll
            for (int k = 0; k < 8; ++k) p[k * kDP] = f[k];
        }
    }
    __syncwarp();
    // ---- phase 4: 1-D IDCT along the rows (v -> column), level shift restored ----
#pragma unroll
    for (int rnd = 0; rnd < kRounds; ++rnd) {
        const int task = lane + 32 * rnd;
        if (kTasks % 32 == 0 || task < kTasks) {
            float* p = blk + (task >> 3) * kDBlk + (task & 7) * kDP;
            float f[8];
            row_load8(p, f);
            idct8(f);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += 128.0f;
            row_store8(p, f);
        }
    }
    __syncwarp();
    // ---- phase 5: chroma nearest x2 (diffjpeg.py:397-402), YCbCr -> RGB (:415-431), clamp, /255 (:476-479), optional 8-bit
    //      lattice; the lane's own 8 pixels of each MCU again ----
    if (!row_ok) return;
#pragma unroll
    for (int p = 0; p < NM; ++p) {
        const int x0 = (mx0 + p) * 16 + half * 8;
        if (x0 >= W) break;
        const bool interior = rows_in && ((mx0 + p) * 16 + 16 <= W);
        const float* bp = blk + p * 6 * kDBlk;
        float yrow[8];
        row_load8(bp + ((ly >> 3) * 2 + half) * kDBlk + (ly & 7) * kDP, yrow);
        const float4 vb = *reinterpret_cast<const float4*>(bp + 4 * kDBlk + (ly >> 1) * kDP + half * 4);
        const float4 vr = *reinterpret_cast<const float4*>(bp + 5 * kDBlk + (ly >> 1) * kDP + half * 4);
        const float cbv[4] = {vb.x, vb.y, vb.z, vb.w}, crv[4] = {vr.x, vr.y, vr.z, vr.w};
        float res[3][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float Y = yrow[k], Cb = cbv[k >> 1] - 128.0f, Cr = crv[k >> 1] - 128.0f;   // (+128 was restored on every plane)
            float R = fmaf(Cr, 1.402f, Y);
            float G = fmaf(Cr, -0.714136f, fmaf(Cb, -0.344136f, Y));
            float Bc = fmaf(Cb, 1.772f, Y);
            R = div255(fminf(fmaxf(R, 0.0f), 255.0f));
            G = div255(fminf(fmaxf(G, 0.0f), 255.0f));
            Bc = div255(fminf(fmaxf(Bc, 0.0f), 255.0f));
            if (round8_out) { R = quantise8_unit(R); G = quantise8_unit(G); Bc = quantise8_unit(Bc); }
            res[0][k] = R; res[1][k] = G; res[2][k] = Bc;
        }
        if (ct.lq_out) {  // only the crop window, into the dense (B,3,p,p) output
            const int yy = y - top;
            if (yy < 0 || yy >= ct.p) continue;
            float* lp = ct.lq_out + ((size_t)b * 3 * ct.p + yy) * ct.p;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int xx = x0 + k - left;
                    if (xx >= 0 && xx < ct.p && x0 + k < W) lp[(size_t)c * ct.p * ct.p + xx] = res[c][k];
                }
            continue;
        }
        float* op = out + (size_t)b * 3 * hw + (size_t)y * W + x0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (interior) {
                reinterpret_cast<float4*>(op + c * hw)[0] = make_float4(res[c][0], res[c][1], res[c][2], res[c][3]);
                reinterpret_cast<float4*>(op + c * hw)[1] = make_float4(res[c][4], res[c][5], res[c][6], res[c][7]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (x0 + k < W) op[c * hw + k] = res[c][k];
            }
        }
    }
}

// Host side of the launch: MCUs per warp, warps per CTA, grid, dynamic shared memory.
struct JpegLaunch {
    int nm, wpc, ctas;
    size_t smem;
};
static JpegLaunch jpeg_launch(int B, int mcu_x, int mcu_y, int force_wpc) {
    JpegLaunch l;
    l.nm = 1;
    const int64_t warps = (int64_t)B * ((mcu_x + l.nm - 1) / l.nm) * mcu_y;
    l.wpc = force_wpc ? force_wpc : (warps >= (int64_t)kNumSMs * 16 ? 4 : 1);  // few MCUs -> one-warp CTAs so they spread over all 148 SMs
    l.ctas = (int)ceil_div(warps, l.wpc);
    l.smem = (size_t)l.wpc * (l.nm * 6 * kDBlk + 128) * sizeof(float);
    return l;
}
template <typename... Args>
static cudaError_t launch_jpeg(const JpegLaunch& l, int differentiable, int extra_ctas, cudaStream_t st, Args... args) {
    const dim3 grid(l.ctas + extra_ctas), block(32 * l.wpc);
    return differentiable ? launch_chain(diffjpeg_kernel<true, 1>, grid, block, l.smem, st, args...)
                          : launch_chain(diffjpeg_kernel<false, 1>, grid, block, l.smem, st, args...);
}

__global__ void quality_to_factor_kernel(float* q, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) q[i] = quality_to_factor_dev(q[i]);
}

}  // namespace otf

extern "C" int otf_quality_to_factor_f32(float* quality_dev, int B, void* stream) {
    using namespace otf;
    OTF_REQUIRE(quality_dev && B > 0, OTF_ERR_BAD_ARG, "quality_to_factor: bad args");
    quality_to_factor_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(quality_dev, B);
    OTF_LAUNCH_CHECK("quality_to_factor_kernel");
    return OTF_OK;
}

extern "C" int otf_diffjpeg_f32(const float* img, int B, int H, int W, const float* factor_dev, float factor_scalar,
                                int factor_is_quality, int differentiable, int clamp_in, int round8_out, float* out,
                                void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out, OTF_ERR_BAD_ARG, "diffjpeg: null pointer");
    OTF_REQUIRE(B > 0 && H > 0 && W > 0, OTF_ERR_BAD_ARG, "diffjpeg: bad extents");
    const int mcu_x = ceil_div(W, 16), mcu_y = ceil_div(H, 16);
    const int vec_ok = (W % 4 == 0) && (((uintptr_t)img & 15) == 0) && (((uintptr_t)out & 15) == 0);
    CropTail ct;
    memset(&ct, 0, sizeof(ct));
    const JpegLaunch l = jpeg_launch(B, mcu_x, mcu_y, 0);
    ct.jpeg_ctas = l.ctas;
    launch_jpeg(l, differentiable, 0, (cudaStream_t)stream, img, out, B, H, W, mcu_x, mcu_y, factor_dev, factor_scalar, clamp_in, round8_out,
                   vec_ok, factor_is_quality, ct);
    OTF_LAUNCH_CHECK("diffjpeg_kernel");
    return OTF_OK;
}

extern "C" int otf_diffjpeg_crop_pair_f32(const float* img, int B, int H, int W, const float* factor_dev, float factor_scalar,
                                          int factor_is_quality, int differentiable, int clamp_in, const float* gt, int Hg, int Wg,
                                          int top, int left, const int32_t* top_left_dev, int lq_patch, int scale, float* gt_out,
                                          float* lq_out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && lq_out && (gt || !gt_out), OTF_ERR_BAD_ARG, "diffjpeg_crop_pair: null pointer");
    OTF_REQUIRE(B > 0 && B * 3 <= 65535 && H > 0 && W > 0 && scale > 0 && lq_patch > 0, OTF_ERR_BAD_ARG, "diffjpeg_crop_pair: bad extents");
    OTF_REQUIRE(Hg == H * scale && Wg == W * scale, OTF_ERR_BAD_ARG, "diffjpeg_crop_pair: GT (%d, %d) is not %dx LQ (%d, %d)", Hg, Wg, scale, H, W);
    OTF_REQUIRE(top >= 0 && left >= 0 && top + lq_patch <= H && left + lq_patch <= W, OTF_ERR_BAD_ARG, "diffjpeg_crop_pair: window outside LQ");
    OTF_REQUIRE(!gt_out || ((lq_patch * scale) % 4 == 0 && (((uintptr_t)gt_out) & 15) == 0), OTF_ERR_UNSUPPORTED,
                "diffjpeg_crop_pair: GT patch must be a multiple of 4 pixels wide (use otf_diffjpeg_f32 + otf_crop_pair_f32)");
    const int mcu_x = ceil_div(W, 16), mcu_y = ceil_div(H, 16);
    const JpegLaunch l2 = jpeg_launch(B, mcu_x, mcu_y, 4);
    const int64_t warps = (int64_t)l2.ctas * 4;
    const int vec_ok = (W % 4 == 0) && (((uintptr_t)img & 15) == 0);
    const int wpc = 4;  // (the GT copy behind the codec wants full CTAs)
    CropTail ct;
    memset(&ct, 0, sizeof(ct));
    ct.lq_out = lq_out; ct.gt = gt; ct.gt_out = gt_out; ct.tl_dev = top_left_dev;
    ct.top = top; ct.left = left; ct.p = lq_patch; ct.scale = scale; ct.Hg = Hg; ct.Wg = Wg; ct.planes = B * 3;
    ct.jpeg_ctas = ceil_div(warps, wpc);
    // device-side offsets: the alignment of the GT window start is only known when scale % 4 == 0
    ct.vec_gt = (Wg % 4 == 0) && (top_left_dev ? scale % 4 == 0 : (left * scale) % 4 == 0) && (((uintptr_t)gt & 15) == 0);
    const int64_t gquads = (int64_t)B * 3 * lq_patch * scale * (lq_patch * scale / 4);
    int copy_ctas = (int)((gquads / 4 + 127) / 128);
    if (copy_ctas > kNumSMs * 8) copy_ctas = kNumSMs * 8;
    if (copy_ctas < 1) copy_ctas = 1;
    if (!gt_out) copy_ctas = 0;  // the GT window stays a view of the caller's tensor (what the reference's crop returns)
    launch_jpeg(l2, differentiable, copy_ctas, (cudaStream_t)stream, img, (float*)nullptr, B, H, W, mcu_x, mcu_y, factor_dev, factor_scalar,
                   clamp_in, 1, vec_ok, factor_is_quality, ct);
    OTF_LAUNCH_CHECK("diffjpeg_kernel (fused crop)");
    return OTF_OK;
}

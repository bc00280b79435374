// Fused DiffJPEG: one kernel for the whole of traiNNer/utils/diffjpeg.py:503-527
// (x255 -> RGB2YCbCr -> 4:2:0 -> 8x8 DCT -> quantise by table*factor[b] -> round -> dequantise ->
//  IDCT -> chroma x2 -> YCbCr2RGB -> clamp -> /255 -> crop), which the reference runs as ~60 ATen
// launches and ~25 image-sized HBM round trips.  Here the image is read once and written once.
//
// Mapping: one warp = one 16x16 MCU (4 Y blocks + Cb + Cr).  Lane l owns the 8-pixel row segment
// (block yb = l>>3, row r = l&7) of the MCU, i.e. row (yb>>1)*8+r, columns (yb&1)*8..+7 — 32 B per
// channel per lane, float4-vectorised.  The 8x8 DCT is register resident: an 8-point transform
// along the lane's own row, an 8x8 transpose across the 8 lanes of the block (12 shuffles), a second
// 8-point transform; chroma rows are gathered onto lanes 0-7 (Cb) / 8-15 (Cr) with shuffles.
// Arithmetic order (documented for parity): 1-D DCT-II with the orthonormal matrix
// C[k][n] = 0.5*alpha_k*cos((2n+1)k*pi/16), sums taken n = 0..7 with FFMA; quantisation divides by
// fl(table*factor) exactly as the reference (no reciprocal), torch.round = rintf (half to even).
#include <string.h>

#include "otf_common.cuh"

namespace otf {

// C[k][n] = 0.5*alpha_k*cos((2n+1)k*pi/16) rounded to fp32 (orthonormal DCT-II; the reference's
// 0.25*alpha_u*alpha_v*cos*cos of diffjpeg.py:155-164 is its outer product)
__constant__ float c_dct[8][8] = {
    {3.535533845e-01f, 3.535533845e-01f, 3.535533845e-01f, 3.535533845e-01f, 3.535533845e-01f, 3.535533845e-01f, 3.535533845e-01f, 3.535533845e-01f},
    {4.903926253e-01f, 4.157347977e-01f, 2.777851224e-01f, 9.754516184e-02f, -9.754516184e-02f, -2.777851224e-01f, -4.157347977e-01f, -4.903926253e-01f},
    {4.619397521e-01f, 1.913417131e-01f, -1.913417131e-01f, -4.619397521e-01f, -4.619397521e-01f, -1.913417131e-01f, 1.913417131e-01f, 4.619397521e-01f},
    {4.157347977e-01f, -9.754516184e-02f, -4.903926253e-01f, -2.777851224e-01f, 2.777851224e-01f, 4.903926253e-01f, 9.754516184e-02f, -4.157347977e-01f},
    {3.535533845e-01f, -3.535533845e-01f, -3.535533845e-01f, 3.535533845e-01f, 3.535533845e-01f, -3.535533845e-01f, -3.535533845e-01f, 3.535533845e-01f},
    {2.777851224e-01f, -4.903926253e-01f, 9.754516184e-02f, 4.157347977e-01f, -4.157347977e-01f, -9.754516184e-02f, 4.903926253e-01f, -2.777851224e-01f},
    {1.913417131e-01f, -4.619397521e-01f, 4.619397521e-01f, -1.913417131e-01f, -1.913417131e-01f, 4.619397521e-01f, -4.619397521e-01f, 1.913417131e-01f},
    {9.754516184e-02f, -2.777851224e-01f, 4.157347977e-01f, -4.903926253e-01f, 4.903926253e-01f, -4.157347977e-01f, 2.777851224e-01f, -9.754516184e-02f},
};
// Quantisation tables indexed [u][v] exactly as the reference stores them: the Annex-K luminance
// table TRANSPOSED (diffjpeg.py:18-31) and the chroma table (:32-37, symmetric).
__constant__ float c_ytab[8][8] = {
    {16, 12, 14, 14, 18, 24, 49, 72},     {11, 12, 13, 17, 22, 35, 64, 92},   {10, 14, 16, 22, 37, 55, 78, 95},
    {16, 19, 24, 29, 56, 64, 87, 98},     {24, 26, 40, 51, 68, 81, 103, 112}, {40, 58, 57, 87, 109, 104, 121, 100},
    {51, 60, 69, 80, 103, 113, 120, 103}, {61, 55, 56, 62, 77, 92, 101, 99},
};
__constant__ float c_ctab[8][8] = {
    {17, 18, 24, 47, 99, 99, 99, 99}, {18, 21, 26, 66, 99, 99, 99, 99}, {24, 26, 56, 99, 99, 99, 99, 99},
    {47, 66, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99},
    {99, 99, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99},
};

// y[k] = sum_n C[k][n] x[n]   (forward)      x[n] = sum_k C[k][n] y[k]   (inverse)
__device__ __forceinline__ void dct8(float (&x)[8]) {
    float y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float s = c_dct[k][0] * x[0];
#pragma unroll
        for (int n = 1; n < 8; ++n) s = fmaf(c_dct[k][n], x[n], s);
        y[k] = s;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = y[k];
}
__device__ __forceinline__ void idct8(float (&y)[8]) {
    float x[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        float s = c_dct[0][n] * y[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) s = fmaf(c_dct[k][n], y[k], s);
        x[n] = s;
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) y[n] = x[n];
}

// 8x8 transpose across the 8 lanes of a group: lane g holds row g in a[0..7]; afterwards lane g
// holds column g.
__device__ __forceinline__ void transpose8(float (&a)[8], int lane) {
#pragma unroll
    for (int s = 1; s < 8; s <<= 1) {
        const bool up = lane & s;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i & s) continue;
            const float send = up ? a[i] : a[i | s];
            const float recv = __shfl_xor_sync(0xffffffffu, send, s);
            if (up) a[i] = recv; else a[i | s] = recv;
        }
    }
}

// Quantise + dequantise an 8x8 block held as: lane v (within its group of 8) owns coefficients
// D[u][v], u = register index.  tab is [u][v]: luma or chroma table, selected at run time so that the
// codec below is ONE loop body for both planes (half the code: the fully unrolled kernel overflowed
// the instruction cache — "no_instruction" was its second largest stall in profiles/).
__device__ __forceinline__ void quant_dequant(float (&d)[8], int v, bool luma, float factor, bool differentiable) {
    const float(*tab)[8] = luma ? c_ytab : c_ctab;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float t = __fmul_rn(tab[u][v], factor);  // table * factor
        const float x = __fdiv_rn(d[u], t);            // diffjpeg.py:207-212
        float q = rintf(x);                            // torch.round
        if (differentiable) {
            const float e = __fsub_rn(x, q);
            q = __fadd_rn(q, __fmul_rn(__fmul_rn(e, e), e));  // :40-42
        }
        d[u] = __fmul_rn(q, t);  // :300-306
    }
}

// Full 2-D DCT -> quantise -> dequantise -> 2-D IDCT on the block whose row `lane&7` is in f[].
// Input is level-shifted (f-128); output has +128 restored.
__device__ __forceinline__ void block_codec(float (&f)[8], int lane, bool luma, float factor, bool differentiable) {
    dct8(f);             // along the row (y index -> v)
    transpose8(f, lane); // lane now = v, registers = x (row index)
    dct8(f);             // along x -> u
    quant_dequant(f, lane & 7, luma, factor, differentiable);
    idct8(f);            // u -> x
    transpose8(f, lane); // lane = x, registers = v
    idct8(f);            // v -> y
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] += 128.0f;
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

__global__ void __launch_bounds__(128) diffjpeg_kernel(const float* __restrict__ img, float* __restrict__ out, int B, int H,
                                                       int W, int mcu_x, int mcu_y, const float* __restrict__ factor_dev,
                                                       float factor_scalar, int differentiable, int clamp_in,
                                                       int round8_out, int vec_ok, int factor_is_quality,
                                                       const __grid_constant__ CropTail ct) {
    int top = ct.top, left = ct.left;
    if (ct.lq_out) {
        if (ct.tl_dev) {  // per-step offsets of a captured chain, clamped so that a bad upload cannot leave the image
            top = clampi(ct.tl_dev[0], 0, H - ct.p);
            left = clampi(ct.tl_dev[1], 0, W - ct.p);
        }
        if ((int)blockIdx.x >= ct.jpeg_ctas) {  // the GT crop rides in the same launch
            const int64_t q0 = (int64_t)(blockIdx.x - ct.jpeg_ctas) * blockDim.x + threadIdx.x;
            const int64_t qs = (int64_t)(gridDim.x - ct.jpeg_ctas) * blockDim.x;
            if (ct.vec_gt) copy_window<true>(ct.gt, ct.Hg, ct.Wg, top * ct.scale, left * ct.scale, ct.p * ct.scale, ct.gt_out, ct.planes, q0, qs);
            else copy_window<false>(ct.gt, ct.Hg, ct.Wg, top * ct.scale, left * ct.scale, ct.p * ct.scale, ct.gt_out, ct.planes, q0, qs);
            return;
        }
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total = (int64_t)B * mcu_x * mcu_y;
    if (warp >= total) return;  // warp-uniform
    const int b = (int)(warp / ((int64_t)mcu_x * mcu_y));
    const int m = (int)(warp - (int64_t)b * mcu_x * mcu_y);
    const int my = m / mcu_x, mx = m - my * mcu_x;
    float factor = factor_dev ? factor_dev[b] : factor_scalar;
    if (factor_is_quality) factor = quality_to_factor_dev(factor);  // diffjpeg.py:57-61 fused (no extra launch)

    const int yb = lane >> 3, r = lane & 7;
    const int by = yb >> 1, bx = yb & 1;
    const int y = my * 16 + by * 8 + r;
    const int x0 = mx * 16 + bx * 8;
    const size_t hw = (size_t)H * W;
    const float* ip = img + (size_t)b * 3 * hw + (size_t)y * W + x0;
    const bool row_ok = y < H;
    const bool full = row_ok && (x0 + 8 <= W);

    // ---- load 8 px x 3 channels (zero padding outside the image: diffjpeg.py:515-522) ----
    float px[3][8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (full && vec_ok) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(ip + c * hw));
            const float4 d = __ldg(reinterpret_cast<const float4*>(ip + c * hw) + 1);
            px[c][0] = a.x; px[c][1] = a.y; px[c][2] = a.z; px[c][3] = a.w;
            px[c][4] = d.x; px[c][5] = d.y; px[c][6] = d.z; px[c][7] = d.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) px[c][k] = (row_ok && x0 + k < W) ? __ldg(ip + c * hw + k) : 0.0f;
        }
    }
    // ---- x255, RGB -> YCbCr (diffjpeg.py:70-91), level shift for Y folded in ----
    float yv[8], cb[8], cr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float R = px[0][k], G = px[1][k], Bc = px[2][k];
        if (clamp_in) { R = clamp01(R); G = clamp01(G); Bc = clamp01(Bc); }
        R = __fmul_rn(R, 255.0f); G = __fmul_rn(G, 255.0f); Bc = __fmul_rn(Bc, 255.0f);
        yv[k] = fmaf(Bc, 0.114f, fmaf(G, 0.587f, R * 0.299f)) - 128.0f;
        cb[k] = fmaf(Bc, 0.5f, fmaf(G, -0.331264f, R * -0.168736f));  // the +128 shift is folded away, see below
        cr[k] = fmaf(Bc, -0.081312f, fmaf(G, -0.418688f, R * 0.5f));
    }
    // ---- chroma 2x2 mean (diffjpeg.py:112-125); the +128 shift and the -128 level shift cancel ----
    float cbs[4], crs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float s0 = cb[2 * k] + cb[2 * k + 1], s1 = cr[2 * k] + cr[2 * k + 1];
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        cbs[k] = s0 * 0.25f;
        crs[k] = s1 * 0.25f;
    }
    // gather chroma rows: target lane t (&15): comp = t>>3 (0 Cb, 1 Cr), chroma row mrow = t&7
    //   source block row by' = mrow>>2, source lane row r' = (mrow&3)*2, left half from bx'=0, right from bx'=1
    float ch[8];
    {
        const int t = lane & 15, mrow = t & 7;
        const int src_l = (((mrow >> 2) * 2 + 0) << 3) + (mrow & 3) * 2;
        const int src_r = (((mrow >> 2) * 2 + 1) << 3) + (mrow & 3) * 2;
        const bool is_cr = t >> 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float lb = __shfl_sync(0xffffffffu, cbs[k], src_l), lr = __shfl_sync(0xffffffffu, crs[k], src_l);
            const float rb = __shfl_sync(0xffffffffu, cbs[k], src_r), rr = __shfl_sync(0xffffffffu, crs[k], src_r);
            ch[k] = is_cr ? lr : lb;
            ch[4 + k] = is_cr ? rr : rb;
        }
    }
    // ---- codec: one loop body, pass 0 = the four luma blocks, pass 1 = Cb | Cr (lanes 16-31 redo 0-15) ----
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        float blk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) blk[k] = pass ? ch[k] : yv[k];
        block_codec(blk, lane, pass == 0, factor, differentiable);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (pass) ch[k] = blk[k]; else yv[k] = blk[k];
        }
    }
    // ---- chroma back to pixel lanes: nearest x2 (diffjpeg.py:397-402) ----
    {
        const int mrow = by * 4 + (r >> 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float b_lo = __shfl_sync(0xffffffffu, ch[k], mrow), b_hi = __shfl_sync(0xffffffffu, ch[4 + k], mrow);
            const float r_lo = __shfl_sync(0xffffffffu, ch[k], 8 + mrow), r_hi = __shfl_sync(0xffffffffu, ch[4 + k], 8 + mrow);
            const float vb = bx ? b_hi : b_lo, vr = bx ? r_hi : r_lo;
            cb[2 * k] = vb; cb[2 * k + 1] = vb;
            cr[2 * k] = vr; cr[2 * k + 1] = vr;
        }
    }
    // ---- YCbCr -> RGB (diffjpeg.py:415-431), clamp, /255 (:476-479), optional 8-bit lattice ----
    float* op = out + (size_t)b * 3 * hw + (size_t)y * W + x0;
    float res[3][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float Y = yv[k], Cb = cb[k] - 128.0f, Cr = cr[k] - 128.0f;   // block_codec restored +128 on every plane
        float R = fmaf(Cr, 1.402f, Y);
        float G = fmaf(Cr, -0.714136f, fmaf(Cb, -0.344136f, Y));
        float Bc = fmaf(Cb, 1.772f, Y);
        R = __fdiv_rn(fminf(fmaxf(R, 0.0f), 255.0f), 255.0f);
        G = __fdiv_rn(fminf(fmaxf(G, 0.0f), 255.0f), 255.0f);
        Bc = __fdiv_rn(fminf(fmaxf(Bc, 0.0f), 255.0f), 255.0f);
        if (round8_out) { R = quantise8(R); G = quantise8(G); Bc = quantise8(Bc); }
        res[0][k] = R; res[1][k] = G; res[2][k] = Bc;
    }
    if (!row_ok) return;
    if (ct.lq_out) {  // only the crop window, into the dense (B,3,p,p) output
        const int yy = y - top;
        if (yy < 0 || yy >= ct.p) return;
        float* lp = ct.lq_out + ((size_t)b * 3 * ct.p + yy) * ct.p;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int xx = x0 + k - left;
                if (xx >= 0 && xx < ct.p && x0 + k < W) lp[(size_t)c * ct.p * ct.p + xx] = res[c][k];
            }
        return;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (full && vec_ok) {
            reinterpret_cast<float4*>(op + c * hw)[0] = make_float4(res[c][0], res[c][1], res[c][2], res[c][3]);
            reinterpret_cast<float4*>(op + c * hw)[1] = make_float4(res[c][4], res[c][5], res[c][6], res[c][7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (x0 + k < W) op[c * hw + k] = res[c][k];
        }
    }
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
    const int64_t warps = (int64_t)B * mcu_x * mcu_y;
    const int vec_ok = (W % 4 == 0) && (((uintptr_t)img & 15) == 0) && (((uintptr_t)out & 15) == 0);
    // one warp per MCU; few MCUs (64^2 LQ stage) -> one-warp CTAs so they spread over all 148 SMs
    const int wpc = warps >= (int64_t)kNumSMs * 32 ? 4 : 1;
    CropTail ct;
    memset(&ct, 0, sizeof(ct));
    diffjpeg_kernel<<<ceil_div(warps, wpc), 32 * wpc, 0, (cudaStream_t)stream>>>(img, out, B, H, W, mcu_x, mcu_y, factor_dev,
                                                                             factor_scalar, differentiable, clamp_in,
                                                                             round8_out, vec_ok, factor_is_quality, ct);
    OTF_LAUNCH_CHECK("diffjpeg_kernel");
    return OTF_OK;
}

extern "C" int otf_diffjpeg_crop_pair_f32(const float* img, int B, int H, int W, const float* factor_dev, float factor_scalar,
                                          int factor_is_quality, int differentiable, int clamp_in, const float* gt, int Hg, int Wg,
                                          int top, int left, const int32_t* top_left_dev, int lq_patch, int scale, float* gt_out,
                                          float* lq_out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && gt && gt_out && lq_out, OTF_ERR_BAD_ARG, "diffjpeg_crop_pair: null pointer");
    OTF_REQUIRE(B > 0 && B * 3 <= 65535 && H > 0 && W > 0 && scale > 0 && lq_patch > 0, OTF_ERR_BAD_ARG, "diffjpeg_crop_pair: bad extents");
    OTF_REQUIRE(Hg == H * scale && Wg == W * scale, OTF_ERR_BAD_ARG, "diffjpeg_crop_pair: GT (%d, %d) is not %dx LQ (%d, %d)", Hg, Wg, scale, H, W);
    OTF_REQUIRE(top >= 0 && left >= 0 && top + lq_patch <= H && left + lq_patch <= W, OTF_ERR_BAD_ARG, "diffjpeg_crop_pair: window outside LQ");
    OTF_REQUIRE((lq_patch * scale) % 4 == 0 && (((uintptr_t)gt_out) & 15) == 0, OTF_ERR_UNSUPPORTED,
                "diffjpeg_crop_pair: GT patch must be a multiple of 4 pixels wide (use otf_diffjpeg_f32 + otf_crop_pair_f32)");
    const int mcu_x = ceil_div(W, 16), mcu_y = ceil_div(H, 16);
    const int64_t warps = (int64_t)B * mcu_x * mcu_y;
    const int vec_ok = (W % 4 == 0) && (((uintptr_t)img & 15) == 0);
    const int wpc = 4;
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
    diffjpeg_kernel<<<ct.jpeg_ctas + copy_ctas, 32 * wpc, 0, (cudaStream_t)stream>>>(img, nullptr, B, H, W, mcu_x, mcu_y, factor_dev,
                                                                                  factor_scalar, differentiable, clamp_in, 1, vec_ok,
                                                                                  factor_is_quality, ct);
    OTF_LAUNCH_CHECK("diffjpeg_kernel (fused crop)");
    return OTF_OK;
}

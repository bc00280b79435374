// The JPEG round of the fork's unified compression stage, bit for bit what the reference gets from PIL / libjpeg(-turbo).
//
// Reference: traiNNer/models/paragon_otf_degradations.py:95-158 (`_compress_with_format(..., "jpeg")`): per image a D2H
// copy, `(img.clamp(0,1) * 255).astype(uint8)`, `PIL.Image.save(buffer, format="JPEG", quality=int(q))`,
// `Image.open(buffer).convert("RGB")`, `/ 255`, H2D.  Entropy coding is lossless, so the decoded pixels are fixed by the
// lossy half of baseline JPEG with libjpeg's defaults (4:2:0, Annex-K tables scaled by the quality, the "slow" integer
// DCT, fancy chroma up-sampling) — integer arithmetic that is restated here operation for operation from libjpeg's
// sources (jccolor.c, jcprepct.c, jcsample.c, jfdctint.c, jcdctmgr.c, jcparam.c, jidctint.c, jdsample.c, jdcolor.c; the same
// restatement in numpy is oracle/libjpeg_oracle.py, pinned against PIL itself) and runs on the device in two launches:
//
//   libjpeg_codec_kernel     one warp per 16x16 MCU: uint8 truncation, RGB -> YCbCr, edge expansion, 2x2 chroma
//                            down-sampling, forward DCT, quantise, de-quantise, inverse DCT, range limit for its six
//                            8x8 blocks; decoded Y (full size) and Cb / Cr (half size) go to a byte workspace;
//   libjpeg_upsample_kernel  one thread per chroma sample: triangle-filter up-sampling from the 3x3 chroma neighbourhood
//                            (it crosses MCU borders, hence the second launch), YCbCr -> RGB for the 2x2 pixels, / 255.
//
// Byte / int32 work: ~1.5 workspace bytes per pixel between the launches, no floating point except the first
// multiply by 255 and the last division by 255 (both as the reference does them in float32).
#include "otf_common.cuh"

namespace otf {

struct JpegQuant {
    uint16_t q[2][64];  // luminance, chrominance (natural order), jcparam.c jpeg_set_quality(force_baseline)
};

constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
              F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069, F_2_053119869 = 16819,
              F_2_562915447 = 20995, F_3_072711026 = 25172;
constexpr int JCONST_BITS = 13, JPASS1_BITS = 2;

__device__ __forceinline__ int jdescale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// jfdctint.c: one 1-D pass over eight values.  FIRST: the row pass (results scaled up by 2^PASS1_BITS); else the column
// pass (scaling removed except for the overall factor 8).
template <bool FIRST>
__device__ __forceinline__ void fdct8(int (&d)[8]) {
    const int t0 = d[0] + d[7], t7 = d[0] - d[7], t1 = d[1] + d[6], t6 = d[1] - d[6];
    const int t2 = d[2] + d[5], t5 = d[2] - d[5], t3 = d[3] + d[4], t4 = d[3] - d[4];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int n = FIRST ? JCONST_BITS - JPASS1_BITS : JCONST_BITS + JPASS1_BITS;
    if (FIRST) {
        d[0] = (t10 + t11) << JPASS1_BITS;
        d[4] = (t10 - t11) << JPASS1_BITS;
    } else {
        d[0] = jdescale(t10 + t11, JPASS1_BITS);
        d[4] = jdescale(t10 - t11, JPASS1_BITS);
    }
    int z1 = (t12 + t13) * F_0_541196100;
    d[2] = jdescale(z1 + t13 * F_0_765366865, n);
    d[6] = jdescale(z1 - t12 * F_1_847759065, n);
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * F_1_175875602;
    const int a4 = t4 * F_0_298631336, a5 = t5 * F_2_053119869, a6 = t6 * F_3_072711026, a7 = t7 * F_1_501321110;
    z1 = -z1 * F_0_899976223;
    z2 = -z2 * F_2_562915447;
    z3 = -z3 * F_1_961570560 + z5;
    z4 = -z4 * F_0_390180644 + z5;
    d[7] = jdescale(a4 + z1 + z3, n);
    d[5] = jdescale(a5 + z2 + z4, n);
    d[3] = jdescale(a6 + z2 + z3, n);
    d[1] = jdescale(a7 + z1 + z4, n);
}

// jidctint.c: one 1-D pass.  FIRST: the column pass on de-quantised coefficients; else the row pass with the final
// descale by CONST_BITS + PASS1_BITS + 3.
template <bool FIRST>
__device__ __forceinline__ void idct8(int (&v)[8]) {
    int z1 = (v[2] + v[6]) * F_0_541196100;
    int t2 = z1 - v[6] * F_1_847759065, t3 = z1 + v[2] * F_0_765366865;
    int t0 = (v[0] + v[4]) << JCONST_BITS, t1 = (v[0] - v[4]) << JCONST_BITS;
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    t0 = v[7]; t1 = v[5]; t2 = v[3]; t3 = v[1];
    z1 = t0 + t3;
    int z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
    const int z5 = (z3 + z4) * F_1_175875602;
    t0 *= F_0_298631336; t1 *= F_2_053119869; t2 *= F_3_072711026; t3 *= F_1_501321110;
    z1 = -z1 * F_0_899976223;
    z2 = -z2 * F_2_562915447;
    z3 = -z3 * F_1_961570560 + z5;
    z4 = -z4 * F_0_390180644 + z5;
    t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
    constexpr int n = FIRST ? JCONST_BITS - JPASS1_BITS : JCONST_BITS + JPASS1_BITS + 3;
    v[0] = jdescale(t10 + t3, n); v[7] = jdescale(t10 - t3, n);
    v[1] = jdescale(t11 + t2, n); v[6] = jdescale(t11 - t2, n);
    v[2] = jdescale(t12 + t1, n); v[5] = jdescale(t12 - t1, n);
    v[3] = jdescale(t13 + t0, n); v[4] = jdescale(t13 - t0, n);
}

// `(img.clamp(0, 1) * 255).astype(uint8)`: float32 product, truncation (paragon_otf_degradations.py:119-120)
__device__ __forceinline__ int level8_trunc(float v) { return (int)__fmul_rn(fminf(fmaxf(v, 0.0f), 1.0f), 255.0f); }

constexpr int kJW = 4;          // warps (MCUs) per CTA
// Block rows sit on 16-byte boundaries (pitch 12 words), blocks 104 words apart: a row is two 128-bit shared-memory accesses
// (eight lanes of a block: 16-byte units 3r mod 8, all distinct), and the column pass — lane = (block, column), word
// 104 * block + column + 12 * k — meets 32 distinct banks (8 * block + column mod 32).  Scalar rows at pitch 9 kept the
// shared-memory instruction queue full (ncu: mio_throttle 12 stall cycles per issue).
constexpr int kBP = 12;         // pitch of an 8x8 block row in shared memory
constexpr int kBlk = 8 * kBP + 8;  // words per block
__device__ __forceinline__ void row_load(const int* p, int (&d)[8]) {
    const int4 a = *reinterpret_cast<const int4*>(p), b = *reinterpret_cast<const int4*>(p + 4);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ void row_store(int* p, const int (&d)[8]) {
    *reinterpret_cast<int4*>(p) = make_int4(d[0], d[1], d[2], d[3]);
    *reinterpret_cast<int4*>(p + 4) = make_int4(d[4], d[5], d[6], d[7]);
}

__global__ void __launch_bounds__(32 * kJW) libjpeg_codec_kernel(const float* __restrict__ img, int B, int H, int W, int mcu_x, int mcu_y,
                                                                const __grid_constant__ JpegQuant qt, uint8_t* __restrict__ yplane,
                                                                uint8_t* __restrict__ cplane, int vec_ok) {
    __shared__ __align__(16) int s_blk[kJW][6 * kBlk];  // Y00, Y01, Y10, Y11, Cb, Cr
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int64_t mcu = (int64_t)blockIdx.x * kJW + wrp;
    if (mcu >= (int64_t)B * mcu_x * mcu_y) return;  // warp-uniform
    const int b = (int)(mcu / (mcu_x * mcu_y)), m = (int)(mcu - (int64_t)b * mcu_x * mcu_y);
    const int my = m / mcu_x, mx = m - my * mcu_x;
    int* blk = s_blk[wrp];
    const size_t plane = (size_t)H * W;
    const float* ip = img + (size_t)b * 3 * plane;
    const int ch = (H + 1) >> 1;

    // ---- colour conversion with libjpeg's edge expansion -----------------------------------------------------------------
    // luma: pixel (min(y, H-1), min(x, W-1)).  chroma: columns likewise (expand_right_edge at full resolution); rows are
    // expanded per COMPONENT — chroma row cy >= ch repeats chroma row ch-1, which is built from rows 2(ch-1) and
    // min(2(ch-1)+1, H-1) — so the full-resolution row feeding chroma differs from the luma row inside the padding
    {
        const int ly = lane >> 1, lx0 = (lane & 1) * 8;  // row of the MCU, first of this lane's 8 columns
        const int y = my * 16 + ly;
        const int ry = min(y, H - 1);
        const int cyc = min(y >> 1, ch - 1);
        const int rc = min(2 * cyc + (y & 1), H - 1);
        const float* py = ip + (size_t)ry * W;
        const float* pc = ip + (size_t)rc * W;
        int* yb = blk + ((ly >> 3) * 2 + (lane & 1)) * kBlk + (ly & 7) * kBP;
        // an MCU inside an aligned image (warp-uniform): six 16-byte loads per lane instead of 24 clamped scalar ones
        const bool interior = vec_ok && (mx * 16 + 16 <= W) && (my * 16 + 16 <= H);
        float pr[8], pg[8], pb[8];
        if (interior) {
            const float* p0 = py + mx * 16 + lx0;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(p0) + hf), g4 = __ldg(reinterpret_cast<const float4*>(p0 + plane) + hf),
                             b4 = __ldg(reinterpret_cast<const float4*>(p0 + 2 * plane) + hf);
                pr[4 * hf] = a.x; pr[4 * hf + 1] = a.y; pr[4 * hf + 2] = a.z; pr[4 * hf + 3] = a.w;
                pg[4 * hf] = g4.x; pg[4 * hf + 1] = g4.y; pg[4 * hf + 2] = g4.z; pg[4 * hf + 3] = g4.w;
                pb[4 * hf] = b4.x; pb[4 * hf + 1] = b4.y; pb[4 * hf + 2] = b4.z; pb[4 * hf + 3] = b4.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int x = min(mx * 16 + lx0 + k, W - 1);
                pr[k] = __ldg(py + x); pg[k] = __ldg(py + plane + x); pb[k] = __ldg(py + 2 * plane + x);
            }
        }
        int cbs[4], crs[4];  // horizontal pair sums of the full-resolution Cb / Cr
        int yrow[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = level8_trunc(pr[k]), g = level8_trunc(pg[k]), bl = level8_trunc(pb[k]);
            yrow[k] = ((19595 * r + 38470 * g + 7471 * bl + 32768) >> 16) - 128;  // FIX(0.299), FIX(0.587), FIX(0.114); centred
            int r2 = r, g2 = g, b2 = bl;
            if (rc != ry) {  // (never inside an interior MCU)
                const int x = min(mx * 16 + lx0 + k, W - 1);
                r2 = level8_trunc(__ldg(pc + x)); g2 = level8_trunc(__ldg(pc + plane + x)); b2 = level8_trunc(__ldg(pc + 2 * plane + x));
            }
            // FIX(0.16874) = 11059, FIX(0.33126) = 21709, FIX(0.5) = 32768, FIX(0.41869) = 27439, FIX(0.08131) = 5329;
            // CBCR_OFFSET + ONE_HALF - 1 = (128 << 16) + 32767
            const int cbv = (-11059 * r2 - 21709 * g2 + 32768 * b2 + 8388608 + 32767) >> 16;
            const int crv = (32768 * r2 - 27439 * g2 - 5329 * b2 + 8388608 + 32767) >> 16;
            if (k & 1) { cbs[k >> 1] += cbv; crs[k >> 1] += crv; } else { cbs[k >> 1] = cbv; crs[k >> 1] = crv; }
        }
        row_store(yb, yrow);
        // ---- h2v2_downsample: 2x2 box, bias 1, 2, 1, 2 ... along the component row.  The horizontal pairs sit in this lane,
        //      the row below / above in lane ^ 2: one shuffle per pair sum, the even rows store the four chroma samples ----
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cbs[j] += __shfl_xor_sync(0xffffffffu, cbs[j], 2);
            crs[j] += __shfl_xor_sync(0xffffffffu, crs[j], 2);
        }
        if (!(ly & 1)) {
            const int cy = ly >> 1, cx0 = lx0 >> 1;
            // (an MCU starts at an even component column: the bias alternates with cx)
            *reinterpret_cast<int4*>(blk + 4 * kBlk + cy * kBP + cx0) =
                make_int4(((cbs[0] + 1) >> 2) - 128, ((cbs[1] + 2) >> 2) - 128, ((cbs[2] + 1) >> 2) - 128, ((cbs[3] + 2) >> 2) - 128);
            *reinterpret_cast<int4*>(blk + 5 * kBlk + cy * kBP + cx0) =
                make_int4(((crs[0] + 1) >> 2) - 128, ((crs[1] + 2) >> 2) - 128, ((crs[2] + 1) >> 2) - 128, ((crs[3] + 2) >> 2) - 128);
        }
    }
    __syncwarp();
    // ---- forward DCT rows: 48 row tasks (block, row) ----------------------------------------------------------------------
#pragma unroll
    for (int rnd = 0; rnd < 2; ++rnd) {
        const int task = lane + 32 * rnd;
        if (task < 48) {
            int* p = blk + (task >> 3) * kBlk + (task & 7) * kBP;
            int d[8];
            row_load(p, d);
            fdct8<true>(d);
            row_store(p, d);
        }
    }
    __syncwarp();
    // ---- per column: forward DCT, quantise (jcdctmgr.c: round-half-up division by quantval << 3), de-quantise, inverse
    //      DCT column pass — the eight coefficients of a column never leave the lane ----------------------------------------
#pragma unroll
    for (int rnd = 0; rnd < 2; ++rnd) {
        const int task = lane + 32 * rnd;
        if (task < 48) {
            const int bi = task >> 3, c = task & 7;
            int* p = blk + bi * kBlk + c;
            const uint16_t* q = qt.q[bi >= 4 ? 1 : 0] + c;
            int d[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] = p[k * kBP];
            fdct8<false>(d);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int qv = q[8 * k], dv = qv << 3;
                const int a = abs(d[k]);
                const int r = (a + (dv >> 1)) / dv;
                d[k] = (d[k] < 0 ? -r : r) * qv;
            }
            idct8<true>(d);
#pragma unroll
            for (int k = 0; k < 8; ++k) p[k * kBP] = d[k];
        }
    }
    __syncwarp();
    // ---- inverse DCT rows, range limit, store --------------------------------------------------------------------------------
    const int Wp = mcu_x * 16, Hp = mcu_y * 16, Wc = Wp >> 1, Hc = Hp >> 1;
    uint8_t* yp = yplane + (size_t)b * Hp * Wp;
    uint8_t* cp = cplane + (size_t)b * 2 * Hc * Wc;
#pragma unroll
    for (int rnd = 0; rnd < 2; ++rnd) {
        const int task = lane + 32 * rnd;
        if (task < 48) {
            const int bi = task >> 3, r = task & 7;
            const int* p = blk + bi * kBlk + r * kBP;
            int d[8];
            row_load(p, d);
            idct8<false>(d);
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                lo |= (uint32_t)min(max(d[k] + 128, 0), 255) << (8 * k);
                hi |= (uint32_t)min(max(d[k + 4] + 128, 0), 255) << (8 * k);
            }
            uint8_t* dst;
            if (bi < 4) dst = yp + (size_t)(my * 16 + (bi >> 1) * 8 + r) * Wp + mx * 16 + (bi & 1) * 8;
            else dst = cp + (size_t)(bi - 4) * Hc * Wc + (size_t)(my * 8 + r) * Wc + mx * 8;
            *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);  // (8-byte aligned: plane pitches and offsets are multiples of 8)
        }
    }
}

// jdcolor.c ycc_rgb_convert for one pixel, then the reference's `.float() / 255.0`
__device__ __forceinline__ void ycc_store(float* __restrict__ op, size_t plane, size_t off, int y, int cb, int cr) {
    const int xr = cr - 128, xb = cb - 128;
    const int r = y + ((91881 * xr + 32768) >> 16);                         // FIX(1.40200)
    const int bb = y + ((116130 * xb + 32768) >> 16);                       // FIX(1.77200)
    const int g = y + ((-22554 * xb + 32768 - 46802 * xr) >> 16);           // FIX(0.34414), FIX(0.71414)
    op[off] = div255((float)min(max(r, 0), 255));  // (== __fdiv_rn(x, 255) on [0, 255], three FP32 operations)
    op[plane + off] = div255((float)min(max(g, 0), 255));
    op[2 * plane + off] = div255((float)min(max(bb, 0), 255));
}

__global__ void __launch_bounds__(256) libjpeg_upsample_kernel(const uint8_t* __restrict__ yplane, const uint8_t* __restrict__ cplane, int B,
                                                               int H, int W, int Hp, int Wp, float* __restrict__ out) {
    const int ch = (H + 1) >> 1, cw = (W + 1) >> 1, Hc = Hp >> 1, Wc = Wp >> 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * ch * cw) return;
    const int cx = (int)(idx % cw), cy = (int)((idx / cw) % ch), b = (int)(idx / ((int64_t)cw * ch));
    const uint8_t* yp = yplane + (size_t)b * Hp * Wp;
    const uint8_t* cbp = cplane + (size_t)b * 2 * Hc * Wc;
    const uint8_t* crp = cbp + (size_t)Hc * Wc;
    const size_t plane = (size_t)H * W;
    float* op = out + (size_t)b * 3 * plane;
    // the decoder knows the real extent (ch x cw) of the component: neighbours outside it repeat the edge sample
    const int ym = max(cy - 1, 0), yq = min(cy + 1, ch - 1), xm = max(cx - 1, 0), xq = min(cx + 1, cw - 1);
    int cbv[2][2], crv[2][2];  // [output row v][output column u]
    if (cw > 2) {              // h2v2_fancy_upsample (jdsample.c): 3/4 nearer + 1/4 further, vertically then horizontally
#pragma unroll
        for (int comp = 0; comp < 2; ++comp) {
            const uint8_t* c = comp ? crp : cbp;
            int col[3][3];  // [row: above, this, below][column: left, this, right]
            const int ys[3] = {ym, cy, yq}, xs[3] = {xm, cx, xq};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) col[i][j] = c[(size_t)ys[i] * Wc + xs[j]];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int nb = v ? 2 : 0;
                const int l = 3 * col[1][0] + col[nb][0], t = 3 * col[1][1] + col[nb][1], r = 3 * col[1][2] + col[nb][2];
                // first / last column of the component: (thiscolsum * 4 + 8) >> 4 and (thiscolsum * 4 + 7) >> 4
                const int even = cx == 0 ? (t * 4 + 8) >> 4 : (t * 3 + l + 8) >> 4;
                const int odd = cx == cw - 1 ? (t * 4 + 7) >> 4 : (t * 3 + r + 7) >> 4;
                (comp ? crv : cbv)[v][0] = even;
                (comp ? crv : cbv)[v][1] = odd;
            }
        }
    } else {  // components at most two samples wide: plain replication (jdsample.c h2v2_upsample)
        const int cbs = cbp[(size_t)cy * Wc + cx], crs = crp[(size_t)cy * Wc + cx];
#pragma unroll
        for (int v = 0; v < 2; ++v)
#pragma unroll
            for (int u = 0; u < 2; ++u) { cbv[v][u] = cbs; crv[v][u] = crs; }
    }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        const int y = 2 * cy + v;
        if (y >= H) continue;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int x = 2 * cx + u;
            if (x >= W) continue;
            ycc_store(op, plane, (size_t)y * W + x, yp[(size_t)y * Wp + x], cbv[v][u], crv[v][u]);
        }
    }
}

}  // namespace otf

extern "C" int64_t otf_libjpeg_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return -1;
    const int64_t Hp = (H + 15) / 16 * 16, Wp = (W + 15) / 16 * 16;
    return (B * Hp * Wp * 3 / 2 + 255) / 256 * 256;
}

extern "C" int otf_libjpeg_roundtrip_f32(const float* img, int B, int H, int W, int quality, void* workspace, int64_t workspace_bytes,
                                         float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && workspace, OTF_ERR_BAD_ARG, "libjpeg_roundtrip: null pointer");
    OTF_REQUIRE(B > 0 && H > 0 && W > 0 && H <= 65500 && W <= 65500, OTF_ERR_BAD_ARG, "libjpeg_roundtrip: bad extents (%d, %d, %d)", B, H, W);
    OTF_REQUIRE(workspace_bytes >= otf_libjpeg_workspace_bytes(B, H, W), OTF_ERR_WORKSPACE, "libjpeg_roundtrip: workspace too small");
    OTF_REQUIRE((((uintptr_t)workspace) & 15) == 0, OTF_ERR_BAD_ARG, "libjpeg_roundtrip: workspace must be 16-byte aligned");
    // jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table(force_baseline = TRUE)
    static const uint8_t kLum[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57,
                                     69, 56, 14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64,
                                     81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    static const uint8_t kChr[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                     99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
    const int q = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = q < 50 ? 5000 / q : 200 - 2 * q;
    JpegQuant qt;
    for (int i = 0; i < 64; ++i) {
        int l = (kLum[i] * scale + 50) / 100, c = (kChr[i] * scale + 50) / 100;
        qt.q[0][i] = (uint16_t)(l < 1 ? 1 : (l > 255 ? 255 : l));
        qt.q[1][i] = (uint16_t)(c < 1 ? 1 : (c > 255 ? 255 : c));
    }
    const int mcu_x = (W + 15) / 16, mcu_y = (H + 15) / 16;
    const int Hp = mcu_y * 16, Wp = mcu_x * 16;
    uint8_t* yplane = (uint8_t*)workspace;
    uint8_t* cplane = yplane + (size_t)B * Hp * Wp;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t mcus = (int64_t)B * mcu_x * mcu_y;
    const int vec_ok = (W % 4 == 0) && ((((uintptr_t)img) & 15) == 0);  // rows, planes and MCU starts on 16-byte boundaries
    libjpeg_codec_kernel<<<(unsigned)((mcus + kJW - 1) / kJW), 32 * kJW, 0, st>>>(img, B, H, W, mcu_x, mcu_y, qt, yplane, cplane, vec_ok);
    OTF_LAUNCH_CHECK("libjpeg_codec_kernel");
    const int64_t samples = (int64_t)B * ((H + 1) / 2) * ((W + 1) / 2);
    libjpeg_upsample_kernel<<<(unsigned)((samples + 255) / 256), 256, 0, st>>>(yplane, cplane, B, H, W, Hp, Wp, out);
    OTF_LAUNCH_CHECK("libjpeg_upsample_kernel");
    return OTF_OK;
}

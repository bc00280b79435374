// Gaussian and Poisson noise with Philox4x32-10, plus the reference's clip/round tails.
// Replaces traiNNer/data/degradations.py:569-633 (Gaussian) and :762-842 (Poisson):
//   torch.rand/randn/poisson + ~10 elementwise launches + 2*B torch.unique host syncs
// with one streaming kernel (Gaussian) or two (Poisson: presence masks, then sampling).
// HBM-bound: Gaussian reads N, writes N; Poisson reads N twice, writes N.
#include <stdlib.h>

#include "otf_common.cuh"

namespace otf {

// ------------------------------------------------------------------ Gaussian ----
// grid = (chunks, B): a CTA works inside one sample, so sigma/gray are block-uniform and all index
// math is 32-bit.  One thread = 4 consecutive elements of the sample's flat (C,H,W) block, two such
// quads per loop iteration (two independent Philox chains in flight).
// Philox use per quad: one call for the colour field; one more for the batch-shared gray field,
// and only for samples whose gray flag is set (flag 0 -> noise*1 + ng*0 == noise exactly, flag 1 ->
// noise*0 + ng*1 == ng exactly, so the unused field is never generated).
// EXACT (injected fields, or the `rounds` tail whose rint() is a cliff) evaluates the reference's
// operations one by one: (N*sigma)/255, noise*(1-g) + ng*g.  The production path folds the
// block-uniform factors first — N*(sigma/255*(1-g)) + G*(sigma/255*g) — which differs from the
// reference by <= 2 ulp of the noise (~1e-9 on a [0,1] image) and costs 2 FMAs instead of 2 IEEE
// divisions + 5 ops per element: at 6.4 TB/s an elementwise kernel has ~6 issue slots per float4.
// Philox positions are laid out per image ROW: a quad is four consecutive pixels of one row, quad (row, xq) of sample b
// has index b*(C*H*QW) + row*QW + xq with QW = ceil(W/4) (row = c*H + y), the shared gray field uses y*QW + xq.  For
// W % 4 == 0 this is the flat quad numbering; for any other width rows simply end with a partial quad.  The row layout is
// what lets the resize kernels draw the very same field in their epilogue (resize.cu: four lanes own four consecutive
// output columns), so that the fused resize + noise launch is bit-identical to resize followed by this kernel.
struct GaussQuad {
    float v[4], nc[4], ng[4];
};

template <bool VEC, bool EXACT>
__device__ __forceinline__ void gauss_load(GaussQuad& g, const Philox& ph, const float* __restrict__ ip, int q, int nq,
                                           int W, int QW, int H, size_t base, int b, const float* __restrict__ ncol,
                                           const float* __restrict__ ngray, bool need_color, bool need_gray,
                                           uint64_t offset) {
    // VEC: W % 4 == 0, the quad's first element is 4q;  otherwise row = q / QW, xq = q % QW
    const int row = VEC ? 0 : q / QW, xq = VEC ? 0 : q - row * QW;
    const int e0 = VEC ? (q << 2) : row * W + (xq << 2);
    const int cnt = VEC ? 4 : min(4, W - (xq << 2));
#pragma unroll
    for (int k = 0; k < 4; ++k) { g.v[k] = 0.f; g.nc[k] = 0.f; g.ng[k] = 0.f; }
    if (VEC) {
        const float4 t = *reinterpret_cast<const float4*>(ip + e0);
        g.v[0] = t.x; g.v[1] = t.y; g.v[2] = t.z; g.v[3] = t.w;
    } else {
        for (int k = 0; k < cnt; ++k) g.v[k] = ip[e0 + k];
    }
    if (ncol) {
        for (int k = 0; k < cnt; ++k) g.nc[k] = ncol[base + e0 + k];
    } else if (need_color) {
        const float4 t = normal4(ph, (uint64_t)b * nq + q, offset * 8 + STREAM_COLOR);
        g.nc[0] = t.x; g.nc[1] = t.y; g.nc[2] = t.z; g.nc[3] = t.w;
    }
    if (need_gray) {
        const int hw = H * W;
        const int p0 = e0 % hw;  // pixel index of the first element inside its channel plane
        if (ngray) {
            for (int k = 0; k < cnt; ++k) g.ng[k] = ngray[p0 + k];
        } else {
            // ONE (h,w) field shared by the whole batch and all channels (degradations.py:593-596)
            const int gq = VEC ? (p0 >> 2) : (row % H) * QW + xq;
            const float4 t = normal4(ph, (uint64_t)gq, offset * 8 + STREAM_GRAY);
            g.ng[0] = t.x; g.ng[1] = t.y; g.ng[2] = t.z; g.ng[3] = t.w;
        }
    }
}

template <bool VEC, bool EXACT>
__global__ void __launch_bounds__(256) gaussian_noise_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                             int C, int H, int W,
                                                             const float* __restrict__ sigma, const float* __restrict__ gray,
                                                             const float* __restrict__ ncol, const float* __restrict__ ngray,
                                                             uint64_t seed, uint64_t offset,
                                                             const uint64_t* __restrict__ offset_dev, int flags) {
    pdl_enter();
    const Philox ph(seed);
    if (offset_dev) offset += *offset_dev;  // device-resident counter: a replayed CUDA graph draws a fresh field
    const int b = blockIdx.y;
    const bool raw = (flags & OTF_NOISE_RAW_FIELD) != 0;  // ncol is the finished noise field: added as it is
    const float sg = raw ? 0.0f : sigma[b];
    const float g = gray ? gray[b] : 0.0f;
    const bool use_gray = gray != nullptr && !raw;
    const float one_minus_g = __fsub_rn(1.0f, g);
    const float s255 = __fdiv_rn(sg, 255.0f);
    const float ca = use_gray ? s255 * one_minus_g : s255, cb = use_gray ? s255 * g : 0.0f;  // folded factors
    const int QW = (W + 3) >> 2;
    const int chw = C * H * W;
    const int nq = C * H * QW;
    const size_t base = (size_t)b * chw;
    const float* ip = img + base;
    float* op = out + base;
    const bool need_color = !(use_gray && g == 1.0f);
    const bool need_gray = use_gray && (ncol != nullptr || g != 0.0f);
    const int stride = gridDim.x * blockDim.x;
    for (int q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < nq; q0 += 2 * stride) {
        GaussQuad gq[2];
        const int qs[2] = {q0, q0 + stride};
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (qs[u] < nq) gauss_load<VEC, EXACT>(gq[u], ph, ip, qs[u], nq, W, QW, H, base, b, ncol, ngray, need_color, need_gray, offset);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (qs[u] >= nq) continue;
            const int row = VEC ? 0 : qs[u] / QW, xq = VEC ? 0 : qs[u] - row * QW;
            const int e0 = VEC ? (qs[u] << 2) : row * W + (xq << 2);
            const int cnt = VEC ? 4 : min(4, W - (xq << 2));
            float r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float noise;
                if (EXACT && raw) {
                    noise = gq[u].nc[k];  // degradations.py:625 / :833: out = img + noise, the field generated elsewhere
                } else if (EXACT) {
                    // degradations.py:598: noise = randn * sigma / 255 ; :603: noise*(1-gray) + noise_gray*gray
                    noise = __fdiv_rn(__fmul_rn(gq[u].nc[k], sg), 255.0f);
                    if (use_gray) {
                        const float ngv = __fdiv_rn(__fmul_rn(gq[u].ng[k], sg), 255.0f);
                        noise = __fadd_rn(__fmul_rn(noise, one_minus_g), __fmul_rn(ngv, g));
                    }
                } else {
                    noise = fmaf(gq[u].ng[k], cb, gq[u].nc[k] * ca);
                }
                r[k] = (flags & OTF_NOISE_FIELD_ONLY) ? noise : noise_tail(__fadd_rn(gq[u].v[k], noise), flags);
            }
            if (VEC) {
                *reinterpret_cast<float4*>(op + e0) = make_float4(r[0], r[1], r[2], r[3]);
            } else {
                for (int k = 0; k < cnt; ++k) op[e0 + k] = r[k];
            }
        }
    }
}

__global__ void philox_fill_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t offset, int normal) {
    const Philox ph(seed);
    const int64_t nq = (n + 3) >> 2;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        float t[4];
        if (normal) {
            const float4 v = normal4(ph, (uint64_t)q, offset * 8 + STREAM_COLOR);
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        } else {
            const uint4 r = ph((uint64_t)q, offset * 8 + STREAM_COLOR);
            t[0] = u01(r.x); t[1] = u01(r.y); t[2] = u01(r.z); t[3] = u01(r.w);
        }
        for (int k = 0; k < 4; ++k)
            if ((q << 2) + k < n) out[(q << 2) + k] = t[k];
    }
}

// ------------------------------------------------------------------- Poisson ----
// Counter-based Poisson(lambda) sampler. Each sample owns Philox counters (index, stream) with
// sub-counters in the top bits of the stream word, so it can draw as many uniforms as it needs.
struct UniformStream {
    const Philox& ph;
    uint64_t index, stream;
    uint4 buf;
    int used;
    uint32_t sub;
    __device__ UniformStream(const Philox& p, uint64_t idx, uint64_t st) : ph(p), index(idx), stream(st), used(4), sub(0) {}
    __device__ __forceinline__ float next() {
        if (used == 4) {
            buf = ph(index, stream + ((uint64_t)sub << 40));
            ++sub;
            used = 0;
        }
        const uint32_t w = used == 0 ? buf.x : used == 1 ? buf.y : used == 2 ? buf.z : buf.w;
        ++used;
        return u01(w);
    }
};

// log(k!) — exact table for small k, Stirling series otherwise (abs err < 1e-7 for k >= 10 before the
// ~1e-6 relative error of the MUFU logarithm)
__constant__ float c_logfact[10] = {0.0f, 0.0f, 0.69314718f, 1.79175947f, 3.17805383f, 4.78749174f,
                                    6.57925121f, 8.52516136f, 10.60460290f, 12.80182748f};
__device__ __forceinline__ float log_factorial(float k) {
    if (k < 10.0f) return c_logfact[(int)k];
    const float x = k + 1.0f;
    const float ix = __fdividef(1.0f, x), ix2 = ix * ix;
    return (x - 0.5f) * __logf(x) - x + 0.91893853f + ix * (0.083333333f - ix2 * (0.0027777778f - ix2 * 0.00079365079f));
}

// Poisson(lambda), lambda in [0, 256].  MUFU-based exp/log/div/sqrt: the acceptance inequality of the
// rejection branch is evaluated to ~1e-4 absolute on the log scale, a bias far below what a chi-square
// test at 2M samples can see (tests/test_rng_gpu.py), for roughly a fifth of the instructions.
__device__ float poisson_sample(float lam, UniformStream& us) {
    if (!(lam > 0.0f)) return 0.0f;
    if (lam < 10.0f) {
        // inversion by sequential search on the CDF (one uniform)
        float p = __expf(-lam), k = 0.0f;
        const float u = us.next();
        float cdf = p;
        while (u > cdf && k < 100.0f) {
            k += 1.0f;
            p *= __fdividef(lam, k);
            cdf += p;
        }
        return k;
    }
    // PTRS — W. Hörmann, "The transformed rejection method for generating Poisson random
    // variables", Insurance: Mathematics and Economics 12 (1993) 39-45.
    float slam;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(slam) : "f"(lam));
    const float loglam = __logf(lam);
    const float bb = 0.931f + 2.53f * slam;
    const float a = -0.059f + 0.02483f * bb;
    const float invalpha = 1.1239f + __fdividef(1.1328f, bb - 3.4f);
    const float vr = 0.9277f - __fdividef(3.6224f, bb - 2.0f);
    for (int it = 0; it < 64; ++it) {
        const float U = us.next() - 0.5f;
        const float V = us.next();
        const float us_ = 0.5f - fabsf(U);
        const float k = floorf((__fdividef(2.0f * a, us_) + bb) * U + lam + 0.43f);
        if (us_ >= 0.07f && V <= vr) return k;
        if (k < 0.0f || (us_ < 0.013f && V > us_)) continue;
        // log(V) + log(invalpha) - log(a/us^2 + b) as one logarithm
        if (__logf(__fdividef(V * invalpha, __fdividef(a, us_ * us_) + bb)) <= -lam + k * loglam - log_factorial(k)) return k;
    }
    return floorf(lam + 0.5f);
}

__global__ void philox_poisson_kernel(const float* __restrict__ lam, float* __restrict__ out, int64_t n, uint64_t seed,
                                      uint64_t offset) {
    const Philox ph(seed);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        UniformStream us(ph, (uint64_t)i, offset * 8 + STREAM_POIS_COLOR);
        out[i] = poisson_sample(lam[i], us);
    }
}

__device__ __forceinline__ float gray_of(float r, float g, float b) {
    // torchvision rgb_to_grayscale (degradations.py:787): (0.2989*r + 0.587*g + 0.114*b), left to right
    return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}
__device__ __forceinline__ int level8(float x) {  // clamp(round(x*255),0,255) as an integer level
    return (int)fminf(fmaxf(rintf(__fmul_rn(x, 255.0f)), 0.0f), 255.0f);
}

// Pass 1: 256-bit presence masks per sample (colour over C,H,W; gray over H,W).
// masks[b*16 + 0..7] colour, [b*16 + 8..15] gray.  grid = (chunks, B).
__global__ void __launch_bounds__(256) poisson_presence_kernel(const float* __restrict__ img, int hw,
                                                               uint32_t* __restrict__ masks) {
    pdl_enter();
    // One byte flag per level in shared memory: setting it is a plain store (every writer stores the same 1, so the
    // race is benign) — 4 stores per pixel instead of the ~100 select instructions of per-thread register masks or the
    // contended atomics of a shared bit mask.  At the end 16 ballots pack the flags into the sample's 2 x 256-bit masks.
    const int b = blockIdx.y;
    const float* ip = img + (size_t)b * 3 * hw;
    __shared__ uint8_t flag[512];  // [0,256) colour levels, [256,512) gray levels
    flag[threadIdx.x] = 0;
    flag[256 + threadIdx.x] = 0;
    __syncthreads();
    if ((hw & 3) == 0 && (((uintptr_t)img) & 15) == 0) {  // three independent 16-byte loads per trip
        const float4* i4 = reinterpret_cast<const float4*>(ip);
        const int q4 = hw >> 2;
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < q4; p += gridDim.x * blockDim.x) {
            const float4 r = __ldg(i4 + p), g = __ldg(i4 + q4 + p), bl = __ldg(i4 + 2 * q4 + p);
            const float rr[4] = {r.x, r.y, r.z, r.w}, gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {bl.x, bl.y, bl.z, bl.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                flag[level8(rr[k])] = 1;
                flag[level8(gg[k])] = 1;
                flag[level8(bb[k])] = 1;
                flag[256 + level8(gray_of(rr[k], gg[k], bb[k]))] = 1;
            }
        }
    } else {
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
            const float r = ip[p], g = ip[hw + p], bl = ip[2 * hw + p];
            flag[level8(r)] = 1;
            flag[level8(g)] = 1;
            flag[level8(bl)] = 1;
            flag[256 + level8(gray_of(r, g, bl))] = 1;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;  // warp w packs levels [32w, 32w+32) of both kinds
    const uint32_t mc = __ballot_sync(0xffffffffu, flag[threadIdx.x] != 0);
    const uint32_t mg = __ballot_sync(0xffffffffu, flag[256 + threadIdx.x] != 0);
    if (lane == 0) {
        if (mc) atomicOr(&masks[b * 16 + wid], mc);
        if (mg) atomicOr(&masks[b * 16 + 8 + wid], mg);
    }
}

__device__ __forceinline__ float vals_from_mask(const uint32_t* m) {
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) cnt += __popc(m[i]);
    // 2 ** ceil(log2(cnt)) — degradations.py:792, :803
    int v = 1;
    while (v < cnt) v <<= 1;
    return (float)v;
}

// Pass 2: one thread per pixel (all three channels). grid = (chunks, B).
// (Tried and dropped in round 1: a per-thread state machine that runs one PTRS trial per loop iteration so that
// lanes never wait for the slowest rejection loop — bit-identical output, but 0.23 ms vs 0.165 ms here at
// 64x3x192^2: the bookkeeping costs more than the ~1.5 extra warp-level trials it saves.)
// one pixel (all three channels) of the generic path: any gray flag, injected counts, exported lambdas
__device__ __forceinline__ void poisson_pixel_generic(const float* __restrict__ ip, float* __restrict__ op, int hw, int b, int p,
                                                      float vc, float vg, float sc, const float* gray, float gf,
                                                      const float* __restrict__ counts_c, const float* __restrict__ counts_g,
                                                      const Philox& ph, uint64_t offset, int flags, float* __restrict__ lam_c_out,
                                                      float* __restrict__ lam_g_out) {
    const float px[3] = {ip[p], ip[hw + p], ip[2 * hw + p]};
    // The gray flag is block-uniform.  flag 0: noise*1 + noise_g*0 == noise exactly, flag 1: noise*0 + noise_g*1
    // == noise_g exactly — so the field that the mix discards is not sampled (unless a test asked for its lambda
    // or injected its counts).
    const bool need_g = gray && (gf != 0.0f || counts_g || lam_g_out);
    const bool need_c = !(gray && gf == 1.0f) || counts_c || lam_c_out;
    float noise_g = 0.0f;
    if (need_g) {
        // degradations.py:787-795 — gray image, quantised; noise relative to the quantised value
        const float qg = quantise8(gray_of(px[0], px[1], px[2]));
        const float lam = __fmul_rn(qg, vg);
        if (lam_g_out) lam_g_out[(size_t)b * hw + p] = lam;
        float cnt;
        if (counts_g) {
            cnt = counts_g[(size_t)b * hw + p];
        } else {
            UniformStream us(ph, (uint64_t)b * hw + p, offset * 8 + STREAM_POIS_GRAY);
            cnt = poisson_sample(lam, us);
        }
        noise_g = __fsub_rn(__fdiv_rn(cnt, vg), qg);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const size_t e = (size_t)b * 3 * hw + (size_t)c * hw + p;
        float noise = 0.0f;
        if (need_c) {
            const float qc = quantise8(px[c]);  // :800
            const float lam = __fmul_rn(qc, vc);
            if (lam_c_out) lam_c_out[e] = lam;
            float cnt;
            if (counts_c) {
                cnt = counts_c[e];
            } else {
                UniformStream us(ph, (uint64_t)e, offset * 8 + STREAM_POIS_COLOR);
                cnt = poisson_sample(lam, us);
            }
            noise = __fsub_rn(__fdiv_rn(cnt, vc), qc);  // :805-806
        }
        if (gray) noise = __fadd_rn(__fmul_rn(noise, __fsub_rn(1.0f, gf)), __fmul_rn(noise_g, gf));  // :808
        noise = __fmul_rn(noise, sc);                                                              // :811
        op[(size_t)c * hw + p] =
            (flags & OTF_NOISE_FIELD_ONLY) ? noise : noise_tail(__fadd_rn(px[c], noise), flags);           // :834-841
    }
}

__global__ void __launch_bounds__(256) poisson_apply_kernel(const float* __restrict__ img, float* __restrict__ out, int hw,
                                                            const float* __restrict__ scale, const float* __restrict__ gray,
                                                            const float* __restrict__ counts_c, const float* __restrict__ counts_g,
                                                            uint64_t seed, uint64_t offset, const uint64_t* __restrict__ offset_dev,
                                                            int flags,
                                                            const uint32_t* __restrict__ masks, float* __restrict__ vals_out,
                                                            float* __restrict__ lam_c_out, float* __restrict__ lam_g_out) {
    pdl_enter();
    if (offset_dev) offset += *offset_dev;
    const int b = blockIdx.y;
    __shared__ float s_vals[2];
    if (threadIdx.x < 2) {
        s_vals[threadIdx.x] = vals_from_mask(masks + b * 16 + threadIdx.x * 8);
        if (vals_out && blockIdx.x == 0) vals_out[b * 2 + threadIdx.x] = s_vals[threadIdx.x];
    }
    __syncthreads();
    const float vc = s_vals[0], vg = s_vals[1];
    const float sc = scale[b];
    const float gf = gray ? gray[b] : 0.0f;
    const Philox ph(seed);
    const float* ip = img + (size_t)b * 3 * hw;
    float* op = out + (size_t)b * 3 * hw;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x)
        poisson_pixel_generic(ip, op, hw, b, p, vc, vg, sc, gray, gf, counts_c, counts_g, ph, offset, flags, lam_c_out, lam_g_out);
}

// ---- alias tables (the production path) --------------------------------------------------------------
// lambda = q * vals with q = level / 255 and vals = 2^v, v in 0..8 (a sample has at most 256 distinct levels): only
// 9 x 256 = 2304 different lambdas can ever occur, whatever the image.  For each of them the Poisson pmf on a
// 256-wide window [k_lo, k_lo + 256) around lambda (>= 7.9 sigma at lambda = 256, tail mass < 1e-14) is turned ONCE per
// device into a Walker / Vose alias table with 256 columns (fp64 construction, otf_poisson_build_tables): 2304 rows x
// 256 columns x 8 bytes = 4.7 MB, L2 resident.  A sample then costs ONE 24-bit uniform and ONE 8-byte table read: the top
// 8 bits pick the column, the low 16 bits are compared with the column's 16-bit threshold and select the column's own
// count or its alias — no search, no loop, no divergence, one Philox call per pixel, and every outcome's probability is
// right to 2^-24 (the resolution of the uniform).  Round 1 inverted the CDF instead (a guide byte, k_lo and ~1.3 CDF
// reads per sample): ncu showed that kernel waiting on its ~3.3 scattered reads per sample (long-scoreboard 8.2) — the
// alias form makes it one.  The rejection kernel above pays ~2.6 warp-level PTRS attempts per element.
constexpr int kPoisWin = 256, kPoisRows = 9 * 256;

struct PoissonTables {
    const uint2* alias;  // [9][256][256]: .x = threshold (0..65536), .y = own count | alias count << 16
};
static int64_t poisson_tables_bytes() { return (int64_t)kPoisRows * kPoisWin * 8; }
static PoissonTables poisson_tables(const void* dev) {
    PoissonTables t;
    t.alias = (const uint2*)dev;
    return t;
}

// one thread per (vals exponent, level) row — runs once per device
__global__ void __launch_bounds__(64) poisson_tables_kernel(uint2* __restrict__ alias_out) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= kPoisRows) return;
    const int v = row >> 8, level = row & 255;
    const float lamf = __fmul_rn(__fdiv_rn((float)level, 255.0f), (float)(1 << v));  // exactly what the apply kernel forms
    const double lam = (double)lamf;
    int klo = (int)floor(lam - 8.0 * sqrt(lam));
    if (klo < 0) klo = 0;
    // scaled pmf: pr[k] = 256 * P(klo + k), renormalised over the window (the mass outside is < 1e-14)
    double pr[kPoisWin];
    uint8_t small[kPoisWin], large[kPoisWin];
    double p = lam > 0.0 ? exp(-lam + klo * log(lam) - lgamma((double)klo + 1.0)) : (klo == 0 ? 1.0 : 0.0);
    double tot = 0.0;
    for (int k = 0; k < kPoisWin; ++k) {
        pr[k] = p;
        tot += p;
        p = p * lam / (double)(klo + k + 1);
    }
    int ns = 0, nl = 0;
    for (int k = 0; k < kPoisWin; ++k) {
        pr[k] = pr[k] / tot * (double)kPoisWin;
        if (pr[k] < 1.0) small[ns++] = (uint8_t)k; else large[nl++] = (uint8_t)k;
    }
    uint2* out = alias_out + (size_t)row * kPoisWin;
    // Vose: pair a deficient column with a surplus one until one list runs out
    while (ns > 0 && nl > 0) {
        const int sidx = small[--ns], lidx = large[nl - 1];
        const double thr = fmin(fmax(pr[sidx], 0.0), 1.0);
        out[sidx] = make_uint2((unsigned)llrint(thr * 65536.0), (unsigned)(klo + sidx) | ((unsigned)(klo + lidx) << 16));
        pr[lidx] = (pr[lidx] + pr[sidx]) - 1.0;
        if (pr[lidx] < 1.0) { --nl; small[ns++] = (uint8_t)lidx; }
    }
    while (nl > 0) { const int k = large[--nl]; out[k] = make_uint2(65536u, (unsigned)(klo + k) | ((unsigned)(klo + k) << 16)); }
    while (ns > 0) { const int k = small[--ns]; out[k] = make_uint2(65536u, (unsigned)(klo + k) | ((unsigned)(klo + k) << 16)); }  // rounding leftovers
}

__device__ __forceinline__ float poisson_by_table(const PoissonTables& tab, int row, uint32_t bits) {
    const uint32_t u24 = bits >> 8;
    const uint2 e = __ldg(tab.alias + (size_t)row * kPoisWin + (u24 >> 16));
    return (float)(((u24 & 0xffffu) < e.x) ? (e.y & 0xffffu) : (e.y >> 16));
}

// Pass 2, production flags (no injected counts, no exports, per-sample gray flag exactly 0 or 1): one thread per
// pixel, one Philox call per pixel (3 of its 4 words for the colour channels, 1 for a gray sample).
__global__ void __launch_bounds__(256) poisson_apply_table_kernel(const float* __restrict__ img, float* __restrict__ out, int hw,
                                                                  const float* __restrict__ scale, const float* __restrict__ gray,
                                                                  uint64_t seed, uint64_t offset,
                                                                  const uint64_t* __restrict__ offset_dev, int flags,
                                                                  const uint32_t* __restrict__ masks, PoissonTables tab) {
    pdl_enter();
    if (offset_dev) offset += *offset_dev;
    const int b = blockIdx.y;
    const float gf = gray ? gray[b] : 0.0f;
    const bool is_gray = gray && gf == 1.0f;
    __shared__ float s_vals[2];
    if (threadIdx.x < 2) s_vals[threadIdx.x] = vals_from_mask(masks + b * 16 + threadIdx.x * 8);
    __syncthreads();
    const float sc = scale[b];
    const Philox ph(seed);
    const float* ip = img + (size_t)b * 3 * hw;
    float* op = out + (size_t)b * 3 * hw;
    if (gray && gf != 0.0f && gf != 1.0f) {
        // a fractional flag mixes both fields (the API allows it; the reference only ever draws 0 or 1): rejection sampler
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x)
            poisson_pixel_generic(ip, op, hw, b, p, s_vals[0], s_vals[1], sc, gray, gf, nullptr, nullptr, ph, offset, flags, nullptr, nullptr);
        return;
    }
    const float vsel = s_vals[is_gray ? 1 : 0];
    const float inv_vals = 1.0f / vsel;  // vals is a power of two: the reciprocal and the product are exact
    const int row0 = (31 - __clz((int)vsel)) * 256;  // table rows of this sample's vals
    __shared__ float s_q[256];  // level / 255 with the IEEE division the reference performs, once per CTA
    s_q[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);
    __syncthreads();
    const uint64_t stream = offset * 8 + (is_gray ? STREAM_POIS_GRAY : STREAM_POIS_COLOR);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
        const float px[3] = {ip[p], ip[hw + p], ip[2 * hw + p]};
        const uint4 r = ph((uint64_t)b * hw + p, stream);
        float noise[3];
        if (is_gray) {
            const int lv = level8(gray_of(px[0], px[1], px[2]));
            const float cnt = poisson_by_table(tab, row0 + lv, r.x);
            // mix with flag 1: noise * 0 + noise_g * 1 == noise_g exactly (degradations.py:808)
            noise[0] = noise[1] = noise[2] = __fsub_rn(__fmul_rn(cnt, inv_vals), s_q[lv]);
        } else {
            // the three channels' table reads are independent: issue them side by side
            const uint32_t w[3] = {r.x, r.y, r.z};
            int lv[3];
            uint2 e[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                lv[ch] = level8(px[ch]);
                e[ch] = __ldg(tab.alias + (size_t)(row0 + lv[ch]) * kPoisWin + (w[ch] >> 24));
            }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float cnt = (float)((((w[ch] >> 8) & 0xffffu) < e[ch].x) ? (e[ch].y & 0xffffu) : (e[ch].y >> 16));
                noise[ch] = __fsub_rn(__fmul_rn(cnt, inv_vals), s_q[lv[ch]]);  // :805-806
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float nz = __fmul_rn(noise[c], sc);  // :811
            op[(size_t)c * hw + p] = (flags & OTF_NOISE_FIELD_ONLY) ? nz : noise_tail(__fadd_rn(px[c], nz), flags);
        }
    }
}

static int stream_grid(int64_t work_items, int threads) {
    int64_t blocks = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace otf

extern "C" int otf_gaussian_noise_f32(const float* img, int B, int C, int H, int W, const float* sigma_dev,
                                      const float* gray_dev, const float* noise_color_dev, const float* noise_gray_dev,
                                      uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int flags, float* out,
                                      void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && (sigma_dev || (flags & OTF_NOISE_RAW_FIELD)), OTF_ERR_BAD_ARG, "gaussian_noise: null pointer");
    OTF_REQUIRE(!(flags & OTF_NOISE_RAW_FIELD) || (noise_color_dev && !noise_gray_dev), OTF_ERR_BAD_ARG,
                "gaussian_noise: OTF_NOISE_RAW_FIELD takes the finished field in noise_color_dev (and no gray field)");
    OTF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OTF_ERR_BAD_ARG, "gaussian_noise: bad extents");
    OTF_REQUIRE(!(noise_gray_dev && !noise_color_dev), OTF_ERR_BAD_ARG, "gaussian_noise: inject both fields or neither");
    OTF_REQUIRE((int64_t)C * H * W < (1ll << 30), OTF_ERR_UNSUPPORTED, "gaussian_noise: sample too large");
    OTF_REQUIRE(B <= 65535, OTF_ERR_UNSUPPORTED, "gaussian_noise: B > 65535");
    if (flags & OTF_NOISE_RAW_FIELD) gray_dev = nullptr;
    const float* ng = gray_dev ? noise_gray_dev : nullptr;
    OTF_REQUIRE(!(gray_dev && noise_color_dev && !noise_gray_dev), OTF_ERR_BAD_ARG,
                "gaussian_noise: gray flags with an injected colour field need the injected gray field too");
    const bool vec = (W % 4 == 0) && (((uintptr_t)img & 15) == 0) && (((uintptr_t)out & 15) == 0);
    int chunks = ceil_div((int64_t)C * H * ((W + 3) / 4), 512);
    const int cap = ceil_div(kNumSMs * 16, B);
    if (chunks > cap) chunks = cap;
    const dim3 grid(chunks, B);
    const bool exact = noise_color_dev != nullptr || (flags & OTF_NOISE_ROUNDS);
    cudaStream_t st = (cudaStream_t)stream;
#define OTF_GAUSS(V, E)                                                                                              \
    launch_chain(gaussian_noise_kernel<V, E>, dim3(grid), dim3(256), 0, st, img, out, C, H, W, sigma_dev, gray_dev, noise_color_dev, ng, \
                                                      seed, offset, offset_dev, flags)
    if (vec && exact) OTF_GAUSS(true, true);
    else if (vec) OTF_GAUSS(true, false);
    else if (exact) OTF_GAUSS(false, true);
    else OTF_GAUSS(false, false);
#undef OTF_GAUSS
    OTF_LAUNCH_CHECK("gaussian_noise_kernel");
    return OTF_OK;
}

extern "C" int otf_philox_normal_f32(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
    using namespace otf;
    OTF_REQUIRE(out && n > 0, OTF_ERR_BAD_ARG, "philox_normal: bad args");
    philox_fill_kernel<<<stream_grid((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset, 1);
    OTF_LAUNCH_CHECK("philox_fill_kernel");
    return OTF_OK;
}

extern "C" int otf_philox_uniform_f32(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
    using namespace otf;
    OTF_REQUIRE(out && n > 0, OTF_ERR_BAD_ARG, "philox_uniform: bad args");
    philox_fill_kernel<<<stream_grid((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset, 0);
    OTF_LAUNCH_CHECK("philox_fill_kernel");
    return OTF_OK;
}

extern "C" int otf_philox_poisson_f32(const float* lambda_dev, float* out, int64_t n, uint64_t seed, uint64_t offset,
                                      void* stream) {
    using namespace otf;
    OTF_REQUIRE(lambda_dev && out && n > 0, OTF_ERR_BAD_ARG, "philox_poisson: bad args");
    philox_poisson_kernel<<<stream_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(lambda_dev, out, n, seed, offset);
    OTF_LAUNCH_CHECK("philox_poisson_kernel");
    return OTF_OK;
}

extern "C" int64_t otf_poisson_tables_bytes(void) { return otf::poisson_tables_bytes(); }

extern "C" int otf_poisson_build_tables(void* tables_dev, void* stream) {
    using namespace otf;
    OTF_REQUIRE(tables_dev && (((uintptr_t)tables_dev) & 15) == 0, OTF_ERR_BAD_ARG, "poisson_build_tables: null or unaligned pointer");
    char* p = (char*)tables_dev;
    poisson_tables_kernel<<<kPoisRows / 64, 64, 0, (cudaStream_t)stream>>>((uint2*)p);
    OTF_LAUNCH_CHECK("poisson_tables_kernel");
    return OTF_OK;
}

extern "C" int otf_poisson_noise_f32(const float* img, int B, int C, int H, int W, const float* scale_dev,
                                     const float* gray_dev, const float* counts_color_dev, const float* counts_gray_dev,
                                     uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int flags, uint32_t* masks_dev,
                                     const void* tables_dev, float* vals_out_dev, float* lambda_color_dev, float* lambda_gray_dev,
                                     float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && scale_dev && masks_dev, OTF_ERR_BAD_ARG, "poisson_noise: null pointer");
    OTF_REQUIRE(C == 3, OTF_ERR_UNSUPPORTED, "poisson_noise: C must be 3 (rgb_to_grayscale), got %d", C);
    OTF_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, OTF_ERR_BAD_ARG, "poisson_noise: bad extents");
    OTF_REQUIRE(!(lambda_gray_dev && !gray_dev), OTF_ERR_BAD_ARG, "poisson_noise: lambda_gray needs gray flags");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = H * W;
    cudaError_t e = cudaMemsetAsync(masks_dev, 0, (size_t)B * 16 * sizeof(uint32_t), st);
    if (e != cudaSuccess) return cuda_fail(e, "poisson masks memset");
    int chunks = ceil_div(hw, 256 * 4);
    const int max_chunks = ceil_div(kNumSMs * 8, B);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    launch_chain(poisson_presence_kernel, dim3(dim3(chunks, B)), dim3(256), 0, st, img, hw, masks_dev);
    OTF_LAUNCH_CHECK("poisson_presence_kernel");
    int chunks2 = ceil_div(hw, 256);
    const int max_chunks2 = ceil_div(kNumSMs * 16, B);
    if (chunks2 > max_chunks2) chunks2 = max_chunks2;
    // production flags + room for the CDF tables: exact table inversion; OTF_POISSON_IMPL=ptrs forces the rejection sampler
    static const bool force_ptrs = [] { const char* e = getenv("OTF_POISSON_IMPL"); return e && e[0] == 'p'; }();
    const bool plain = !counts_color_dev && !counts_gray_dev && !vals_out_dev && !lambda_color_dev && !lambda_gray_dev;
    if (plain && !force_ptrs && tables_dev) {
        launch_chain(poisson_apply_table_kernel, dim3(dim3(chunks2, B)), dim3(256), 0, st, img, out, hw, scale_dev, gray_dev, seed, offset, offset_dev, flags,
                                                                     masks_dev, poisson_tables(tables_dev));
        OTF_LAUNCH_CHECK("poisson_apply_table_kernel");
        return OTF_OK;
    }
    launch_chain(poisson_apply_kernel, dim3(dim3(chunks2, B)), dim3(256), 0, st, img, out, hw, scale_dev, gray_dev, counts_color_dev,
                                                           counts_gray_dev, seed, offset, offset_dev, flags, masks_dev, vals_out_dev,
                                                           lambda_color_dev, lambda_gray_dev);
    OTF_LAUNCH_CHECK("poisson_apply_kernel");
    return OTF_OK;
}

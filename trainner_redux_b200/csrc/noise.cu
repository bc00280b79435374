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
struct GaussQuad {
    float v[4], nc[4], ng[4];
};

template <bool VEC, bool EXACT>
__device__ __forceinline__ void gauss_load(GaussQuad& g, const Philox& ph, const float* __restrict__ ip, int q, int nq,
                                           int chw, int hw, size_t base, int b, const float* __restrict__ ncol,
                                           const float* __restrict__ ngray, bool need_color, bool need_gray,
                                           uint64_t offset) {
    const int e0 = q << 2;
    const int cnt = min(4, chw - e0);
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
        const int p0 = e0 % hw;  // pixel index of the first element inside its channel plane
        if (ngray) {
            for (int k = 0; k < cnt; ++k) { int p = p0 + k; if (p >= hw) p -= hw; g.ng[k] = ngray[p]; }
        } else if ((p0 & 3) == 0 && p0 + 3 < hw) {
            // ONE (h,w) field shared by the whole batch and all channels (degradations.py:593-596)
            const float4 t = normal4(ph, (uint64_t)(p0 >> 2), offset * 8 + STREAM_GRAY);
            g.ng[0] = t.x; g.ng[1] = t.y; g.ng[2] = t.z; g.ng[3] = t.w;
        } else {
            for (int k = 0; k < cnt; ++k) {
                int p = p0 + k; if (p >= hw) p -= hw;
                const float4 t = normal4(ph, (uint64_t)(p >> 2), offset * 8 + STREAM_GRAY);
                const float tt[4] = {t.x, t.y, t.z, t.w};
                g.ng[k] = tt[p & 3];
            }
        }
    }
}

template <bool VEC, bool EXACT>
__global__ void __launch_bounds__(256) gaussian_noise_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                             int chw, int hw,
                                                             const float* __restrict__ sigma, const float* __restrict__ gray,
                                                             const float* __restrict__ ncol, const float* __restrict__ ngray,
                                                             uint64_t seed, uint64_t offset, int flags) {
    const Philox ph(seed);
    const int b = blockIdx.y;
    const float sg = sigma[b];
    const float g = gray ? gray[b] : 0.0f;
    const bool use_gray = gray != nullptr;
    const float one_minus_g = __fsub_rn(1.0f, g);
    const float s255 = __fdiv_rn(sg, 255.0f);
    const float ca = use_gray ? s255 * one_minus_g : s255, cb = use_gray ? s255 * g : 0.0f;  // folded factors
    const int nq = (chw + 3) >> 2;
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
            if (qs[u] < nq) gauss_load<VEC, EXACT>(gq[u], ph, ip, qs[u], nq, chw, hw, base, b, ncol, ngray, need_color, need_gray, offset);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (qs[u] >= nq) continue;
            const int e0 = qs[u] << 2;
            const int cnt = min(4, chw - e0);
            float r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float noise;
                if (EXACT) {
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
    const int b = blockIdx.y;
    const float* ip = img + (size_t)b * 3 * hw;
    uint32_t mc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
        const float r = ip[p], g = ip[hw + p], bl = ip[2 * hw + p];
        const int lv[4] = {level8(r), level8(g), level8(bl), level8(gray_of(r, g, bl))};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t bit = 1u << (lv[k] & 31);
            const int word = lv[k] >> 5;
#pragma unroll
            for (int wd = 0; wd < 8; ++wd) {
                const uint32_t m = word == wd ? bit : 0u;
                if (k < 3) mc[wd] |= m; else mg[wd] |= m;
            }
        }
    }
    __shared__ uint32_t sm[16];
    if (threadIdx.x < 16) sm[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int wd = 0; wd < 8; ++wd) {
        const uint32_t c = __reduce_or_sync(0xffffffffu, mc[wd]);
        const uint32_t g = __reduce_or_sync(0xffffffffu, mg[wd]);
        if ((threadIdx.x & 31) == 0) {
            if (c) atomicOr(&sm[wd], c);
            if (g) atomicOr(&sm[8 + wd], g);
        }
    }
    __syncthreads();
    if (threadIdx.x < 16 && sm[threadIdx.x]) atomicOr(&masks[b * 16 + threadIdx.x], sm[threadIdx.x]);
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
__global__ void __launch_bounds__(256) poisson_apply_kernel(const float* __restrict__ img, float* __restrict__ out, int hw,
                                                            const float* __restrict__ scale, const float* __restrict__ gray,
                                                            const float* __restrict__ counts_c, const float* __restrict__ counts_g,
                                                            uint64_t seed, uint64_t offset, int flags,
                                                            const uint32_t* __restrict__ masks, float* __restrict__ vals_out,
                                                            float* __restrict__ lam_c_out, float* __restrict__ lam_g_out) {
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
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
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
                                      uint64_t seed, uint64_t offset, int flags, float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && sigma_dev, OTF_ERR_BAD_ARG, "gaussian_noise: null pointer");
    OTF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OTF_ERR_BAD_ARG, "gaussian_noise: bad extents");
    OTF_REQUIRE(!(noise_gray_dev && !noise_color_dev), OTF_ERR_BAD_ARG, "gaussian_noise: inject both fields or neither");
    OTF_REQUIRE((int64_t)C * H * W < (1ll << 30), OTF_ERR_UNSUPPORTED, "gaussian_noise: sample too large");
    OTF_REQUIRE(B <= 65535, OTF_ERR_UNSUPPORTED, "gaussian_noise: B > 65535");
    const float* ng = gray_dev ? noise_gray_dev : nullptr;
    OTF_REQUIRE(!(gray_dev && noise_color_dev && !noise_gray_dev), OTF_ERR_BAD_ARG,
                "gaussian_noise: gray flags with an injected colour field need the injected gray field too");
    const int chw = C * H * W;
    const bool vec = (chw % 4 == 0) && (((uintptr_t)img & 15) == 0) && (((uintptr_t)out & 15) == 0);
    int chunks = ceil_div((chw + 3) / 4, 512);
    const int cap = ceil_div(kNumSMs * 16, B);
    if (chunks > cap) chunks = cap;
    const dim3 grid(chunks, B);
    const bool exact = noise_color_dev != nullptr || (flags & OTF_NOISE_ROUNDS);
    cudaStream_t st = (cudaStream_t)stream;
#define OTF_GAUSS(V, E)                                                                                              \
    gaussian_noise_kernel<V, E><<<grid, 256, 0, st>>>(img, out, chw, H * W, sigma_dev, gray_dev, noise_color_dev, ng, \
                                                      seed, offset, flags)
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

extern "C" int otf_poisson_noise_f32(const float* img, int B, int C, int H, int W, const float* scale_dev,
                                     const float* gray_dev, const float* counts_color_dev, const float* counts_gray_dev,
                                     uint64_t seed, uint64_t offset, int flags, uint32_t* masks_dev, float* vals_out_dev,
                                     float* lambda_color_dev, float* lambda_gray_dev, float* out, void* stream) {
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
    poisson_presence_kernel<<<dim3(chunks, B), 256, 0, st>>>(img, hw, masks_dev);
    OTF_LAUNCH_CHECK("poisson_presence_kernel");
    int chunks2 = ceil_div(hw, 256);
    const int max_chunks2 = ceil_div(kNumSMs * 16, B);
    if (chunks2 > max_chunks2) chunks2 = max_chunks2;
    poisson_apply_kernel<<<dim3(chunks2, B), 256, 0, st>>>(img, out, hw, scale_dev, gray_dev, counts_color_dev,
                                                           counts_gray_dev, seed, offset, flags, masks_dev, vals_out_dev,
                                                           lambda_color_dev, lambda_gray_dev);
    OTF_LAUNCH_CHECK("poisson_apply_kernel");
    return OTF_OK;
}

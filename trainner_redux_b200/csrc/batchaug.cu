// MoA batch augment on the finished pair (traiNNer/ops/batchaug.py:21-509).
// The host draws the augmentation, the mixing ratio, the batch permutation and the box in the
// reference's order; the device side is two streaming kernels (plus the resize kernels of resize.cu):
//   mixup_kernel     out[b] = lam * x[b] + (1 - lam) * x[perm[b]]     (batchaug.py:150-158)
//   copy_box_kernel  dst[b, :, dy:dy+bh, dx:dx+bw] = src[perm[b], :, sy:sy+bh, sx:sx+bw]
//                    (cutmix paste :222-227, resizemix paste :318-319, cutblur paste :394-401,
//                     the crops of `up` :476-477)
// The permutation travels in the kernel parameter bank (B <= 512): no H2D copy, no host sync.
#include <string.h>

#include "otf_common.cuh"

namespace otf {

struct Perm {
    int32_t idx[512];
};

// torch evaluates lam * x, (1 - lam) * y and the sum as three fp32 ops: no FMA contraction here.
__device__ __forceinline__ float mix1(float x, float y, float a, float b) {
    return __fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y));
}

template <bool VEC>
__global__ void __launch_bounds__(256) mixup_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n,
                                                    float a, float b, const __grid_constant__ Perm perm) {
    const int s = blockIdx.y;
    const float* xs = x + (int64_t)s * n;
    const float* ys = x + (int64_t)perm.idx[s] * n;
    float* os = out + (int64_t)s * n;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, ts = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        const int64_t nq = n >> 2;
        constexpr int UNR = 2;
        for (int64_t qb = t0; qb < nq; qb += ts * UNR) {
            float4 u[UNR], v[UNR];
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                const int64_t q = qb + k * ts;
                if (q < nq) {
                    u[k] = __ldg(reinterpret_cast<const float4*>(xs) + q);
                    v[k] = __ldg(reinterpret_cast<const float4*>(ys) + q);
                }
            }
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                const int64_t q = qb + k * ts;
                if (q < nq)
                    reinterpret_cast<float4*>(os)[q] = make_float4(mix1(u[k].x, v[k].x, a, b), mix1(u[k].y, v[k].y, a, b),
                                                                   mix1(u[k].z, v[k].z, a, b), mix1(u[k].w, v[k].w, a, b));
            }
        }
    } else {
        for (int64_t i = t0; i < n; i += ts) os[i] = mix1(__ldg(xs + i), __ldg(ys + i), a, b);
    }
}

// One thread moves VEC (4 or 1) consecutive pixels of a box row; blockIdx.y = destination sample.
template <int VEC>
__global__ void __launch_bounds__(256) copy_box_kernel(const float* __restrict__ src, int Hs, int Ws, int sy, int sx,
                                                       float* __restrict__ dst, int Hd, int Wd, int dy, int dx, int bh,
                                                       int bw, int planes, int use_perm,
                                                       const __grid_constant__ Perm perm) {
    const int s = blockIdx.y;
    const int from = use_perm ? perm.idx[s] : s;
    const float* sp = src + (int64_t)from * planes * Hs * Ws;
    float* dp = dst + (int64_t)s * planes * Hd * Wd;
    const int qrow = bw / VEC;
    const int64_t n = (int64_t)planes * bh * qrow;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int xq = (int)(i % qrow);
        const int64_t t = i / qrow;
        const int y = (int)(t % bh);
        const int64_t pl = t / bh;
        const float* a = sp + (pl * Hs + sy + y) * Ws + sx + VEC * xq;
        float* o = dp + (pl * Hd + dy + y) * Wd + dx + VEC * xq;
        if (VEC == 4) *reinterpret_cast<float4*>(o) = __ldg(reinterpret_cast<const float4*>(a));
        else *o = __ldg(a);
    }
}

static int fill_perm(Perm& p, const int32_t* perm_host, int B, const char* who) {
    memset(&p, 0, sizeof(p));
    if (!perm_host) return OTF_OK;
    for (int i = 0; i < B; ++i) {
        OTF_REQUIRE(perm_host[i] >= 0 && perm_host[i] < B, OTF_ERR_BAD_ARG, "%s: perm[%d] = %d outside [0, %d)", who, i, perm_host[i], B);
        p.idx[i] = perm_host[i];
    }
    return OTF_OK;
}

}  // namespace otf

extern "C" int otf_mixup_f32(const float* img, const int32_t* perm_host, int B, int64_t sample_elems, float lam,
                             float one_minus_lam, float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && out && perm_host, OTF_ERR_BAD_ARG, "mixup: null pointer");
    OTF_REQUIRE(img != out, OTF_ERR_BAD_ARG, "mixup: in place is not allowed (samples are read through the permutation)");
    OTF_REQUIRE(B > 0 && B <= 512 && sample_elems > 0, OTF_ERR_BAD_ARG, "mixup: need 0 < B <= 512 (got %d)", B);
    Perm p;
    if (int rc = fill_perm(p, perm_host, B, "mixup")) return rc;
    const bool vec = (sample_elems % 4 == 0) && ((((uintptr_t)img | (uintptr_t)out) & 15) == 0);
    int bx = ceil_div(vec ? sample_elems / 4 : sample_elems, 256 * 2);
    const int cap = ceil_div((int64_t)kNumSMs * 8, B);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    if (vec) mixup_kernel<true><<<dim3(bx, B), 256, 0, (cudaStream_t)stream>>>(img, out, sample_elems, lam, one_minus_lam, p);
    else mixup_kernel<false><<<dim3(bx, B), 256, 0, (cudaStream_t)stream>>>(img, out, sample_elems, lam, one_minus_lam, p);
    OTF_LAUNCH_CHECK("mixup_kernel");
    return OTF_OK;
}

extern "C" int otf_copy_box_f32(const float* src, int Hs, int Ws, int sy, int sx, float* dst, int Hd, int Wd, int dy, int dx,
                                int bh, int bw, int B, int planes_per_sample, const int32_t* perm_host, void* stream) {
    using namespace otf;
    OTF_REQUIRE(src && dst, OTF_ERR_BAD_ARG, "copy_box: null pointer");
    OTF_REQUIRE(B > 0 && B <= 512 && planes_per_sample > 0, OTF_ERR_BAD_ARG, "copy_box: need 0 < B <= 512 (got %d)", B);
    OTF_REQUIRE(bh >= 0 && bw >= 0 && sy >= 0 && sx >= 0 && dy >= 0 && dx >= 0 && sy + bh <= Hs && sx + bw <= Ws &&
                    dy + bh <= Hd && dx + bw <= Wd,
                OTF_ERR_BAD_ARG, "copy_box: box (%d, %d) at src (%d, %d) / dst (%d, %d) outside (%d, %d) / (%d, %d)", bh, bw, sy,
                sx, dy, dx, Hs, Ws, Hd, Wd);
    if (bh == 0 || bw == 0) return OTF_OK;  // an empty slice assignment is a no-op in the reference too
    OTF_REQUIRE(src != dst || !perm_host, OTF_ERR_BAD_ARG, "copy_box: a permuted copy cannot run in place");
    Perm p;
    if (int rc = fill_perm(p, perm_host, B, "copy_box")) return rc;
    const bool vec = (bw % 4 == 0) && (Ws % 4 == 0) && (Wd % 4 == 0) && (sx % 4 == 0) && (dx % 4 == 0) &&
                     ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0);
    const int64_t n = (int64_t)planes_per_sample * bh * (vec ? bw / 4 : bw);
    int bx = ceil_div(n, 256);
    const int cap = ceil_div((int64_t)kNumSMs * 8, B);
    if (bx > cap) bx = cap;
    if (vec) copy_box_kernel<4><<<dim3(bx, B), 256, 0, (cudaStream_t)stream>>>(src, Hs, Ws, sy, sx, dst, Hd, Wd, dy, dx, bh, bw, planes_per_sample, perm_host != nullptr, p);
    else copy_box_kernel<1><<<dim3(bx, B), 256, 0, (cudaStream_t)stream>>>(src, Hs, Ws, sy, sx, dst, Hd, Wd, dy, dx, bh, bw, planes_per_sample, perm_host != nullptr, p);
    OTF_LAUNCH_CHECK("copy_box_kernel");
    return OTF_OK;
}

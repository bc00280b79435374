// Small streaming kernels of the path: 8-bit lattice clamp/round, paired crop, layout
// normalisation, pair-pool slot gather/scatter; plus the library's error plumbing.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "otf_common.cuh"

namespace otf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool pdl_enabled() {
    // off by default: measured on B200 (profiles/r02_pdl_ab.json) the early-resident dependents cost more than the launch
    // gaps they hide — single-stream feed_data 227 k pairs/s with the attribute vs 248 k without, and the prefetcher loop
    // (copy stream + events around the captured chain) collapses; OTF_PDL=1 turns the attribute on for further study
    static const bool on = [] { const char* e = getenv("OTF_PDL"); return e && e[0] == '1'; }();
    return on;
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return OTF_ERR_CUDA;
}

// ---- a7: traiNNer/models/realesrgan_model.py:616 ------------------------------------------
__global__ void __launch_bounds__(256) clamp_round_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
    pdl_enter();
    const int64_t nq = n >> 2;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[q];
        reinterpret_cast<float4*>(out)[q] = make_float4(quantise8(v.x), quantise8(v.y), quantise8(v.z), quantise8(v.w));
    }
    const int64_t tail = (nq << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tail < n) out[tail] = quantise8(x[tail]);
}

// ---- uint8 GT upload: x / 255 on the device ------------------------------
// The reference normalises on the host (traiNNer/utils/img_util.py:65-109 `img2tensor`: float32(img) / 255) and
// ships fp32 over PCIe; uploading the 8-bit image and dividing here moves 4x fewer bytes.  Same IEEE division.
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
    pdl_enter();
    const int64_t nq = n >> 2;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        const uchar4 u = reinterpret_cast<const uchar4*>(src)[q];
        reinterpret_cast<float4*>(dst)[q] = make_float4(__fdiv_rn((float)u.x, 255.0f), __fdiv_rn((float)u.y, 255.0f),
                                                        __fdiv_rn((float)u.z, 255.0f), __fdiv_rn((float)u.w, 255.0f));
    }
    const int64_t tail = (nq << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tail < n) dst[tail] = __fdiv_rn((float)src[tail], 255.0f);
}

// ---- the way back: an image on the 8-bit lattice as bytes ------------------------------------
// What traiNNer/utils/img_util.py:112-181 (`tensor2img`: `(img * 255.0).round()` -> uint8) does on the host after a
// full-precision read-back: clamp(round(x * 255), 0, 255) on the device, so that a finished LQ batch (which lies on that
// lattice already, realesrgan_model.py:616) crosses PCIe as one byte per value and `u8.float() / 255` restores it bit for bit.
__global__ void __launch_bounds__(256) f32_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t n) {
    pdl_enter();
    auto level = [](float v) { return (unsigned)fminf(fmaxf(rintf(__fmul_rn(v, 255.0f)), 0.0f), 255.0f); };
    const int64_t nq = n >> 2;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
        reinterpret_cast<uchar4*>(dst)[q] = make_uchar4(level(v.x), level(v.y), level(v.z), level(v.w));
    }
    const int64_t tail = (nq << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tail < n) dst[tail] = (uint8_t)level(src[tail]);
}

// (copy_window, the quad-per-thread window copy of a8, lives in otf_common.cuh: the fused DiffJPEG + crop launch uses it too)
template <bool VEC_GT, bool VEC_LQ, bool ROUND8>
__global__ void __launch_bounds__(256) crop_pair_kernel(const float* __restrict__ gt, int Hg, int Wg,
                                                        const float* __restrict__ lq, int Hl, int Wl, int top, int left,
                                                        const int32_t* __restrict__ top_left_dev,
                                                        int p, int scale, int planes, float* __restrict__ gt_out,
                                                        float* __restrict__ lq_out) {
    pdl_enter();
    if (top_left_dev) {  // per-step offsets of a captured chain, clamped so that a bad upload cannot read outside the image
        top = clampi(top_left_dev[0], 0, Hl - p);
        left = clampi(top_left_dev[1], 0, Wl - p);
    }
    const int64_t q0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, qs = (int64_t)gridDim.x * blockDim.x;
    if (gt_out) copy_window<VEC_GT>(gt, Hg, Wg, top * scale, left * scale, p * scale, gt_out, planes, q0, qs);  // (NULL: the GT window stays a view)
    copy_window<VEC_LQ, ROUND8>(lq, Hl, Wl, top, left, p, lq_out, planes, q0, qs);
}

// any patch size (p % 4 != 0): one element per thread
__global__ void __launch_bounds__(256) crop_pair_scalar_kernel(const float* __restrict__ gt, int Hg, int Wg,
                                                               const float* __restrict__ lq, int Hl, int Wl, int top,
                                                               int left, const int32_t* __restrict__ top_left_dev, int p,
                                                               int scale, int planes, int round8,
                                                               float* __restrict__ gt_out, float* __restrict__ lq_out) {
    pdl_enter();
    if (top_left_dev) {
        top = clampi(top_left_dev[0], 0, Hl - p);
        left = clampi(top_left_dev[1], 0, Wl - p);
    }
    const int G = p * scale;
    const int64_t ng = gt_out ? (int64_t)planes * G * G : 0, nl = (int64_t)planes * p * p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ng + nl; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < ng) {
            const int x = (int)(i % G), y = (int)((i / G) % G);
            const int64_t pl = i / ((int64_t)G * G);
            gt_out[i] = __ldg(gt + ((size_t)pl * Hg + top * scale + y) * Wg + left * scale + x);
        } else {
            const int64_t j = i - ng;
            const int x = (int)(j % p), y = (int)((j / p) % p);
            const int64_t pl = j / ((int64_t)p * p);
            const float v = __ldg(lq + ((size_t)pl * Hl + top + y) * Wl + left + x);
            lq_out[j] = round8 ? quantise8(v) : v;
        }
    }
}

// ---- strided (channels_last, sliced views...) -> dense NCHW -------------------------------
__global__ void __launch_bounds__(256) copy_strided_kernel(const float* __restrict__ src, int64_t sb, int64_t sc, int64_t sh,
                                                           int64_t sw, int C, int H, int W, float* __restrict__ dst,
                                                           int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        int64_t t = i / W;
        const int y = (int)(t % H);
        t /= H;
        const int c = (int)(t % C);
        const int64_t b = t / C;
        dst[i] = __ldg(src + b * sb + c * sc + y * sh + x * sw);
    }
}

// ---- pair-pool slot movement ---------------------------------------------
struct SlotIdx {
    int32_t idx[512];
};
template <bool SCATTER>
__global__ void __launch_bounds__(256) move_slots_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                         int64_t slot_elems, const __grid_constant__ SlotIdx map) {
    const int s = blockIdx.y;
    const float* sp = src + (SCATTER ? (int64_t)s : (int64_t)map.idx[s]) * slot_elems;
    float* dp = dst + (SCATTER ? (int64_t)map.idx[s] : (int64_t)s) * slot_elems;
    const bool vec = (slot_elems % 4 == 0) && ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0);
    if (vec) {
        const int64_t nq = slot_elems >> 2;
        for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x)
            reinterpret_cast<float4*>(dp)[q] = __ldg(reinterpret_cast<const float4*>(sp) + q);
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slot_elems; i += (int64_t)gridDim.x * blockDim.x)
            dp[i] = __ldg(sp + i);
    }
}

template <bool SCATTER>
static int move_slots(const float* src, const int32_t* idx_host, int n, int64_t slot_elems, float* dst, void* stream) {
    OTF_REQUIRE(src && dst && idx_host, OTF_ERR_BAD_ARG, "move_slots: null pointer");
    OTF_REQUIRE(n > 0 && n <= 512 && slot_elems > 0, OTF_ERR_BAD_ARG, "move_slots: need 0 < n <= 512 (got %d)", n);
    SlotIdx m;
    memset(&m, 0, sizeof(m));
    for (int i = 0; i < n; ++i) {
        OTF_REQUIRE(idx_host[i] >= 0, OTF_ERR_BAD_ARG, "move_slots: negative slot index");
        m.idx[i] = idx_host[i];
    }
    int bx = ceil_div(slot_elems / 4 + 1, 256 * 4);
    if (bx > 64) bx = 64;
    move_slots_kernel<SCATTER><<<dim3(bx, n), 256, 0, (cudaStream_t)stream>>>(src, dst, slot_elems, m);
    OTF_LAUNCH_CHECK("move_slots_kernel");
    return OTF_OK;
}

// One pool step in ONE launch (realesrgan_model.py:430-447): for every listed slot of both queues, hand the stored
// pair out (dequeue; skipped while the pool is still filling) and store the new pair in its place (enqueue).
// grid = (chunks, n, 2): z selects the LQ or the GT queue.
__global__ void __launch_bounds__(256) pool_exchange_kernel(float* __restrict__ q_lq, float* __restrict__ q_gt,
                                                            const float* __restrict__ in_lq, const float* __restrict__ in_gt,
                                                            float* __restrict__ out_lq, float* __restrict__ out_gt,
                                                            int64_t lq_elems, int64_t gt_elems, const __grid_constant__ SlotIdx map) {
    const int s = blockIdx.y;
    const bool gt = blockIdx.z != 0;
    const int64_t n = gt ? gt_elems : lq_elems;
    float* qp = (gt ? q_gt : q_lq) + (int64_t)map.idx[s] * n;
    const float* ip = (gt ? in_gt : in_lq) + (int64_t)s * n;
    float* op = gt ? out_gt : out_lq;
    if (op) op += (int64_t)s * n;
    const bool vec = (n % 4 == 0) && ((((uintptr_t)qp | (uintptr_t)ip | (uintptr_t)op) & 15) == 0);
    if (vec) {
        const int64_t nq = n >> 2;
        for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
            const float4 fresh = __ldg(reinterpret_cast<const float4*>(ip) + q);
            if (op) reinterpret_cast<float4*>(op)[q] = reinterpret_cast<const float4*>(qp)[q];
            reinterpret_cast<float4*>(qp)[q] = fresh;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const float fresh = __ldg(ip + i);
            if (op) op[i] = qp[i];
            qp[i] = fresh;
        }
    }
}

}  // namespace otf

extern "C" int otf_pool_exchange_f32(float* queue_lq, float* queue_gt, const int32_t* idx_host, int n, int64_t lq_elems,
                                     int64_t gt_elems, const float* in_lq, const float* in_gt, float* out_lq, float* out_gt,
                                     void* stream) {
    using namespace otf;
    OTF_REQUIRE(queue_lq && queue_gt && idx_host && in_lq && in_gt, OTF_ERR_BAD_ARG, "pool_exchange: null pointer");
    OTF_REQUIRE((out_lq == nullptr) == (out_gt == nullptr), OTF_ERR_BAD_ARG, "pool_exchange: give both outputs or neither");
    OTF_REQUIRE(n > 0 && n <= 512 && lq_elems > 0 && gt_elems > 0, OTF_ERR_BAD_ARG, "pool_exchange: need 0 < n <= 512 (got %d)", n);
    SlotIdx m;
    memset(&m, 0, sizeof(m));
    for (int i = 0; i < n; ++i) {
        OTF_REQUIRE(idx_host[i] >= 0, OTF_ERR_BAD_ARG, "pool_exchange: negative slot index");
        m.idx[i] = idx_host[i];
    }
    const int64_t big = gt_elems > lq_elems ? gt_elems : lq_elems;
    int bx = ceil_div(big / 4 + 1, 256 * 4);
    if (bx > 64) bx = 64;
    pool_exchange_kernel<<<dim3(bx, n, 2), 256, 0, (cudaStream_t)stream>>>(queue_lq, queue_gt, in_lq, in_gt, out_lq, out_gt, lq_elems,
                                                                        gt_elems, m);
    OTF_LAUNCH_CHECK("pool_exchange_kernel");
    return OTF_OK;
}

extern "C" int otf_abi_version(void) { return OTF_ABI_VERSION; }
extern "C" const char* otf_last_error(void) { return otf::g_err; }
extern "C" int otf_device_cc(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return otf::cuda_fail(cudaGetLastError(), "cudaGetDevice");
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return major * 10 + minor;
}

extern "C" int otf_clamp_round_f32(const float* x, int64_t n, float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(x && out && n > 0, OTF_ERR_BAD_ARG, "clamp_round: bad args");
    OTF_REQUIRE(((((uintptr_t)x) | ((uintptr_t)out)) & 15) == 0, OTF_ERR_BAD_ARG, "clamp_round: pointers must be 16-byte aligned");
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    if (blocks < 1) blocks = 1;
    launch_chain(clamp_round_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, x, out, n);
    OTF_LAUNCH_CHECK("clamp_round_kernel");
    return OTF_OK;
}

extern "C" int otf_u8_to_f32(const uint8_t* src, int64_t n, float* dst, void* stream) {
    using namespace otf;
    OTF_REQUIRE(src && dst && n > 0, OTF_ERR_BAD_ARG, "u8_to_f32: bad args");
    OTF_REQUIRE((((uintptr_t)src) & 3) == 0 && (((uintptr_t)dst) & 15) == 0, OTF_ERR_BAD_ARG, "u8_to_f32: misaligned pointers");
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    if (blocks < 1) blocks = 1;
    launch_chain(u8_to_f32_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, src, dst, n);
    OTF_LAUNCH_CHECK("u8_to_f32_kernel");
    return OTF_OK;
}

extern "C" int otf_f32_to_u8(const float* src, int64_t n, uint8_t* dst, void* stream) {
    using namespace otf;
    OTF_REQUIRE(src && dst && n > 0, OTF_ERR_BAD_ARG, "f32_to_u8: bad args");
    OTF_REQUIRE((((uintptr_t)src) & 15) == 0 && (((uintptr_t)dst) & 3) == 0, OTF_ERR_BAD_ARG, "f32_to_u8: misaligned pointers");
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    if (blocks < 1) blocks = 1;
    launch_chain(f32_to_u8_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, src, dst, n);
    OTF_LAUNCH_CHECK("f32_to_u8_kernel");
    return OTF_OK;
}

extern "C" int otf_crop_pair_f32(const float* gt, int planes, int Hg, int Wg, const float* lq, int Hl, int Wl, int top,
                                 int left, const int32_t* top_left_dev, int lq_patch, int scale, int lq_round8, float* gt_out,
                                 float* lq_out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(lq && lq_out && (gt || !gt_out), OTF_ERR_BAD_ARG, "crop_pair: null pointer");
    OTF_REQUIRE(planes > 0 && planes <= 65535 && scale > 0 && lq_patch > 0, OTF_ERR_BAD_ARG, "crop_pair: bad extents");
    OTF_REQUIRE(Hg == Hl * scale && Wg == Wl * scale, OTF_ERR_BAD_ARG, "crop_pair: GT (%d, %d) is not %dx LQ (%d, %d)", Hg, Wg, scale, Hl, Wl);
    OTF_REQUIRE(top >= 0 && left >= 0 && top + lq_patch <= Hl && left + lq_patch <= Wl, OTF_ERR_BAD_ARG, "crop_pair: window outside LQ");
    const int G = lq_patch * scale;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t work = ((gt_out ? (int64_t)planes * G * G : 0) + (int64_t)planes * lq_patch * lq_patch) / 4;
    int blocks = (int)((work / 4 + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    const bool quads = (lq_patch % 4 == 0) && ((((uintptr_t)gt_out | (uintptr_t)lq_out) & 15) == 0);
    if (!quads) {
        launch_chain(crop_pair_scalar_kernel, dim3(blocks), dim3(256), 0, st, gt, Hg, Wg, lq, Hl, Wl, top, left, top_left_dev, lq_patch, scale, planes, lq_round8, gt_out, lq_out);
        OTF_LAUNCH_CHECK("crop_pair_scalar_kernel");
        return OTF_OK;
    }
    // device-side offsets: the alignment of the window start is only known for the GT window at scale % 4 == 0
    const bool vg = (Wg % 4 == 0) && (top_left_dev ? scale % 4 == 0 : (left * scale) % 4 == 0) && (((uintptr_t)gt & 15) == 0);
    const bool vl = (Wl % 4 == 0) && !top_left_dev && (left % 4 == 0) && (((uintptr_t)lq & 15) == 0);
#define OTF_CROP(VG, VL, R8) \
    launch_chain(crop_pair_kernel<VG, VL, R8>, dim3(blocks), dim3(256), 0, st, gt, Hg, Wg, lq, Hl, Wl, top, left, top_left_dev, lq_patch, scale, planes, gt_out, lq_out)
    if (lq_round8) {
        if (vg && vl) OTF_CROP(true, true, true); else if (vg) OTF_CROP(true, false, true);
        else if (vl) OTF_CROP(false, true, true); else OTF_CROP(false, false, true);
    } else {
        if (vg && vl) OTF_CROP(true, true, false); else if (vg) OTF_CROP(true, false, false);
        else if (vl) OTF_CROP(false, true, false); else OTF_CROP(false, false, false);
    }
#undef OTF_CROP
    OTF_LAUNCH_CHECK("crop_pair_kernel");
    return OTF_OK;
}

extern "C" int otf_copy_strided_f32(const float* src, const int64_t strides[4], int B, int C, int H, int W, float* dst,
                                    void* stream) {
    using namespace otf;
    OTF_REQUIRE(src && dst && strides, OTF_ERR_BAD_ARG, "copy_strided: null pointer");
    OTF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OTF_ERR_BAD_ARG, "copy_strided: bad extents");
    const int64_t n = (int64_t)B * C * H * W;
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    copy_strided_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, strides[0], strides[1], strides[2], strides[3], C, H, W, dst, n);
    OTF_LAUNCH_CHECK("copy_strided_kernel");
    return OTF_OK;
}

extern "C" int otf_gather_slots_f32(const float* src, const int32_t* idx_host, int n, int64_t slot_elems, float* dst,
                                    void* stream) {
    return otf::move_slots<false>(src, idx_host, n, slot_elems, dst, stream);
}
extern "C" int otf_scatter_slots_f32(const float* src, const int32_t* idx_host, int n, int64_t slot_elems, float* dst,
                                     void* stream) {
    return otf::move_slots<true>(src, idx_host, n, slot_elems, dst, stream);
}

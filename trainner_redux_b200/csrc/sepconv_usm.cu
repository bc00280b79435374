// 1-D reflect-padded correlation (horizontal / vertical) and the fused USM sharpen.
// Replaces traiNNer/utils/img_process_util.py:35-55 (USMSharp: two 51x51 filter2d calls +
// ~8 elementwise launches) and the Lanczos prefilter of traiNNer/data/degradations.py:982-998.
//
// The USM Gaussian is an exact outer product (cv2.getGaussianKernel(r) x itself), so the
// 2601-tap 2-D correlation is evaluated as 51 + 51 taps.  HBM-bound: every pass streams the
// plane once; taps ride in the kernel parameter bank (uniform loads, no LSU traffic).
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
__global__ void __launch_bounds__(256) sepconv_h_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W,
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
__global__ void __launch_bounds__(256) sepconv_v_kernel(const float* __restrict__ tmp, float* __restrict__ out, int H, int W,
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

static int launch_h(const float* img, int planes, int H, int W, const Taps& t, float* out, cudaStream_t st) {
    const dim3 grid(ceil_div(W, H_TW), ceil_div(H, H_ROWS), planes);
    const size_t smem = (size_t)H_ROWS * (H_TW + 2 * (t.n / 2)) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sepconv_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "sepconv_h smem attribute");
    }
    sepconv_h_kernel<<<grid, 256, smem, st>>>(img, out, H, W, t);
    OTF_LAUNCH_CHECK("sepconv_h_kernel");
    return OTF_OK;
}

template <int EPI>
static int launch_v(const float* tmp, int planes, int H, int W, const Taps& t, float* out, const float* img, float* aux,
                    float weight, float threshold, cudaStream_t st) {
    const dim3 grid(ceil_div(W, V_TW), ceil_div(H, V_TH), planes);
    const size_t smem = (size_t)(V_TH + 2 * (t.n / 2)) * V_TW * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sepconv_v_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "sepconv_v smem attribute");
    }
    sepconv_v_kernel<EPI><<<grid, 256, smem, st>>>(tmp, out, H, W, t, img, aux, weight, threshold);
    OTF_LAUNCH_CHECK("sepconv_v_kernel");
    return OTF_OK;
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
    const size_t n = (size_t)planes * H * W;
    float* tmp = (float*)workspace_dev;
    float* mask = tmp + n;
    float* sharp = mask + n;
    int rc;
    if ((rc = launch_h(img, planes, H, W, t, tmp, st))) return rc;                                       // blur, rows
    if ((rc = launch_v<EPI_USM_MASK>(tmp, planes, H, W, t, mask, img, sharp, weight, threshold, st))) return rc;  // blur, cols + mask/sharp
    if ((rc = launch_h(mask, planes, H, W, t, tmp, st))) return rc;                                      // soft mask, rows
    return launch_v<EPI_USM_BLEND>(tmp, planes, H, W, t, out, img, sharp, weight, threshold, st);       // soft mask, cols + blend
}

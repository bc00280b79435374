// Shared device/host helpers for libotf_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/otf_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libotf_b200 is written for sm_100a (B200) only"
#endif

namespace otf {

// ---------------------------------------------------------------- errors ----
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define OTF_REQUIRE(cond, code, ...)          \
    do {                                      \
        if (!(cond)) {                        \
            ::otf::set_error(__VA_ARGS__);    \
            return (code);                    \
        }                                     \
    } while (0)

#define OTF_LAUNCH_CHECK(what)                                        \
    do {                                                              \
        cudaError_t e__ = cudaGetLastError();                         \
        if (e__ != cudaSuccess) return ::otf::cuda_fail(e__, what);   \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized against this

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------ programmatic dependent launch ----
// The chain's kernels are launched back to back on one stream (or as consecutive nodes of a captured graph).  With
// programmatic stream serialisation the NEXT kernel's CTAs may become resident while the last wave of this one drains:
// every kernel on the chain starts with pdl_enter() — wait for the predecessor grid to complete (nothing of its output
// is touched before, so plain stream order is preserved, ping-pong buffers included), then allow the successor's launch.
// What is saved is the launch latency and CTA ramp-up at each of the ~8 kernel boundaries of a step (single stream:
// ~4 us each).  The attribute is only attached with OTF_PDL=1 in the environment (see pdl_enabled(): it measured
// slower); without it griddepcontrol.* are no-ops.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// --------------------------------------------------------------- indexing ---
// torch F.pad(mode="reflect"): mirror without repeating the edge sample.
__host__ __device__ __forceinline__ int reflect_idx(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}
__host__ __device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ------------------------------------------------------- exact fp32 pieces --
// The reference evaluates these as separate ATen ops (no FMA contraction), and a
// rounding cliff follows them, so they are written with non-contractable
// intrinsics to stay bit-identical.
__device__ __forceinline__ float quantise8(float x) {
    // clamp(round(x*255),0,255)/255 — realesrgan_model.py:616, degradations.py:789
    float v = rintf(__fmul_rn(x, 255.0f));
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    return __fdiv_rn(v, 255.0f);
}
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

__device__ __forceinline__ float noise_tail(float v, int flags) {
    // degradations.py:626-632
    const bool clip = flags & OTF_NOISE_CLIP, rounds = flags & OTF_NOISE_ROUNDS;
    if (clip && rounds) return quantise8(v);
    if (clip) return clamp01(v);
    if (rounds) return __fdiv_rn(rintf(__fmul_rn(v, 255.0f)), 255.0f);
    return v;
}

// x / 255 with IEEE rounding in three FP32-pipe operations instead of the ~10 of a general division: q = x * RN(1/255),
// one exact residual, one correction (Markstein).  Equal to __fdiv_rn(x, 255.0f) for EVERY float in [0, 255] — checked
// exhaustively on the device (profiles/experiments/div255_exact.cu) — which is the whole range the codecs divide.
__device__ __forceinline__ float div255(float x) {
    const float y = 3.9215688593685627e-03f;  // RN(1 / 255) = 0x3B808081
    const float q = __fmul_rn(x, y);
    return fmaf(fmaf(-q, 255.0f, x), y, q);
}

// ------------------------------------------------------------ Philox4x32-10 -
struct Philox {
    uint32_t key[2];
    __device__ __forceinline__ Philox(uint64_t seed) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
    }
    // counter = (lo64 = index, hi64 = stream/offset)
    __device__ __forceinline__ uint4 operator()(uint64_t index, uint64_t stream) const {
        uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32);
        uint32_t c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// (0,1] uniform from 32 random bits (never 0, so log() is finite)
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }

// Box-Muller: two standard normals from two 32-bit words
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    // MUFU-based log/sin/cos: abs error ~1e-6 on the normal deviates, far below what the
    // distribution tests (KS at 4M samples) or an 8-bit image can resolve
    const float u1 = u01(a), u2 = u01(b);
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    return make_float2(r * c, r * s);
}
// four standard normals for element-quad `quad` of a stream
__device__ __forceinline__ float4 normal4(const Philox& ph, uint64_t quad, uint64_t stream) {
    const uint4 r = ph(quad, stream);
    const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
    return make_float4(a.x, a.y, b.x, b.y);
}

// ---- a8: traiNNer/data/transforms.py:124-135 followed by .contiguous() --------------------
// One launch copies both windows.  A thread moves one quad (4 consecutive output pixels) per
// iteration and keeps UNR of them in flight; 16-byte loads when the window start is aligned.
template <bool VEC, bool ROUND8 = false>
__device__ __forceinline__ void copy_window(const float* __restrict__ src, int Hs, int Ws, int top, int left, int n,
                                            float* __restrict__ dst, int planes, int64_t q0, int64_t qstride) {
    const int qrow = n >> 2;                          // quads per output row (n % 4 == 0 on this path)
    const int64_t nq = (int64_t)planes * n * qrow;
    constexpr int UNR = 4;
    for (int64_t qb = q0; qb < nq; qb += qstride * UNR) {
        float4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t q = qb + u * qstride;
            if (q < nq) {
                const int xq = (int)(q % qrow);
                const int64_t t = q / qrow;
                const int y = (int)(t % n);
                const int64_t pl = t / n;
                const float* sp = src + ((size_t)pl * Hs + (top + y)) * Ws + left + 4 * xq;
                if (VEC) v[u] = __ldg(reinterpret_cast<const float4*>(sp));
                else v[u] = make_float4(__ldg(sp), __ldg(sp + 1), __ldg(sp + 2), __ldg(sp + 3));
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t q = qb + u * qstride;
            if (q < nq) {
                if (ROUND8) v[u] = make_float4(quantise8(v[u].x), quantise8(v[u].y), quantise8(v[u].z), quantise8(v[u].w));  // a7 fused
                reinterpret_cast<float4*>(dst)[q] = v[u];
            }
        }
    }
}

// streams used by the noise kernels (hi 64 bits of the Philox counter = offset*8 + id)
enum { STREAM_COLOR = 0, STREAM_GRAY = 1, STREAM_POIS_COLOR = 2, STREAM_POIS_GRAY = 3 };

}  // namespace otf

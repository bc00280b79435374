// filter2d: per-sample KxK cross-correlation with reflect padding.
// Replaces traiNNer/utils/img_process_util.py:8-32 (F.pad reflect + grouped F.conv2d).
//
// Roofline note: 2*K*K flop per 8 algorithmic bytes — FP32-FMA
// bound for K > ~9.  The kernel is therefore built around FFMA issue efficiency:
//   * a halo tile of the plane is staged once in shared memory (reflect resolved at
//     load time), the per-sample taps beside it;
//   * each thread owns a TY x TX register block of outputs; an image row of the tile
//     is loaded once into registers (LDS.128) and reused for every output row and tap
//     it contributes to; taps arrive as warp-broadcast LDS.128 (1 wavefront);
//   * the loops are specialised on the kernel's TRUE support (zero-padded 21x21
//     kernels of true size 7..21 are the norm: realesrgan_dataset.py:171-172), found
//     on the device so the host never synchronises.
// Summation order per output: kernel rows ascending, taps left to right, one FFMA each.
#include "otf_common.cuh"

namespace otf {

// ---- true support of each kernel -----------------------------------------------------
__global__ void kernel_support_kernel(const float* __restrict__ kern, int K, int32_t* __restrict__ support) {
    const int kb = blockIdx.x, c = K / 2;
    int r = 0;
    for (int idx = threadIdx.x; idx < K * K; idx += blockDim.x) {
        if (kern[(size_t)kb * K * K + idx] != 0.0f) {
            const int i = idx / K, j = idx - i * K;
            r = max(r, max(abs(i - c), abs(j - c)));
        }
    }
    r = __reduce_max_sync(0xffffffffu, r);
    __shared__ int smax;
    if (threadIdx.x == 0) smax = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax(&smax, r);
    __syncthreads();
    if (threadIdx.x == 0) support[kb] = smax;
}

constexpr int kMaxRT = 10;  // register-blocked path covers radius <= 10 (K <= 21)
constexpr int kWPitch = 24; // taps row pitch in smem (float4 broadcast loads)

template <int TX, int TY, int KT>
__device__ __forceinline__ void accumulate_rows(const float* __restrict__ tile_thread, int pitch,
                                                const float* __restrict__ wsm, float (&acc)[TY][TX]) {
    constexpr int NROW = TX + KT - 1;  // multiple of 4 for TX in {4,8}, KT in {5,9,13,17,21}
    static_assert(NROW % 4 == 0, "row window must be float4 sized");
#pragma unroll 1
    for (int r = 0; r < TY + KT - 1; ++r) {
        float row[NROW];
        const float4* rp = reinterpret_cast<const float4*>(tile_thread + r * pitch);
#pragma unroll
        for (int q = 0; q < NROW / 4; ++q) {
            const float4 v = rp[q];
            row[4 * q + 0] = v.x; row[4 * q + 1] = v.y; row[4 * q + 2] = v.z; row[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int oy = 0; oy < TY; ++oy) {
            const int i = r - oy;  // kernel row feeding output row oy from image row r
            if (i >= 0 && i < KT) {
                const float4* wp = reinterpret_cast<const float4*>(wsm + i * kWPitch);
#pragma unroll
                for (int q = 0; q < (KT + 3) / 4; ++q) {
                    const float4 w4 = wp[q];
                    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int j = 4 * q + t;
                        if (j < KT) {
#pragma unroll
                            for (int ox = 0; ox < TX; ++ox) acc[oy][ox] = fmaf(w[t], row[ox + j], acc[oy][ox]);
                        }
                    }
                }
            }
        }
    }
}

template <int TX, int TY, int BX, int BY>
__global__ void __launch_bounds__(BX* BY) filter2d_kernel(const float* __restrict__ img, const float* __restrict__ kern,
                                                          const int32_t* __restrict__ support, float* __restrict__ out,
                                                          int C, int H, int W, int K, int kernel_batch, int vec_ok) {
    constexpr int TILE_W = TX * BX, TILE_H = TY * BY, NT = BX * BY;
    constexpr int P = TILE_W + 2 * kMaxRT;  // smem row pitch (multiple of 4)
    static_assert(P % 4 == 0, "pitch");
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;                                  // (TILE_H + 2*RT) x P
    float* wsm = smem + (TILE_H + 2 * kMaxRT) * P;       // 21 x kWPitch

    const int plane = blockIdx.z;
    const int b = plane / C;
    const int kb = kernel_batch == 1 ? 0 : b;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.x;
    const int rs = support ? min(support[kb], K / 2) : K / 2;  // true radius (block-uniform)
    // radius variant: 0 (pulse), 2, 4, 6, 8, 10
    const int RT = rs == 0 ? 0 : ((rs + 1) & ~1);
    const int KT = 2 * RT + 1;
    const int c = K / 2;

    // taps -> smem, cropped/padded to the KT x KT centre
    const float* kp = kern + (size_t)kb * K * K;
    for (int idx = tid; idx < 21 * kWPitch; idx += NT) {
        const int i = idx / kWPitch, j = idx - i * kWPitch;
        const int si = c - RT + i, sj = c - RT + j;
        float v = 0.0f;
        if (i < KT && j < KT && si >= 0 && si < K && sj >= 0 && sj < K) v = kp[si * K + sj];
        wsm[idx] = v;
    }
    // halo tile -> smem, reflect resolved here (one warp per row, lanes along x)
    const float* ip = img + (size_t)plane * H * W;
    const int th = TILE_H + 2 * RT, tw = TILE_W + 2 * RT;
    for (int yy = tid >> 5; yy < th; yy += NT / 32) {
        const int gy = clampi(reflect_idx(y0 - RT + yy, H), 0, H - 1);
        const float* rowp = ip + (size_t)gy * W;
        for (int xx = tid & 31; xx < tw; xx += 32) {
            const int gx = clampi(reflect_idx(x0 - RT + xx, W), 0, W - 1);
            tile[yy * P + xx] = __ldg(rowp + gx);
        }
    }
    __syncthreads();

    const int tx = tid % BX, ty = tid / BX;
    float acc[TY][TX];
#pragma unroll
    for (int oy = 0; oy < TY; ++oy)
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) acc[oy][ox] = 0.0f;

    const float* tt = tile + (ty * TY) * P + tx * TX;
    switch (RT) {
        case 0: {
            const float w = wsm[0];
#pragma unroll
            for (int oy = 0; oy < TY; ++oy)
#pragma unroll
                for (int ox = 0; ox < TX; ++ox) acc[oy][ox] = w * tt[oy * P + ox];
        } break;
        case 2: accumulate_rows<TX, TY, 5>(tt, P, wsm, acc); break;
        case 4: accumulate_rows<TX, TY, 9>(tt, P, wsm, acc); break;
        case 6: accumulate_rows<TX, TY, 13>(tt, P, wsm, acc); break;
        case 8: accumulate_rows<TX, TY, 17>(tt, P, wsm, acc); break;
        default: accumulate_rows<TX, TY, 21>(tt, P, wsm, acc); break;
    }

    float* op = out + (size_t)plane * H * W;
    const int ox0 = x0 + tx * TX;
#pragma unroll
    for (int oy = 0; oy < TY; ++oy) {
        const int y = y0 + ty * TY + oy;
        if (y >= H) break;
        float* orow = op + (size_t)y * W + ox0;
        if (vec_ok && ox0 + TX <= W) {
#pragma unroll
            for (int q = 0; q < TX / 4; ++q)
                reinterpret_cast<float4*>(orow)[q] =
                    make_float4(acc[oy][4 * q], acc[oy][4 * q + 1], acc[oy][4 * q + 2], acc[oy][4 * q + 3]);
        } else {
#pragma unroll
            for (int ox = 0; ox < TX; ++ox)
                if (ox0 + ox < W) orow[ox] = acc[oy][ox];
        }
    }
}

// Generic path for K > 21 (e.g. a 51x51 USM kernel pushed through filter2d): one
// output per thread, taps and pixels straight from L1/L2.  Correct, not fast; the
// fast USM route is otf_usm_sharp_f32 (exactly separable).
__global__ void filter2d_generic_kernel(const float* __restrict__ img, const float* __restrict__ kern,
                                        float* __restrict__ out, int C, int H, int W, int K, int kernel_batch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int plane = blockIdx.z;
    if (x >= W || y >= H) return;
    const int kb = kernel_batch == 1 ? 0 : plane / C;
    const float* kp = kern + (size_t)kb * K * K;
    const float* ip = img + (size_t)plane * H * W;
    const int r = K / 2;
    float acc = 0.0f;
    for (int i = 0; i < K; ++i) {
        const int gy = reflect_idx(y - r + i, H);
        for (int j = 0; j < K; ++j) {
            const int gx = reflect_idx(x - r + j, W);
            acc = fmaf(__ldg(kp + i * K + j), __ldg(ip + (size_t)gy * W + gx), acc);
        }
    }
    out[(size_t)plane * H * W + (size_t)y * W + x] = acc;
}

template <int TX, int TY, int BX, int BY>
static int launch_blocked(const float* img, int B, int C, int H, int W, const float* kernel, int kernel_batch, int K,
                          const int32_t* support, float* out, cudaStream_t st) {
    constexpr int TILE_W = TX * BX, TILE_H = TY * BY;
    constexpr int P = TILE_W + 2 * kMaxRT;
    const size_t smem = ((size_t)(TILE_H + 2 * kMaxRT) * P + 21 * kWPitch) * sizeof(float);
    auto kfn = filter2d_kernel<TX, TY, BX, BY>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "filter2d smem attribute");
    }
    const dim3 grid(ceil_div(W, TILE_W), ceil_div(H, TILE_H), B * C);
    const int vec_ok = (W % 4 == 0) && (((uintptr_t)out & 15) == 0);
    kfn<<<grid, BX * BY, smem, st>>>(img, kernel, support, out, C, H, W, K, kernel_batch, vec_ok);
    OTF_LAUNCH_CHECK("filter2d_kernel");
    return OTF_OK;
}

}  // namespace otf

extern "C" int otf_filter2d_f32(const float* img, int B, int C, int H, int W, const float* kernel, int kernel_batch,
                                int K, int32_t* support_dev, float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(img && kernel && out, OTF_ERR_BAD_ARG, "filter2d: null pointer");
    OTF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OTF_ERR_BAD_ARG, "filter2d: bad extents %d %d %d %d", B, C, H, W);
    OTF_REQUIRE(K > 0 && (K % 2) == 1, OTF_ERR_BAD_ARG, "Wrong kernel size");
    OTF_REQUIRE(kernel_batch == 1 || kernel_batch == B, OTF_ERR_BAD_ARG, "filter2d: kernel batch %d vs %d", kernel_batch, B);
    OTF_REQUIRE(K / 2 < H && K / 2 < W, OTF_ERR_BAD_ARG, "filter2d: reflect pad %d needs H,W > pad (got %dx%d)", K / 2, H, W);
    OTF_REQUIRE((int64_t)B * C <= 65535, OTF_ERR_UNSUPPORTED, "filter2d: B*C > 65535");
    OTF_REQUIRE(img != out, OTF_ERR_BAD_ARG, "filter2d: in-place not supported");
    cudaStream_t st = (cudaStream_t)stream;
    if (K > 2 * kMaxRT + 1) {
        const dim3 blk(32, 8), grid(ceil_div(W, 32), ceil_div(H, 8), B * C);
        filter2d_generic_kernel<<<grid, blk, 0, st>>>(img, kernel, out, C, H, W, K, kernel_batch);
        OTF_LAUNCH_CHECK("filter2d_generic_kernel");
        return OTF_OK;
    }
    if (support_dev) {
        kernel_support_kernel<<<kernel_batch, 128, 0, st>>>(kernel, K, support_dev);
        OTF_LAUNCH_CHECK("kernel_support_kernel");
    }
    // big planes: 64x64 tiles, 8x8 outputs per thread (64 threads); small planes: 32x32 tiles, 4x4 per thread
    const int64_t big_tiles = (int64_t)ceil_div(W, 64) * ceil_div(H, 64) * B * C;
    if (big_tiles >= 2 * kNumSMs) return launch_blocked<8, 8, 8, 8>(img, B, C, H, W, kernel, kernel_batch, K, support_dev, out, st);
    return launch_blocked<4, 4, 8, 8>(img, B, C, H, W, kernel, kernel_batch, K, support_dev, out, st);
}

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
#include <cooperative_groups.h>
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "otf_common.cuh"

namespace otf {

// ---- per-kernel analysis: true support, rank-1 factors, processing order --------------
//   support[kb]  largest |offset| with a non-zero tap (kernels arrive zero-padded to 21x21);
//   rank1[kb]    bit 1 (value 2): the kernel is left-right mirror symmetric (horizontal fold);
//                bit 0: set when the kernel is an outer product u v^T to within fp32 rounding of its own
//                taps (isotropic Gaussians — 45 % of the reference's default kernel_prob — and the
//                USM kernel are), with the factors in uv[kb][0][*] (rows) and uv[kb][1][*] (columns),
//                stored centred in 21 slots; the blur is then evaluated as K + K taps instead of K*K.
//                Acceptance: |K_ij - u_i v_j| <= 4e-7 |K_ij| + 1e-9 max|K|, i.e. the separable result
//                differs from the full sum by <= ~4e-7 * sum|K_ij| * max|img| — far inside the 1e-5 bar;
//   order[0..kb) one packed record per launch position (sample | radius << 16 | flags << 24), most expensive
//                sample first: CTAs are launched in that order, so the CTAs resident on an SM at any time run
//                the same specialisation (instruction-cache locality) and the expensive tiles go first (LPT).
//   staged[kb]   the taps exactly as filter2d_kernel wants them in shared memory (22 x 24 float2 slots per sample): the
//                rank-1 factors (u in words [0,24), v in [24,48)), or the folded / dense PAIRED taps — row i holds
//                (w[i][.], w[i-1][.]) for the FFMA2 loops.  A CTA then fetches its taps with ONE bulk copy that
//                completes on the same mbarrier as its tile instead of 4-5 rounds of dependent global loads.
constexpr int kUVPitch = 24;
constexpr int kW2Pitch = 24;                      // float2 per paired tap row
constexpr int kStagedWords = 22 * kW2Pitch * 2;   // per sample
__host__ __device__ inline size_t staged_offset(int kb) { return ((size_t)3 * kb + 3) & ~(size_t)3; }
__host__ __device__ inline size_t scratch_words(int kb) { return staged_offset(kb) + (size_t)kb * kStagedWords; }

struct KernelSets {
    const float* ptr[4];  // up to 4 kernel tensors (kernel1, kernel2, sinc_kernel, ...) analysed by one launch
};

// One thread-block CLUSTER of 8 CTAs per kernel tensor, one WARP per kernel (64 warps: a batch of 64 kernels is analysed in
// one pass, larger batches loop), then — behind the cluster barrier, which orders the CTAs' global writes — CTA 0 of the
// cluster ranks the samples: analysis and launch order in ONE launch with no host-initialised ticket.
constexpr int kAnalyseWarps = 8, kAnalyseCluster = 8;
__global__ void __cluster_dims__(kAnalyseCluster, 1, 1) __launch_bounds__(kAnalyseWarps * 32)
    kernel_analyse_kernel(const __grid_constant__ KernelSets sets, int K, int kernel_batch, int32_t* __restrict__ scratch_base, int interleave, float trim_tol) {
    pdl_enter();
    __shared__ float sk_all[kAnalyseWarps][21 * 21 + 7];
    extern __shared__ int s_sup[];  // [kernel_batch] supports, for the ranking pass (CTA 0 of the cluster)
    const float* kern = sets.ptr[blockIdx.y];
    int32_t* scratch = scratch_base + (size_t)blockIdx.y * scratch_words(kernel_batch);
    int32_t* support = scratch;
    int32_t* rank1 = scratch + 2 * kernel_batch;
    float* staged = reinterpret_cast<float*>(scratch + staged_offset(kernel_batch));
    const int c = K / 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = K * K;  // K <= 21 on this path
    float* sk = sk_all[warp];
    for (int kb = blockIdx.x * kAnalyseWarps + warp; kb < kernel_batch; kb += kAnalyseWarps * kAnalyseCluster) {
        const float* kp = kern + (size_t)kb * n;
        // a lane's taps (idx = lane + 32 t, t < 14 covers 21 x 21) live in registers: the 14 global loads are independent
        // (one round trip instead of 14 serialised ones — the launch sits on the chain's critical path in front of blur1)
        // and the passes below reuse value, ring and position without re-deriving them
        constexpr int NIT = 14;
        float tv[NIT];
        int ti[NIT], tj[NIT];
#pragma unroll
        for (int t = 0; t < NIT; ++t) {
            const int idx = lane + 32 * t;
            tv[t] = idx < n ? __ldg(kp + idx) : 0.0f;
            ti[t] = idx / K;
            tj[t] = idx - ti[t] * K;
        }
        int r = 0, imax = 0;
        float amax = 0.0f, tot = 0.0f;
#pragma unroll
        for (int t = 0; t < NIT; ++t) {
            const int idx = lane + 32 * t;
            if (idx < n) {
                sk[idx] = tv[t];
                if (tv[t] != 0.0f) r = max(r, max(abs(ti[t] - c), abs(tj[t] - c)));
                if (fabsf(tv[t]) > amax) { amax = fabsf(tv[t]); imax = idx; }
                tot += fabsf(tv[t]);
            }
        }
        r = __reduce_max_sync(0xffffffffu, r);
        // EFFECTIVE support: outer rings whose taps add up to less than trim_tol * sum|w| are dropped.  A sigma = 1 Gaussian
        // drawn into a 21x21 kernel has non-zero taps out to the corners (exp(-50) is representable), but every ring beyond
        // radius 6 together weighs < 1e-7 of the kernel: leaving them out moves an output by <= trim_tol * sum|w| * max|img|
        // (2e-7 for a normalised kernel on [0,1] pixels — the size of fp32 summation-order noise, 50x inside the 1e-5 bar)
        // and shrinks the K^2 loop to the part of the kernel that matters.  trim_tol = 0 (OTF_F2D_TRIM=0) keeps every tap.
        if (trim_tol > 0.0f && r > 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            float dropped = 0.0f;
            for (int ring = r; ring >= 1; --ring) {
                float rs = 0.0f;
#pragma unroll
                for (int t = 0; t < NIT; ++t)
                    if (lane + 32 * t < n && max(abs(ti[t] - c), abs(tj[t] - c)) == ring) rs += fabsf(tv[t]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
                if (dropped + rs > trim_tol * tot) break;
                dropped += rs;
                r = ring - 1;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {  // pivot = the largest |tap| (ties -> lowest index)
            const float oa = __shfl_xor_sync(0xffffffffu, amax, o);
            const int oi = __shfl_xor_sync(0xffffffffu, imax, o);
            if (oa > amax || (oa == amax && oi < imax)) { amax = oa; imax = oi; }
        }
        __syncwarp();
        const int pi = imax / K, pj = imax - pi * K;
        const float piv = sk[imax];
        bool ok = amax > 0.0f, sym = true;
#pragma unroll
        for (int t = 0; t < NIT; ++t) {
            if (lane + 32 * t < n) {
                const int i = ti[t], j = tj[t];
                const float kij = tv[t];
                if (ok) {
                    const float sep = sk[i * K + pj] * __fdiv_rn(sk[pi * K + j], piv);
                    ok = fabsf(kij - sep) <= 4e-7f * fabsf(kij) + 1e-9f * amax;
                }
                // left-right mirror symmetry, K[i][j] == K[i][K-1-j] bit for bit (every isotropic family and the sinc
                // kernels are: the generators evaluate a function of x^2): such kernels take the horizontal fold
                sym = sym && (kij == sk[i * K + (K - 1 - j)]);
            }
        }
        ok = __all_sync(0xffffffffu, ok);
        sym = __all_sync(0xffffffffu, sym);
        float* st = staged + (size_t)kb * kStagedWords;
        if (ok && r >= 2) {
            if (lane < kUVPitch) {
                // centred in 21 slots: slot s <-> offset s - 10
                const int t = lane - 10 + c;  // tap index for this slot
                const bool in = lane < 21 && t >= 0 && t < K;
                st[lane] = in ? sk[t * K + pj] : 0.0f;
                st[kUVPitch + lane] = in ? __fdiv_rn(sk[pi * K + t], piv) : 0.0f;
                // the same factors over the true support, packed for the FFMA2 loops of rank1_tile:
                //   float2 vpair[m] = (v[m-1], v[m])               at words [48, 96)
                //   float4 upq[m]   = (u[m], u[m-1], u[m-2], u[m-3]) at words [96, 192),   m = 0..23, zero outside [0, kt)
                const int kt = 2 * r + 1, m = lane;
                auto uf = [&](int q) { return (q >= 0 && q < kt) ? sk[(c - r + q) * K + pj] : 0.0f; };
                auto vf = [&](int q) { return (q >= 0 && q < kt) ? __fdiv_rn(sk[pi * K + (c - r + q)], piv) : 0.0f; };
                reinterpret_cast<float2*>(st)[kUVPitch + m] = make_float2(vf(m - 1), vf(m));
                reinterpret_cast<float4*>(st)[kUVPitch + m] = make_float4(uf(m), uf(m - 1), uf(m - 2), uf(m - 3));
            }
        } else {
            // paired taps over the true support: row i holds (w[i][j], w[i-1][j]); mirror-symmetric kernels keep the right
            // half only, (w[i][r+t], w[i-1][r+t]) for t = 0..r  (accumulate_rows_packed / accumulate_rows_folded)
            const bool fo = sym && r >= 2;
            const int kt = 2 * r + 1;
            auto tap = [&](int i, int j) { return (i >= 0 && i < kt && j >= 0 && j < kt) ? sk[(c - r + i) * K + (c - r + j)] : 0.0f; };
            float2* st2 = reinterpret_cast<float2*>(st);
            for (int idx = lane; idx < (kt + 1) * kW2Pitch; idx += 32) {
                const int i = idx / kW2Pitch, t = idx - i * kW2Pitch;
                const int j = fo ? (t <= r ? r + t : -1) : t;
                st2[idx] = make_float2(tap(i, j), tap(i - 1, j));
            }
        }
        if (lane == 0) { support[kb] = r; rank1[kb] = (ok ? 1 : 0) | (sym ? 2 : 0); }
        __syncwarp();
    }
    cooperative_groups::this_cluster().sync();  // (release / acquire at cluster scope: every CTA's supports are visible)
    if (blockIdx.x != 0) return;
    // order[0..kb): one packed record per launch position — sample index | true radius << 16 | flags << 24 — sorted by the
    // sample's COST (FP32-pipe operations per output: dense K^2, folded ~K(R+1), rank-1 ~3K), most expensive first, ties by
    // index: the CTAs resident on an SM at any time run the same specialisation, the expensive tiles go first (LPT), and a
    // CTA learns its sample, radius and kind from ONE load instead of three dependent ones.
    for (int t = threadIdx.x; t < kernel_batch; t += blockDim.x) {
        const int r = support[t], fl = rank1[t], kt = 2 * r + 1;
        const int cost = (r >= 2 && (fl & 1)) ? 3 * kt : (r >= 2 && (fl & 2)) ? kt * (r + 1) + 6 * r : kt * kt;
        s_sup[t] = (cost << 12) | (r << 4) | (fl & 3);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kernel_batch; t += blockDim.x) {
        const int mine = s_sup[t] >> 12;
        int rank = 0;
        for (int u = 0; u < kernel_batch; ++u) {
            const int other = s_sup[u] >> 12;
            rank += (other > mine) || (other == mine && u < t);
        }
        // interleave: launch positions alternate between the expensive and the cheap end of the ranking, so FMA-bound
        // (dense) and latency-bound (rank-1, small K) CTAs share every SM instead of running one kind after the other
        const int pos = !interleave ? rank : (rank < (kernel_batch + 1) / 2 ? 2 * rank : 2 * (kernel_batch - 1 - rank) + 1);
        scratch[kernel_batch + pos] = t | (((s_sup[t] >> 4) & 0xff) << 16) | ((s_sup[t] & 3) << 24);
    }
}

static int analyse_sets(const float* const* kernels, int nsets, int kernel_batch, int K, int32_t* scratch, cudaStream_t st) {
    OTF_REQUIRE(nsets >= 1 && nsets <= 4, OTF_ERR_BAD_ARG, "filter2d_analyse: 1..4 kernel sets per call");
    OTF_REQUIRE(kernel_batch >= 1 && kernel_batch <= 4096, OTF_ERR_UNSUPPORTED, "filter2d: kernel batch must be 1..4096");
    KernelSets sets;
    for (int i = 0; i < 4; ++i) sets.ptr[i] = kernels[i < nsets ? i : 0];
    static const int interleave = getenv("OTF_F2D_INTERLEAVE") ? atoi(getenv("OTF_F2D_INTERLEAVE")) : 0;
    // OTF_F2D_TRIM=<tolerance> (0 = keep every non-zero tap): relative weight of the outer rings the effective support may drop
    static const float trim_tol = getenv("OTF_F2D_TRIM") ? (float)atof(getenv("OTF_F2D_TRIM")) : 2e-7f;
    launch_chain(kernel_analyse_kernel, dim3(dim3(kAnalyseCluster, nsets)), dim3(kAnalyseWarps * 32), kernel_batch * sizeof(int), st, sets, K, kernel_batch, scratch, interleave, trim_tol);
    OTF_LAUNCH_CHECK("kernel_analyse_kernel");
    return OTF_OK;
}

constexpr int kMaxRT = 10;   // register-blocked path covers radius <= 10 (K <= 21)
constexpr int kMaxRA = 12;   // halo rounded up to a float4 boundary
constexpr int kPitchPad = 4;  // makes the row pitch an ODD number of 16-byte chunks (conflict-free two-row LDS.128 phases, rank1_tile)
constexpr int kWPitch = 24;  // taps row pitch in smem (float4 broadcast loads)

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Accumulate a TY x TX register block.  The tile's column 0 sits RA = roundup4(R) pixels left of the
// block's first output column, so every row window starts on a 16-byte boundary (LDS.128) and tap j
// of output ox reads row[(RA - R) + ox + j].
template <int TX, int TY, int KT>
__device__ __forceinline__ void accumulate_rows(const float* __restrict__ tile_thread, int pitch,
                                                const float* __restrict__ wsm, float (&acc)[TY][TX]) {
    constexpr int R = KT / 2, RA = (R + 3) & ~3, OFF = RA - R;
    constexpr int NROW = TX + 2 * RA;
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
                            for (int ox = 0; ox < TX; ++ox) acc[oy][ox] = fmaf(w[t], row[OFF + ox + j], acc[oy][ox]);
                        }
                    }
                }
            }
        }
    }
}

// Packed variant of accumulate_rows for sm_100a: FFMA2 (fma.rn.f32x2) performs two fp32 FMAs per
// issued instruction.  The FMA pipe has the same peak either way (measured 73 TFLOP/s with both,
// profiles/experiments/ffma2_rate.cu) but the packed form needs half the issue slots, which leaves
// room for the LDS / address work that kept the scalar loop at ~2/3 of the FMA peak.
// Pairing is over two vertically adjacent outputs: they read the SAME pixel with the taps of two
// consecutive kernel rows, so operand a = (pixel, pixel), operand b = (w[i][j], w[i-1][j]) comes
// pre-paired from shared memory (rows i = 0..KT, out-of-range taps zero) and the accumulator pair is
// (acc[2p][ox], acc[2p+1][ox]).  Each lane is an ordinary IEEE fp32 FMA: results are bit-identical
// to the scalar loop.
#ifndef OTF_F2D_MINB
#define OTF_F2D_MINB 5
#endif
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
    const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

template <int TX, int TY, int KT>
__device__ __forceinline__ void accumulate_rows_packed(const float* __restrict__ tile_thread, int pitch,
                                                       const float2* __restrict__ wsm2, float (&acc)[TY][TX]) {
    static_assert(TY % 2 == 0, "pairs of output rows");
    constexpr int R = KT / 2, RA = (R + 3) & ~3, OFF = RA - R;
    constexpr int NROW = TX + 2 * RA;
    float2 acc2[TY / 2][TX];
#pragma unroll
    for (int p = 0; p < TY / 2; ++p)
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) acc2[p][ox] = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int r = 0; r < TY + KT - 1; ++r) {
        float2 rd[NROW];  // every pixel of the row window duplicated into a register pair
        const float4* rp = reinterpret_cast<const float4*>(tile_thread + r * pitch);
#pragma unroll
        for (int q = 0; q < NROW / 4; ++q) {
            const float4 v = rp[q];
            rd[4 * q + 0] = make_float2(v.x, v.x); rd[4 * q + 1] = make_float2(v.y, v.y);
            rd[4 * q + 2] = make_float2(v.z, v.z); rd[4 * q + 3] = make_float2(v.w, v.w);
        }
#pragma unroll
        for (int p = 0; p < TY / 2; ++p) {
            const int i = r - 2 * p;  // output row 2p uses kernel row i, output row 2p+1 uses kernel row i-1
            if (i >= 0 && i <= KT) {
                const float4* wp = reinterpret_cast<const float4*>(wsm2 + i * kW2Pitch);
#pragma unroll
                for (int q = 0; q < (KT + 1) / 2; ++q) {
                    const float4 w4 = wp[q];  // taps j = 2q, 2q+1 as (lo, hi) pairs
                    const float2 w[2] = {make_float2(w4.x, w4.y), make_float2(w4.z, w4.w)};
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int j = 2 * q + t;
                        if (j < KT) {
#pragma unroll
                            for (int ox = 0; ox < TX; ++ox) ffma2(acc2[p][ox], rd[OFF + ox + j], w[t]);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < TY / 2; ++p)
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) {
            acc[2 * p][ox] = acc2[p][ox].x;
            acc[2 * p + 1][ox] = acc2[p][ox].y;
        }
}

// Horizontal fold for left-right mirror-symmetric kernels (w[i][R+t] == w[i][R-t]): the two pixels that share a tap
// are added first, s = p[x+t] + p[x-t], and the sum feeds the FFMA2 of both vertically paired output rows — per image
// row and output column R FADDs + 2(R+1) FFMA2 instead of 2(2R+1) FFMA2: 1.56x fewer FMA-pipe cycles at K = 21.
// wsm2f row i holds the paired taps (w[i][R+t], w[i-1][R+t]) for t = 0..R.  The summation order differs from the
// unfolded loop (pairs are pre-added), i.e. by a few ulp — inside the 1e-5 bar like every other order.
template <int TX, int TY, int KT, bool ALL>
__device__ __forceinline__ void fold_row(const float* __restrict__ rowp, const float2* __restrict__ wsm2f, int r,
                                         float2 (&acc2)[TY / 2][TX]) {
    constexpr int R = KT / 2, RA = (R + 3) & ~3, OFF = RA - R;
    constexpr int NROW = TX + 2 * RA;
    float row[NROW];
    const float4* rp = reinterpret_cast<const float4*>(rowp);
#pragma unroll
    for (int q = 0; q < NROW / 4; ++q) {
        const float4 v = rp[q];
        row[4 * q + 0] = v.x; row[4 * q + 1] = v.y; row[4 * q + 2] = v.z; row[4 * q + 3] = v.w;
    }
    // output pair p takes kernel rows (i, i - 1) with i = r - 2p from image row r: live while 0 <= i <= KT (block-uniform)
    bool live[TY / 2];
#pragma unroll
    for (int p = 0; p < TY / 2; ++p) live[p] = ALL || (r - 2 * p >= 0 && r - 2 * p <= KT);
#pragma unroll
    for (int q = 0; q < (R + 2) / 2; ++q) {  // taps t = 2q, 2q + 1
        float4 w4[TY / 2];
#pragma unroll
        for (int p = 0; p < TY / 2; ++p)
            if (live[p]) w4[p] = reinterpret_cast<const float4*>(wsm2f + (r - 2 * p) * kW2Pitch)[q];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = 2 * q + u;
            if (t <= R) {
#pragma unroll
                for (int ox = 0; ox < TX; ++ox) {
                    const int c = OFF + ox + R;
                    const float sv = t == 0 ? row[c] : __fadd_rn(row[c + t], row[c - t]);
                    const float2 s2 = make_float2(sv, sv);
#pragma unroll
                    for (int p = 0; p < TY / 2; ++p)
                        if (live[p])
                            ffma2(acc2[p][ox], s2, u == 0 ? make_float2(w4[p].x, w4[p].y) : make_float2(w4[p].z, w4[p].w));
                }
            }
        }
    }
}

template <int TX, int TY, int KT>
__device__ __forceinline__ void accumulate_rows_folded(const float* __restrict__ tile_thread, int pitch,
                                                       const float2* __restrict__ wsm2f, float (&acc)[TY][TX]) {
    static_assert(TY % 2 == 0, "pairs of output rows");
    float2 acc2[TY / 2][TX];
#pragma unroll
    for (int p = 0; p < TY / 2; ++p)
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) acc2[p][ox] = make_float2(0.0f, 0.0f);
    // image rows 0 .. TY + KT - 2: every pair is live for rows TY - 2 .. KT (one loop body without predicates when that
    // range is not empty), the ramps on either side test the pairs one by one
    constexpr int LO = TY - 2, HI = KT, LAST = TY + KT - 2;
    if constexpr (LO <= HI) {
#pragma unroll 1
        for (int r = 0; r < LO; ++r) fold_row<TX, TY, KT, false>(tile_thread + r * pitch, wsm2f, r, acc2);
#pragma unroll 1
        for (int r = LO; r <= HI; ++r) fold_row<TX, TY, KT, true>(tile_thread + r * pitch, wsm2f, r, acc2);
#pragma unroll 1
        for (int r = HI + 1; r <= LAST; ++r) fold_row<TX, TY, KT, false>(tile_thread + r * pitch, wsm2f, r, acc2);
    } else {
#pragma unroll 1
        for (int r = 0; r <= LAST; ++r) fold_row<TX, TY, KT, false>(tile_thread + r * pitch, wsm2f, r, acc2);
    }
#pragma unroll
    for (int p = 0; p < TY / 2; ++p)
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) {
            acc[2 * p][ox] = acc2[p][ox].x;
            acc[2 * p + 1][ox] = acc2[p][ox].y;
        }
}

// Rank-1 kernels: out = u (x) v correlated with the tile, evaluated per thread as a horizontal
// K-tap pass over each row of its window (kept in registers) followed by the vertical accumulation.
// u and v sit centred in 21 slots (tap t of a KT-tap kernel is slot 10 - R + t).
template <int TX, int TY, int KT>
__device__ __forceinline__ void accumulate_rank1(const float* __restrict__ tile_thread, int pitch,
                                                 const float* __restrict__ u, const float* __restrict__ v,
                                                 float (&acc)[TY][TX]) {
    constexpr int R = KT / 2, RA = (R + 3) & ~3, OFF = RA - R;
    constexpr int NROW = TX + 2 * RA;
    float vr[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) vr[j] = v[10 - R + j];
#pragma unroll 1
    for (int r = 0; r < TY + KT - 1; ++r) {
        float row[NROW];
        const float4* rp = reinterpret_cast<const float4*>(tile_thread + r * pitch);
#pragma unroll
        for (int q = 0; q < NROW / 4; ++q) {
            const float4 t4 = rp[q];
            row[4 * q + 0] = t4.x; row[4 * q + 1] = t4.y; row[4 * q + 2] = t4.z; row[4 * q + 3] = t4.w;
        }
        float h[TX];
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) {
            float a = vr[0] * row[OFF + ox];
#pragma unroll
            for (int j = 1; j < KT; ++j) a = fmaf(vr[j], row[OFF + ox + j], a);
            h[ox] = a;
        }
#pragma unroll
        for (int oy = 0; oy < TY; ++oy) {
            const int i = r - oy;
            if (i >= 0 && i < KT) {
                const float ui = u[10 - R + i];
#pragma unroll
                for (int ox = 0; ox < TX; ++ox) acc[oy][ox] = fmaf(ui, h[ox], acc[oy][ox]);
            }
        }
    }
}

// Rank-1 kernels, CTA-wide two-pass form (the default): the horizontal K-tap pass runs ONCE per tile row for the whole
// CTA and its result replaces the row in shared memory (in place), then every thread accumulates a column strip.  Per
// output that is K (TILE_H + 2R) / TILE_H + K FMAs — 48.6 at K = 21 on a 64-row tile — where the per-thread form above
// recomputes the horizontal pass for each TY-row register block (K (TY + K - 1) / TY + K = 147).  With so little
// arithmetic left the pass is bound by SHARED-MEMORY bandwidth (ncu: l1tex 76 %, short-scoreboard the top stall), so
// both passes are laid out for few, conflict-free wavefronts:
//   horizontal: a lane produces 8 consecutive outputs of one tile row from a (8 + 2 RA)-pixel window (LDS.128s); the eight
//     lanes of an LDS phase are 2 rows x 4 segments, and the row pitch is an odd number of 16-byte chunks, so a phase
//     touches 8 distinct chunks mod 8.  The 32 lanes of a warp cover whole rows: a row is read and rewritten by ONE warp
//     (loads, __syncwarp, FMAs, stores), no block-wide barrier inside the pass.  FFMA2 over two horizontally adjacent
//     outputs that share the pixel: (out[2k+1], out[2k]) += p * (v[m-1], v[m]), p = win[OFF + 2k + m], m = 0..KT.
//   vertical: a thread owns 4 columns x RPT rows (8 consecutive lanes = 8 consecutive chunks), one LDS.128 per tile row,
//     FFMA2 over two vertically adjacent outputs: (acc[2p], acc[2p+1]) += h * (u[i], u[i-1]), i = r - 2p; the factor pairs of
//     two output pairs arrive as one broadcast LDS.128 (upq).
// Every output is still u-major sum of ascending fmaf chains (fma(v0, p, 0) == v0 * p): bit-identical to accumulate_rank1.
template <int P, int TILE_W, int TILE_H, int NT, int KT>
__device__ __forceinline__ void rank1_tile(float* __restrict__ tile, const float* __restrict__ wsm, int tid,
                                           float* __restrict__ op, int x0, int y0, int H, int W, int vec_ok) {
    constexpr int R = KT / 2, RA = (R + 3) & ~3, OFF = RA - R;
    constexpr int SEGS = TILE_W / 8, RPW = 32 / SEGS, NWARP = NT / 32, NWIN = 8 + 2 * RA, ROWS_H = TILE_H + KT - 1;
    static_assert(SEGS == 8 || SEGS == 4, "64- or 32-wide tiles");
    static_assert((P / 4) % 2 == 1, "odd chunk pitch");
    const int lane = tid & 31, warp = tid >> 5;
    {
        const float2* vpair = reinterpret_cast<const float2*>(wsm) + kUVPitch;
        float2 vp[KT + 1];
#pragma unroll
        for (int m = 0; m <= KT; ++m) vp[m] = vpair[m];  // (v[m-1], v[m])
        const int g = lane >> 3, rowbit = (lane >> 2) & 1, segq = lane & 3;
        const int rw = SEGS == 8 ? 2 * (g >> 1) + rowbit : 2 * g + rowbit;
        const int seg = SEGS == 8 ? 4 * (g & 1) + segq : segq;
#pragma unroll 1
        for (int r0 = warp * RPW; r0 < ROWS_H; r0 += NWARP * RPW) {
            const int r = r0 + rw;
            const bool live = r < ROWS_H;
            float* rowp = tile + min(r, ROWS_H - 1) * P + 8 * seg;
            float win[NWIN];
#pragma unroll
            for (int q = 0; q < NWIN / 4; ++q) {
                const float4 t4 = reinterpret_cast<const float4*>(rowp)[q];
                win[4 * q + 0] = t4.x; win[4 * q + 1] = t4.y; win[4 * q + 2] = t4.z; win[4 * q + 3] = t4.w;
            }
            __syncwarp();
            float2 a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = make_float2(0.0f, 0.0f);  // (out[2k+1], out[2k])
#pragma unroll
            for (int m = 0; m <= KT; ++m)
#pragma unroll
                for (int k = 0; k < 4; ++k) ffma2(a[k], make_float2(win[OFF + 2 * k + m], win[OFF + 2 * k + m]), vp[m]);
            if (live) {
                reinterpret_cast<float4*>(rowp + RA)[0] = make_float4(a[0].y, a[0].x, a[1].y, a[1].x);
                reinterpret_cast<float4*>(rowp + RA)[1] = make_float4(a[2].y, a[2].x, a[3].y, a[3].x);
            }
        }
    }
    __syncthreads();
    constexpr int QUADS = TILE_W / 4, STRIPS = NT / QUADS, RPT = TILE_H / STRIPS;
    static_assert(RPT % 4 == 0, "output rows in groups of four");
    const int quad = tid % QUADS, strip = tid / QUADS;
    float2 acc2[RPT / 2][4];
#pragma unroll
    for (int p = 0; p < RPT / 2; ++p)
#pragma unroll
        for (int cx = 0; cx < 4; ++cx) acc2[p][cx] = make_float2(0.0f, 0.0f);
    const float4* upq = reinterpret_cast<const float4*>(wsm) + kUVPitch;  // (u[i], u[i-1], u[i-2], u[i-3])
    const float* base = tile + (strip * RPT) * P + RA + 4 * quad;
#pragma unroll 1
    for (int r = 0; r < RPT + KT - 1; ++r) {
        const float4 h4 = *reinterpret_cast<const float4*>(base + r * P);
        const float h[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
        for (int pp = 0; pp < RPT / 4; ++pp) {
            const int i = r - 4 * pp;  // rows 4pp, 4pp+1 take (u[i], u[i-1]); rows 4pp+2, 4pp+3 take (u[i-2], u[i-3])
            if (i >= 0 && i <= KT + 2) {
                const float4 w = upq[i];
#pragma unroll
                for (int cx = 0; cx < 4; ++cx) {
                    ffma2(acc2[2 * pp][cx], make_float2(h[cx], h[cx]), make_float2(w.x, w.y));
                    ffma2(acc2[2 * pp + 1][cx], make_float2(h[cx], h[cx]), make_float2(w.z, w.w));
                }
            }
        }
    }
    const int x = x0 + 4 * quad;
#pragma unroll
    for (int p = 0; p < RPT / 2; ++p)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int y = y0 + strip * RPT + 2 * p + half;
            if (y < H) {
                float v[4];
#pragma unroll
                for (int cx = 0; cx < 4; ++cx) v[cx] = half ? acc2[p][cx].y : acc2[p][cx].x;
                float* orow = op + (size_t)y * W + x;
                if (vec_ok && x + 4 <= W) {
                    *reinterpret_cast<float4*>(orow) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int cx = 0; cx < 4; ++cx)
                        if (x + cx < W) orow[cx] = v[cx];
                }
            }
        }
}

// ---- TMA / mbarrier plumbing (sm_100a) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // bounded spin: a descriptor mistake must fault the launch, not hang the GPU
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
// 3-D tiled bulk tensor copy global -> shared, completion signalled on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

#ifdef OTF_F2D_TRACE
// profiles/experiments/f2d_trace.cu: per-CTA phase timestamps (globaltimer ns at entry, SM clock at four points)
__device__ unsigned long long g_f2d_trace[8 * 8192];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define OTF_TRACE(slot) do { if (threadIdx.x == 0) g_f2d_trace[8 * trace_lin + (slot)] = gtime(); } while (0)
#else
#define OTF_TRACE(slot) do { } while (0)
#endif

// L2 prefetch of a tile box (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tmap, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(x), "r"(y), "r"(z) : "memory");
}

// 1-D bulk copy global -> shared completing on an mbarrier (size and both addresses multiples of 16 bytes)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int TX, int TY, int BX, int BY, bool PACKED>
__global__ void __launch_bounds__(BX* BY, PACKED ? OTF_F2D_MINB : 1) filter2d_kernel(const __grid_constant__ CUtensorMap tmap,
                                                          const float* __restrict__ img, const float* __restrict__ kern,
                                                          const int32_t* __restrict__ scratch, int use_order,
                                                          float* __restrict__ out, int C, int H, int W, int K,
                                                          int kernel_batch, int vec_ok, int use_tma, int pf_ahead) {
    pdl_enter();
#ifdef OTF_F2D_TRACE
    const int trace_lin = ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) & 8191;
    if (threadIdx.x == 0) g_f2d_trace[8 * trace_lin + 5] = smid();
    OTF_TRACE(0);
#endif
    constexpr int TILE_W = TX * BX, TILE_H = TY * BY, NT = BX * BY;
    constexpr int P = TILE_W + 2 * kMaxRA + kPitchPad;  // smem row pitch = TMA box width (multiple of 4, odd number of float4s)
    constexpr int ROWS = TILE_H + 2 * kMaxRT;  // TMA box height
    static_assert(P % 4 == 0 && P / 4 <= 32, "a warp fills one tile row with one float4 per lane");
    extern __shared__ __align__(128) float smem[];
    float* tile = smem;              // ROWS x P (rows [0, TILE_H + 2R) are used)
    constexpr int WSM_FLOATS = PACKED ? 22 * kW2Pitch * 2 : 21 * kWPitch;
    float* wsm = smem + ROWS * P;    // 21 x kWPitch taps, or 22 x kW2Pitch paired taps
    uint64_t* bar = reinterpret_cast<uint64_t*>(wsm + WSM_FLOATS);

    const int zb = blockIdx.z / C, zc = blockIdx.z - zb * C;
    int b = zb, R = K / 2, kflags = 0;  // sample, true radius (block-uniform), rank-1 / mirror-symmetry flags
    if (scratch && use_order) {
        const int rec = scratch[kernel_batch + zb];  // packed launch-order record (kernel_analyse_kernel)
        b = rec & 0xffff;
        R = min((rec >> 16) & 0xff, K / 2);
        kflags = rec >> 24;
    } else if (scratch) {
        const int kb0 = kernel_batch == 1 ? 0 : zb;
        R = min(scratch[kb0], K / 2);
        kflags = scratch[2 * kernel_batch + kb0];
    }
    const int plane = b * C + zc;
    const int kb = kernel_batch == 1 ? 0 : b;
    const bool rank1 = R >= 2 && (kflags & 1) != 0;
    const bool fold = PACKED && !rank1 && R >= 2 && (kflags & 2) != 0;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.x;
    const int KT = 2 * R + 1, RA = (R + 3) & ~3;
    const int c = K / 2;

    // halo tile -> smem.
    //  * aligned planes (W % 4 == 0): ONE elected thread issues a 3-D TMA box load (cp.async.bulk.tensor,
    //    out-of-bounds zero filled) that completes on an mbarrier; border tiles then mirror the
    //    out-of-plane columns/rows inside shared memory (torch "reflect": the source pixels are in the box);
    //  * any other plane: per-thread cp.async (16 bytes where the four pixels are inside the plane and
    //    aligned, else four 4-byte copies) with the reflect index resolved at issue.
    // Either way nothing is staged through registers and every byte of the tile is in flight at once;
    // the other CTAs resident on the SM (7 at 64x64) compute meanwhile.
    const float* ip = img + (size_t)plane * H * W;
    const int th = TILE_H + 2 * R, nq = (TILE_W + 2 * RA) / 4;
    const int lane = tid & 31;
    const bool staged_ok = scratch != nullptr && (PACKED || rank1);
    const float* stg = staged_ok ? reinterpret_cast<const float*>(scratch + staged_offset(kernel_batch)) + (size_t)kb * kStagedWords : nullptr;
    const uint32_t tap_bytes = rank1 ? 8 * kUVPitch * 4 : (KT + 1) * kW2Pitch * 8;
    if (use_tma) {
        if (tid == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(bar, ROWS * P * 4 + (staged_ok ? tap_bytes : 0u));
            tma_load_3d(tile, &tmap, x0 - RA, y0 - R, plane, bar);
            if (staged_ok) bulk_load(wsm, stg, tap_bytes, bar);
        } else if (tid == 32 && pf_ahead > 0) {
            // pull the tile of the CTA that will take this one's place (pf_ahead launch positions later) into L2 now: its
            // own box load then finds the bytes on chip, and the HBM requests of a whole extra wave are in flight
            const int per_plane = gridDim.x * gridDim.y;
            const int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x + pf_ahead;
            const int z2 = lin / per_plane, rem = lin - z2 * per_plane;
            if (z2 < (int)gridDim.z) {
                const int zb2 = z2 / C, zc2 = z2 - zb2 * C;
                int b2 = zb2, R2 = K / 2;
                if (scratch && use_order) {
                    const int rec = scratch[kernel_batch + zb2];
                    b2 = rec & 0xffff;
                    R2 = min((rec >> 16) & 0xff, K / 2);
                } else if (scratch) {
                    R2 = min(scratch[kernel_batch == 1 ? 0 : zb2], K / 2);
                }
                const int by2 = rem / gridDim.x, bx2 = rem - by2 * gridDim.x;
                tma_prefetch_3d(&tmap, bx2 * TILE_W - ((R2 + 3) & ~3), by2 * TILE_H - R2, b2 * C + zc2);
            }
        }
    } else if (lane < nq) {
        const int gx0 = x0 - RA + 4 * lane;
        const bool fast = vec_ok && gx0 >= 0 && gx0 + 3 < W;
        int gxs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) gxs[k] = clampi(reflect_idx(gx0 + k, W), 0, W - 1);
        for (int yy = tid >> 5; yy < th; yy += NT / 32) {
            const int gy = clampi(reflect_idx(y0 - R + yy, H), 0, H - 1);
            const float* rowp = ip + (size_t)gy * W;
            float* dst = tile + yy * P + 4 * lane;
            if (fast) {
                cp_async16(dst, rowp + gx0);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) cp_async4(dst + k, rowp + gxs[k]);
            }
        }
    }
    // taps -> smem.  With an analysis scratch the taps were laid out for this kernel once per sample (kernel_analyse_kernel:
    // rank-1 factors, or folded / dense paired taps) and arrive by the bulk copy issued above (TMA path) or a plain
    // coalesced copy; without one (raw C-ABI call, no scratch) they are paired here from the zero-padded kernel.
    const float* kp = kern + (size_t)kb * K * K;
    if (staged_ok) {
        if (!use_tma)
            for (int idx = tid; idx < (int)(tap_bytes / 4); idx += NT) wsm[idx] = stg[idx];
    } else if (PACKED && R >= 1) {
        // paired taps: row i holds (w[i][j], w[i-1][j]) for i = 0..KT, zeros outside the KT x KT centre
        float2* w2 = reinterpret_cast<float2*>(wsm);
        auto tap = [&](int i, int j) {
            const int si = c - R + i, sj = c - R + j;
            return (i >= 0 && i < KT && j < KT && si >= 0 && si < K && sj >= 0 && sj < K) ? kp[si * K + sj] : 0.0f;
        };
        const int nj = KT + 1;  // (rows 0..KT, taps 0..KT: the loops read nothing else)
        for (int idx = tid; idx < (KT + 1) * nj; idx += NT) {
            const int i = idx / nj, j = idx - i * nj;
            w2[i * kW2Pitch + j] = make_float2(tap(i, j), tap(i - 1, j));
        }
    } else
    for (int idx = tid; idx < 21 * kWPitch; idx += NT) {
        const int i = idx / kWPitch, j = idx - i * kWPitch;
        const int si = c - R + i, sj = c - R + j;
        float v = 0.0f;
        if (i < KT && j < KT && si >= 0 && si < K && sj >= 0 && sj < K) v = kp[si * K + sj];
        wsm[idx] = v;
    }
    cp_async_wait_all();
    __syncthreads();  // taps + cp.async pixels visible; mbarrier initialised
    if (use_tma) {
        mbar_wait(bar, 0);
        // mirror what the box load zero-filled: columns first (rows inside the plane), then whole rows
        const bool fix_x = (x0 - R < 0) || (x0 + TILE_W + R > W);
        const bool fix_y = (y0 - R < 0) || (y0 + TILE_H + R > H);
        const int warp = tid >> 5;
        if (fix_x) {
            // out-of-plane columns a stored output can reach: R on the left of x = 0, R on the right of x = W-1
            // (a warp per tile row, a lane per column: no index division)
            const int k = lane;
            const int gx = k < R ? -1 - k : W + (k - R);
            const int sc = gx - (x0 - RA), src = reflect_idx(gx, W) - (x0 - RA);
            const bool col_ok = k < 2 * R && sc >= 0 && sc < P && src >= 0 && src < P;
            for (int yy = warp; yy < th; yy += NT / 32) {
                const int gy = y0 - R + yy;
                if (col_ok && gy >= 0 && gy < H) tile[yy * P + sc] = tile[yy * P + src];
            }
            __syncthreads();
        }
        if (fix_y) {
            // rows above / below the plane only: [0, top) and [bot0, th)
            const int top = max(0, R - y0), bot0 = min(th, H - (y0 - R));
            for (int k = warp; k < top + (th - bot0); k += NT / 32) {
                const int yy = k < top ? k : bot0 + (k - top);
                const int src = clampi(clampi(reflect_idx(y0 - R + yy, H), 0, H - 1) - (y0 - R), 0, ROWS - 1);
                if (lane < P / 4) reinterpret_cast<float4*>(tile + yy * P)[lane] = reinterpret_cast<const float4*>(tile + src * P)[lane];
            }
            __syncthreads();
        }
    }

    OTF_TRACE(1);  // tile + taps ready
    const int tx = tid % BX, ty = tid / BX;
    float acc[TY][TX];
#pragma unroll
    for (int oy = 0; oy < TY; ++oy)
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) acc[oy][ox] = 0.0f;

    const float* tt = tile + (ty * TY) * P + tx * TX;
#ifndef OTF_F2D_RANK1_PER_THREAD
    if (rank1) {
        float* opl = out + (size_t)plane * H * W;
#define OTF_R1(KT_) rank1_tile<P, TILE_W, TILE_H, NT, KT_>(tile, wsm, tid, opl, x0, y0, H, W, vec_ok)
        switch (R) {
            case 2: OTF_R1(5); break;
            case 3: OTF_R1(7); break;
            case 4: OTF_R1(9); break;
            case 5: OTF_R1(11); break;
            case 6: OTF_R1(13); break;
            case 7: OTF_R1(15); break;
            case 8: OTF_R1(17); break;
            case 9: OTF_R1(19); break;
            default: OTF_R1(21); break;
        }
#undef OTF_R1
#ifdef OTF_F2D_TRACE
        OTF_TRACE(2); OTF_TRACE(3);
        if (threadIdx.x == 0) g_f2d_trace[8 * trace_lin + 4] = (unsigned long long)R | (1ull << 8);
#endif
        return;
    }
#endif
    if (rank1) {
        const float* u = wsm, *v = wsm + kUVPitch;
#define OTF_R1(KT_) accumulate_rank1<TX, TY, KT_>(tt, P, u, v, acc)
        switch (R) {
            case 2: OTF_R1(5); break;
            case 3: OTF_R1(7); break;
            case 4: OTF_R1(9); break;
            case 5: OTF_R1(11); break;
            case 6: OTF_R1(13); break;
            case 7: OTF_R1(15); break;
            case 8: OTF_R1(17); break;
            case 9: OTF_R1(19); break;
            default: OTF_R1(21); break;
        }
#undef OTF_R1
    } else if (fold) {
        if constexpr (PACKED) {
            const float2* w2 = reinterpret_cast<const float2*>(wsm);
            switch (R) {
                case 2: accumulate_rows_folded<TX, TY, 5>(tt, P, w2, acc); break;
                case 3: accumulate_rows_folded<TX, TY, 7>(tt, P, w2, acc); break;
                case 4: accumulate_rows_folded<TX, TY, 9>(tt, P, w2, acc); break;
                case 5: accumulate_rows_folded<TX, TY, 11>(tt, P, w2, acc); break;
                case 6: accumulate_rows_folded<TX, TY, 13>(tt, P, w2, acc); break;
                case 7: accumulate_rows_folded<TX, TY, 15>(tt, P, w2, acc); break;
                case 8: accumulate_rows_folded<TX, TY, 17>(tt, P, w2, acc); break;
                case 9: accumulate_rows_folded<TX, TY, 19>(tt, P, w2, acc); break;
                default: accumulate_rows_folded<TX, TY, 21>(tt, P, w2, acc); break;
            }
        }
    } else if (PACKED && R >= 1) {
        const float2* w2 = reinterpret_cast<const float2*>(wsm);
        switch (R) {
            case 1: accumulate_rows_packed<TX, TY, 3>(tt, P, w2, acc); break;
            case 2: accumulate_rows_packed<TX, TY, 5>(tt, P, w2, acc); break;
            case 3: accumulate_rows_packed<TX, TY, 7>(tt, P, w2, acc); break;
            case 4: accumulate_rows_packed<TX, TY, 9>(tt, P, w2, acc); break;
            case 5: accumulate_rows_packed<TX, TY, 11>(tt, P, w2, acc); break;
            case 6: accumulate_rows_packed<TX, TY, 13>(tt, P, w2, acc); break;
            case 7: accumulate_rows_packed<TX, TY, 15>(tt, P, w2, acc); break;
            case 8: accumulate_rows_packed<TX, TY, 17>(tt, P, w2, acc); break;
            case 9: accumulate_rows_packed<TX, TY, 19>(tt, P, w2, acc); break;
            default: accumulate_rows_packed<TX, TY, 21>(tt, P, w2, acc); break;
        }
    } else
    switch (R) {
        case 0: {
            const float w = wsm[0];
#pragma unroll
            for (int oy = 0; oy < TY; ++oy)
#pragma unroll
                for (int ox = 0; ox < TX; ++ox) acc[oy][ox] = w * tt[oy * P + ox];
        } break;
        case 1: accumulate_rows<TX, TY, 3>(tt, P, wsm, acc); break;
        case 2: accumulate_rows<TX, TY, 5>(tt, P, wsm, acc); break;
        case 3: accumulate_rows<TX, TY, 7>(tt, P, wsm, acc); break;
        case 4: accumulate_rows<TX, TY, 9>(tt, P, wsm, acc); break;
        case 5: accumulate_rows<TX, TY, 11>(tt, P, wsm, acc); break;
        case 6: accumulate_rows<TX, TY, 13>(tt, P, wsm, acc); break;
        case 7: accumulate_rows<TX, TY, 15>(tt, P, wsm, acc); break;
        case 8: accumulate_rows<TX, TY, 17>(tt, P, wsm, acc); break;
        case 9: accumulate_rows<TX, TY, 19>(tt, P, wsm, acc); break;
        default: accumulate_rows<TX, TY, 21>(tt, P, wsm, acc); break;
    }

    OTF_TRACE(2);  // accumulation done (thread 0's warp)
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
#ifdef OTF_F2D_TRACE
    OTF_TRACE(3);
    if (threadIdx.x == 0) g_f2d_trace[8 * trace_lin + 4] = (unsigned long long)R | ((unsigned long long)(rank1 ? 1 : fold ? 2 : 0) << 8);
#endif
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

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        (void)cudaGetLastError();
    }
    return fn;
}

template <int TX, int TY, int BX, int BY, bool PACKED>
static int launch_blocked(const float* img, int B, int C, int H, int W, const float* kernel, int kernel_batch, int K,
                          const int32_t* scratch, int use_order, float* out, cudaStream_t st) {
    constexpr int TILE_W = TX * BX, TILE_H = TY * BY;
    constexpr int P = TILE_W + 2 * kMaxRA + kPitchPad, ROWS = TILE_H + 2 * kMaxRT;
    const size_t smem = ((size_t)ROWS * P + (PACKED ? 22 * kW2Pitch * 2 : 21 * kWPitch)) * sizeof(float) + sizeof(uint64_t);
    auto kfn = filter2d_kernel<TX, TY, BX, BY, PACKED>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "filter2d smem attribute");
    }
    const dim3 grid(ceil_div(W, TILE_W), ceil_div(H, TILE_H), B * C);
    const int vec_ok = (W % 4 == 0) && (((uintptr_t)out & 15) == 0) && (((uintptr_t)img & 15) == 0);
    // TMA needs 16-byte global strides/base; OTF_F2D_NO_TMA=1 forces the cp.async fill for A/B runs
    static const bool no_tma = getenv("OTF_F2D_NO_TMA") != nullptr;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int use_tma = 0;
    if (vec_ok && !no_tma && encode_tiled_fn()) {
        const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * C};
        const cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
        const cuuint32_t box[3] = {(cuuint32_t)P, (cuuint32_t)ROWS, 1};
        const cuuint32_t estride[3] = {1, 1, 1};
        const CUresult r = encode_tiled_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)img, gdim, gstride, box, estride,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        use_tma = (r == CUDA_SUCCESS);
    }
    // OTF_F2D_PREFETCH=n: L2-prefetch the tile n launch positions ahead (A/B switch; 0 = off)
    static const int pf_ahead = getenv("OTF_F2D_PREFETCH") ? atoi(getenv("OTF_F2D_PREFETCH")) : 0;
    launch_chain(kfn, dim3(grid), dim3(BX * BY), smem, st, tmap, img, kernel, scratch, use_order, out, C, H, W, K, kernel_batch, vec_ok, use_tma,
                 pf_ahead);
    OTF_LAUNCH_CHECK("filter2d_kernel");
    return OTF_OK;
}

}  // namespace otf

extern "C" int64_t otf_filter2d_scratch_words(int kernel_batch) { return (int64_t)otf::scratch_words(kernel_batch); }

extern "C" int otf_filter2d_analyse_f32(const float* const* kernels_host_array, int nsets, int kernel_batch, int K,
                                        int32_t* scratch_dev, void* stream) {
    using namespace otf;
    OTF_REQUIRE(kernels_host_array && scratch_dev, OTF_ERR_BAD_ARG, "filter2d_analyse: null pointer");
    OTF_REQUIRE(K > 0 && (K % 2) == 1 && K <= 2 * kMaxRT + 1, OTF_ERR_BAD_ARG, "filter2d_analyse: K must be odd and <= 21");
    for (int i = 0; i < nsets && i < 4; ++i) OTF_REQUIRE(kernels_host_array[i], OTF_ERR_BAD_ARG, "filter2d_analyse: null kernel set");
    return analyse_sets(kernels_host_array, nsets, kernel_batch, K, scratch_dev, (cudaStream_t)stream);
}

extern "C" int otf_filter2d_f32(const float* img, int B, int C, int H, int W, const float* kernel, int kernel_batch,
                                int K, int32_t* support_dev, int scratch_ready, float* out, void* stream) {
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
    int use_order = 0;
    if (support_dev) {
        // scratch layout (4-byte words): [0,kb) true radius, [kb,2kb) sample order (largest radius first),
        // [2kb,3kb) rank-1 / symmetry flags, then (16-byte aligned) the per-sample staged taps  (otf_filter2d_scratch_words)
        if (!scratch_ready) {
            if (int rc = analyse_sets(&kernel, 1, kernel_batch, K, support_dev, st)) return rc;
        }
        use_order = (kernel_batch == B && B > 1);
    }
    // big planes: 64x64 tiles, 8x4 outputs per thread (128 threads); tiny planes: 32x32 tiles, 4x4 per thread
    const int64_t big_tiles = (int64_t)ceil_div(W, 64) * ceil_div(H, 64) * B * C;
    static const bool scalar = getenv("OTF_F2D_SCALAR") != nullptr;  // A/B switch: scalar FFMA instead of packed FFMA2
    static const bool tall = getenv("OTF_F2D_TALL") != nullptr;      // A/B switch: 4 columns x 8 rows per thread
    if (big_tiles >= 2 * kNumSMs) {
        if (scalar) return launch_blocked<8, 4, 8, 16, false>(img, B, C, H, W, kernel, kernel_batch, K, support_dev, use_order, out, st);
        // Measured alternative (profiles/r02_microbench_f2d_blocks.txt): a 4 x 8 block makes the row loads conflict-free (the
        // eight threads of a quarter-warp read eight consecutive 16-byte chunks; with 8 columns per thread chunks tx and
        // tx + 4 share their banks) and shares each image row's horizontal pass among 8 output rows: rank-1 kernels run 11 %
        // faster, dense and folded ones 4 % slower (twice the weight loads, 28 row iterations instead of 24) — a wash on the
        // default kernel mix, so the 8 x 4 block stays.
        if (tall) return launch_blocked<4, 8, 16, 8, true>(img, B, C, H, W, kernel, kernel_batch, K, support_dev, use_order, out, st);
        return launch_blocked<8, 4, 8, 16, true>(img, B, C, H, W, kernel, kernel_batch, K, support_dev, use_order, out, st);
    }
    // small planes: 64x32 tiles (64 threads) keep the 8x4 register block and halve the tile
    const int64_t mid_tiles = (int64_t)ceil_div(W, 64) * ceil_div(H, 32) * B * C;
    // (few tiles: the thread count, not the FMA pipe, is the limit — at 64x3x64^2 the 64x32 tiling gives 5 warps per SM and
    // each thread walks 24 rows of 8 columns alone; the 32x32 tiling below doubles the threads and halves their work)
    if (mid_tiles >= 4 * kNumSMs && W >= 48) {
        if (scalar) return launch_blocked<8, 4, 8, 8, false>(img, B, C, H, W, kernel, kernel_batch, K, support_dev, use_order, out, st);
        return launch_blocked<8, 4, 8, 8, true>(img, B, C, H, W, kernel, kernel_batch, K, support_dev, use_order, out, st);
    }
    if (scalar) return launch_blocked<4, 4, 8, 8, false>(img, B, C, H, W, kernel, kernel_batch, K, support_dev, use_order, out, st);
    return launch_blocked<4, 4, 8, 8, true>(img, B, C, H, W, kernel, kernel_batch, K, support_dev, use_order, out, st);
}

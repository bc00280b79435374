// Is FFMA2 (fma.rn.f32x2, sm_100+) a way past the one-FFMA-per-issue-slot limit?  Measures FP32 FMA/s
// of a register-only loop written with scalar FFMA and with packed FFMA2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
    unsigned long long da = *reinterpret_cast<unsigned long long*>(&d);
    unsigned long long aa = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long bb = *reinterpret_cast<unsigned long long*>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(da) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&da);
}
template <int MODE>
__global__ void __launch_bounds__(256) rate(float* out, int iters, float a, float b) {
    float2 acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 w = make_float2(a, a), x = make_float2(b, b * 1.0001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) {
                acc[i].x = fmaf(w.x, x.x, acc[i].x);
                acc[i].y = fmaf(w.y, x.y, acc[i].y);
            } else {
                ffma2(acc[i], w, x);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) rate<0><<<148 * 8, 256>>>(out, iters, 0.999f, 1.0f);
            else rate<1><<<148 * 8, 256>>>(out, iters, 0.999f, 1.0f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * 8 * 256 * 32.0 * iters;
        printf("%s: %.3f ms  %.1f TFLOP/s (2 flop per FMA)\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9);
    }
    return 0;
}

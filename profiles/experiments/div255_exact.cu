// Exhaustive check of the three-operation division used by csrc/diffjpeg.cu (div255): for EVERY fp32 x in [0, 255]
// q = x * RN(1/255); r = fma(-q, 255, x); result = fma(r, RN(1/255), q)  must equal the IEEE quotient x / 255.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o div255_exact div255_exact.cu && ./div255_exact
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float div255(float x) {
    const float y = 3.9215688593685627e-03f;
    const float q = __fmul_rn(x, y);
    return fmaf(fmaf(-q, 255.0f, x), y, q);
}
__global__ void check(uint32_t lo, uint32_t hi, unsigned long long* bad, uint32_t* first) {
    for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= hi; i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)i);
        if (__float_as_uint(div255(x)) != __float_as_uint(__fdiv_rn(x, 255.0f))) {
            if (atomicAdd(bad, 1ULL) == 0) *first = (uint32_t)i;
        }
    }
}
int main() {
    unsigned long long* bad; uint32_t* first;
    cudaMallocManaged(&bad, 8); cudaMallocManaged(&first, 4);
    *bad = 0; *first = 0;
    const uint32_t hi = 0x437F0000u;  // 255.0f; every non-negative float up to it (denormals included)
    check<<<148 * 16, 256>>>(0u, hi, bad, first);
    cudaError_t e = cudaDeviceSynchronize();
    printf("checked %u floats in [0, 255]: %llu mismatches (first bits 0x%08x) %s\n", hi + 1, *bad, *first, cudaGetErrorString(e));
    return *bad != 0 || e != cudaSuccess;
}

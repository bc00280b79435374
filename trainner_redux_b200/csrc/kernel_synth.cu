// On-GPU synthesis of the per-sample blur / sinc kernels.
// Replaces the numpy/scipy generators the reference runs in its dataset workers —
// traiNNer/data/degradations.py:22-212 (iso / aniso / generalized / plateau Gaussians) and :472-507
// (circular low-pass sinc via Bessel J1) — and the three (B,21,21) host->device copies that follow.
// The random draws stay on the host (same order as realesrgan_dataset.py:149-206); each kernel is described
// by 8 doubles [type, ksize, sig_x, sig_y, theta, beta, omega_c, pad_to] and evaluated here in fp64
// (64 x 441 values: the cost is nil; fp64 keeps the result within one fp32 ulp of numpy's), normalised,
// zero-padded to 21x21 and rounded to fp32 exactly as the dataset does.
#include <math.h>

#include "otf_common.cuh"

namespace otf {

__global__ void __launch_bounds__(128) synth_kernels_kernel(const double* __restrict__ params, float* __restrict__ out) {
    pdl_enter();
    __shared__ double vals[21 * 21];
    __shared__ double red[4];
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* p = params + (size_t)b * 8;
    const int kind = (int)p[0], k = (int)p[1];
    const double sx = p[2], sy = p[3], theta = p[4], beta = p[5], wc = p[6];
    float* op = out + (size_t)b * 441;
    if (kind == 7 || k < 1 || k > 21) {  // pulse (realesrgan_dataset.py:110-113)
        for (int i = tid; i < 441; i += 128) op[i] = i == 220 ? 1.0f : 0.0f;
        return;
    }
    // inverse of the (rotated) sigma matrix — degradations.py:22-37, :117-125
    double a, bq, c, d;
    if (kind == 0 || kind == 2 || kind == 4) {
        a = d = 1.0 / (sx * sx);
        bq = c = 0.0;
    } else {
        const double ct = cos(theta), st = sin(theta), dx = sx * sx, dy = sy * sy;
        const double s00 = ct * ct * dx + st * st * dy, s01 = ct * st * dx - st * ct * dy, s11 = st * st * dx + ct * ct * dy;
        const double det = s00 * s11 - s01 * s01;
        a = s11 / det; d = s00 / det; bq = c = -s01 / det;
    }
    const double half = (double)(k / 2);  // grid axis: -(k-1)/2 .. (k-1)/2 for odd k (mesh_grid, :40-59)
    double part = 0.0;
    for (int idx = tid; idx < k * k; idx += 128) {
        const int i = idx / k, j = idx - i * k;
        double v;
        if (kind == 6) {
            // :472-507  cutoff*J1(cutoff*r)/(2*pi*r), centre cutoff^2/(4*pi); np.fromfunction: x = row, y = column
            const double cx = (k - 1) * 0.5, dxr = i - cx, dyr = j - cx;
            const double r = sqrt(dxr * dxr + dyr * dyr);
            v = (i == (k - 1) / 2 && j == (k - 1) / 2) ? wc * wc / (4.0 * M_PI) : wc * j1(wc * r) / (2.0 * M_PI * r);
        } else {
            const double x = j - half, y = i - half;  // grid[i][j] = (xx, yy) = (ax[j], ax[i])
            const double q = (x * a + y * c) * x + (x * bq + y * d) * y;
            if (kind <= 1) v = exp(-0.5 * q);
            else if (kind <= 3) v = exp(-0.5 * pow(q, beta));
            else v = 1.0 / (pow(q, beta) + 1.0);
        }
        vals[idx] = v;
        part += v;
    }
    auto block_sum = [&](double x) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = x;
        __syncthreads();
        return red[0] + red[1] + red[2] + red[3];
    };
    const double s1 = block_sum(part);
    double part2 = 0.0;
    const bool twice = kind != 6;  // bivariate_* normalises, then the random_* wrapper normalises again (:258, :313, :369)
    for (int idx = tid; idx < k * k; idx += 128) {
        vals[idx] /= s1;
        part2 += vals[idx];
    }
    const double s2 = twice ? block_sum(part2) : 1.0;
    const int pad = (21 - k) / 2;
    for (int i = tid; i < 441; i += 128) op[i] = 0.0f;
    __syncthreads();
    for (int idx = tid; idx < k * k; idx += 128) {
        const int i = idx / k, j = idx - i * k;
        op[(i + pad) * 21 + (j + pad)] = (float)(twice ? vals[idx] / s2 : vals[idx]);
    }
}

}  // namespace otf

extern "C" int otf_synth_kernels_f32(const double* params_dev, int B, float* out, void* stream) {
    using namespace otf;
    OTF_REQUIRE(params_dev && out && B > 0, OTF_ERR_BAD_ARG, "synth_kernels: bad args");
    launch_chain(synth_kernels_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, params_dev, out);
    OTF_LAUNCH_CHECK("synth_kernels_kernel");
    return OTF_OK;
}

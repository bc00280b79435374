// Per-CTA phase timeline of filter2d_kernel (experiment, not part of the library):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DOTF_F2D_TRACE -I../../include -o f2d_trace f2d_trace.cu
//   ./f2d_trace [K (true size inside 21x21, 0 = mix 7..21)] [kind: 0 dense, 1 rank-1, 2 symmetric]
// Prints, from %globaltimer stamps written by thread 0 of every CTA: kernel span, per-phase mean durations
// (load = entry -> tile ready, compute, store), and how many CTAs are in each phase at sampled instants.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../trainner_redux_b200/csrc/filter2d.cu"
#include <cstdarg>
namespace otf {
bool pdl_enabled() { return false; }
static char g_err[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
int cuda_fail(cudaError_t e, const char* what) { set_error("%s: %s", what, cudaGetErrorString(e)); return -100; }
}  // namespace otf
extern "C" const char* otf_last_error() { return otf::g_err; }

int main(int argc, char** argv) {
    const int Ktrue = argc > 1 ? atoi(argv[1]) : 0, kind = argc > 2 ? atoi(argv[2]) : 0;
    const int B = 64, C = 3, H = 256, W = 256, K = 21;
    std::vector<float> himg((size_t)B * C * H * W), hk((size_t)B * K * K, 0.f);
    srand(1);
    for (auto& v : himg) v = rand() / (float)RAND_MAX;
    for (int b = 0; b < B; ++b) {
        const int kt = Ktrue ? Ktrue : 7 + 2 * (rand() % 8), r = kt / 2;
        std::vector<float> u(kt), v(kt);
        for (int i = 0; i < kt; ++i) { u[i] = expf(-0.1f * (i - r) * (i - r)); v[i] = u[i]; }
        double sum = 0;
        for (int i = 0; i < kt; ++i)
            for (int j = 0; j < kt; ++j) {
                float w = kind == 1 ? u[i] * v[j] : kind == 2 ? expf(-0.05f * powf((i - r) * (i - r) + (j - r) * (j - r), 0.8f)) : rand() / (float)RAND_MAX;
                hk[(size_t)b * K * K + (10 - r + i) * K + (10 - r + j)] = w;
                sum += w;
            }
        for (int i = 0; i < K * K; ++i) hk[(size_t)b * K * K + i] /= (float)sum;
    }
    float *img, *out, *kern;
    int32_t* scratch;
    cudaMalloc(&img, himg.size() * 4 * 4);  // 4 rotating inputs
    cudaMalloc(&out, himg.size() * 4);
    cudaMalloc(&kern, hk.size() * 4);
    cudaMalloc(&scratch, otf_filter2d_scratch_words(B) * 4);
    for (int i = 0; i < 4; ++i) cudaMemcpy(img + i * himg.size(), himg.data(), himg.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(kern, hk.data(), hk.size() * 4, cudaMemcpyHostToDevice);
    const float* kp = kern;
    if (otf_filter2d_analyse_f32(&kp, 1, B, K, scratch, nullptr)) { printf("analyse: %s\n", otf_last_error()); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(e0);
        if (otf_filter2d_f32(img + (it % 4) * himg.size(), B, C, H, W, kern, B, K, scratch, 1, out, nullptr)) { printf("filter2d: %s\n", otf_last_error()); return 1; }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 1; }
    const int n = 4 * 4 * B * C;
    std::vector<unsigned long long> tr(8 * 8192);
    cudaMemcpyFromSymbol(tr.data(), otf::g_f2d_trace, tr.size() * 8);
    unsigned long long t0 = ~0ull, t1 = 0;
    for (int i = 0; i < n; ++i) { t0 = std::min(t0, tr[8 * i]); t1 = std::max(t1, tr[8 * i + 3]); }
    double ld = 0, cp = 0, st = 0;
    for (int i = 0; i < n; ++i) { ld += tr[8 * i + 1] - tr[8 * i]; cp += tr[8 * i + 2] - tr[8 * i + 1]; st += tr[8 * i + 3] - tr[8 * i + 2]; }
    printf("K=%d kind=%d: event %.1f us, CTA span %.1f us, %d CTAs; mean per CTA: load %.2f us, compute %.2f us, store %.2f us\n", Ktrue, kind,
           ms * 1e3, (t1 - t0) * 1e-3, n, ld / n * 1e-3, cp / n * 1e-3, st / n * 1e-3);
    const int S = 24;
    printf("t(us)  loading computing storing   started\n");
    for (int s = 0; s < S; ++s) {
        const unsigned long long t = t0 + (t1 - t0) * (2 * s + 1) / (2 * S);
        int a = 0, b = 0, c = 0, d = 0;
        for (int i = 0; i < n; ++i) {
            const unsigned long long* q = &tr[8 * i];
            if (q[0] <= t) ++d;
            if (q[0] <= t && t < q[1]) ++a; else if (q[1] <= t && t < q[2]) ++b; else if (q[2] <= t && t < q[3]) ++c;
        }
        printf("%5.1f  %7d %9d %7d %9d\n", (t - t0) * 1e-3, a, b, c, d);
    }
    // per-SM CTA count spread
    int per_sm[256] = {0};
    for (int i = 0; i < n; ++i) per_sm[tr[8 * i + 5] & 255]++;
    int mn = 1 << 30, mx = 0;
    for (int i = 0; i < 148; ++i) { mn = std::min(mn, per_sm[i]); mx = std::max(mx, per_sm[i]); }
    printf("CTAs per SM: min %d max %d\n", mn, mx);
    return 0;
}

// Native stage executor (host code): launches a whole degradation chain from one C call.
// See include/otf_b200.h "native stage executor".  Nothing here touches pixels: every stage is one
// of the library's own entry points, called in sequence on the caller's stream, with the image
// threaded through two ping-pong buffers carved out of the caller's workspace.
// Adjacent stages that have a fused kernel run as ONE launch:
//   resize + Gaussian noise            -> otf_resize_gauss_f32          (resize.cu, noise epilogue)
//   DiffJPEG(+8-bit lattice) + crop    -> otf_diffjpeg_crop_pair_f32    (diffjpeg.cu, crop tail)
//   clamp/round + crop                 -> otf_crop_pair_f32(lq_round8)  (pointwise.cu)
// each bit-identical to the two launches it replaces (tests/test_chain_native_gpu.py); OTF_FUSE=0 turns them off.
#include <stdlib.h>

#include "otf_common.cuh"

namespace otf {

static thread_local int g_last_launches = 0;

static bool fuse_enabled() {
    const char* e = getenv("OTF_FUSE");
    return !(e && e[0] == '0');
}

// resize followed by a Gaussian stage the resize epilogue can draw itself (no injected field, plain / clip tail)
static bool fusable_resize_gauss(const OtfStage& r, const OtfStage& g) {
    return r.op == OTF_OP_RESIZE && g.op == OTF_OP_GAUSS && !r.dst && g.p0 && !g.p2 && !g.p3 &&
           !(g.flags & (OTF_NOISE_ROUNDS | OTF_NOISE_FIELD_ONLY | OTF_NOISE_RAW_FIELD));
}

static inline int64_t align_up(int64_t v, int64_t a = 256) { return (v + a - 1) / a * a; }

// What one pass over the stage list learns: extents after every stage and the scratch each needs.
struct Layout {
    int64_t image_bytes = 0;    // each of the two ping-pong buffers
    int64_t analysis_bytes = 0; // shared filter2d analysis (lives for the whole chain)
    int64_t scratch_bytes = 0;  // per-stage scratch (stream-ordered, so one region of the largest size)
    int final_h = 0, final_w = 0;
};

static int plan_layout(int B, int C, int H, int W, const OtfStage* st, int n, Layout* L) {
    OTF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, OTF_ERR_BAD_ARG, "run_stages: bad image extents");
    OTF_REQUIRE(st && n > 0 && n <= 64, OTF_ERR_BAD_ARG, "run_stages: need 1..64 stages (got %d)", n);
    int h = H, w = W;
    for (int i = 0; i < n; ++i) {
        const OtfStage& s = st[i];
        int64_t scratch = 0;
        bool makes_image = true;
        switch (s.op) {
            case OTF_OP_ANALYSE:
                OTF_REQUIRE(s.n >= 1 && s.n <= 4, OTF_ERR_BAD_ARG, "run_stages[%d]: analyse takes 1..4 kernel tensors", i);
                L->analysis_bytes = align_up((int64_t)s.n * otf_filter2d_scratch_words(s.kb) * 4);
                makes_image = false;
                break;
            case OTF_OP_FILTER2D:
                if (s.n < 0) scratch = otf_filter2d_scratch_words(s.kb) * 4;
                break;
            case OTF_OP_USM:
                scratch = otf_usm_workspace_bytes(B * C, h, w);
                break;
            case OTF_OP_SEPCONV:
            case OTF_OP_GAUSS:
            case OTF_OP_JPEG:
            case OTF_OP_CLAMP_ROUND:
            case OTF_OP_WARP:
            case OTF_OP_GAIN:
            case OTF_OP_SENSOR:
            case OTF_OP_DEMOSAIC:
            case OTF_OP_TRUNC8:
                break;
            case OTF_OP_TAPS_ZERO: {
                OTF_REQUIRE(s.K > 0 && s.p0, OTF_ERR_BAD_ARG, "run_stages[%d]: taps_zero needs a host kernel", i);
                const int pad = s.K / 2;  // F.conv2d(padding=K//2): an even K grows the image by one
                h = h + 2 * pad - s.K + 1;
                w = w + 2 * pad - s.K + 1;
                break;
            }
            case OTF_OP_POISSON:
                scratch = (int64_t)B * 16 * 4;
                break;
            case OTF_OP_LIBJPEG:
                OTF_REQUIRE(C == 3, OTF_ERR_BAD_ARG, "run_stages[%d]: the JPEG round needs 3 channels", i);
                scratch = otf_libjpeg_workspace_bytes(B, h, w);
                break;
            case OTF_OP_RESIZE:
                OTF_REQUIRE(s.oh > 0 && s.ow > 0, OTF_ERR_BAD_ARG, "run_stages[%d]: resize to (%d, %d)", i, s.oh, s.ow);
                if (!s.p0) scratch = otf_resize_workspace_bytes(h, w, s.oh, s.ow, s.mode);
                OTF_REQUIRE(scratch >= 0, OTF_ERR_BAD_ARG, "run_stages[%d]: resize mode %d or extents (%d, %d) -> (%d, %d) not supported", i, s.mode, h, w, s.oh, s.ow);
                h = s.oh;
                w = s.ow;
                break;
            case OTF_OP_CROP_PAIR:
                OTF_REQUIRE(i == n - 1, OTF_ERR_BAD_ARG, "run_stages[%d]: crop_pair must be the last stage", i);
                makes_image = false;
                break;
            default:
                OTF_REQUIRE(false, OTF_ERR_BAD_ARG, "run_stages[%d]: unknown op %d", i, s.op);
        }
        if (makes_image && !s.dst) {
            const int64_t bytes = align_up((int64_t)B * C * h * w * 4);
            if (bytes > L->image_bytes) L->image_bytes = bytes;
        }
        scratch = align_up(scratch);
        if (scratch > L->scratch_bytes) L->scratch_bytes = scratch;
    }
    L->final_h = h;
    L->final_w = w;
    return OTF_OK;
}

}  // namespace otf

extern "C" int64_t otf_run_stages_workspace_bytes(int B, int C, int H, int W, const OtfStage* stages, int nstages) {
    otf::Layout L;
    if (int rc = otf::plan_layout(B, C, H, W, stages, nstages, &L)) return rc;
    return 2 * L.image_bytes + L.analysis_bytes + L.scratch_bytes + 256;
}

extern "C" int otf_run_stages_f32(const float* img, int B, int C, int H, int W, const OtfStage* stages, int nstages,
                                  void* workspace_dev, int64_t workspace_bytes, int* final_h, int* final_w, void* stream) {
    using namespace otf;
    Layout L;
    if (int rc = plan_layout(B, C, H, W, stages, nstages, &L)) return rc;
    const int64_t need = 2 * L.image_bytes + L.analysis_bytes + L.scratch_bytes;
    OTF_REQUIRE(img, OTF_ERR_BAD_ARG, "run_stages: null image");
    OTF_REQUIRE(need == 0 || workspace_dev, OTF_ERR_BAD_ARG, "run_stages: null workspace");
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace_dev) + 255) & ~(uintptr_t)255);
    OTF_REQUIRE(need == 0 || (base - reinterpret_cast<char*>(workspace_dev)) + need <= workspace_bytes, OTF_ERR_BAD_ARG,
                "run_stages: workspace of %lld bytes is too small (need %lld)", (long long)workspace_bytes, (long long)need + 256);
    float* pong[2] = {reinterpret_cast<float*>(base), reinterpret_cast<float*>(base + L.image_bytes)};
    int32_t* analysis = reinterpret_cast<int32_t*>(base + 2 * L.image_bytes);
    void* scratch = base + 2 * L.image_bytes + L.analysis_bytes;
    bool analysed = false;
    int analysed_kb = 0, analysed_sets = 0;

    const bool fuse = fuse_enabled();
    int launches = 0;
    g_last_launches = 0;
    const float* cur = img;
    int h = H, w = W, which = 0;
    for (int i = 0; i < nstages; ++i) {
        const OtfStage& s = stages[i];
        float* out = s.dst ? reinterpret_cast<float*>(s.dst) : pong[which];
        int rc = OTF_OK;
        bool makes_image = true;
        const OtfStage* nx = i + 1 < nstages ? &stages[i + 1] : nullptr;
        if (fuse && nx && fusable_resize_gauss(s, *nx)) {  // resize + noise + clamp: one launch, one image write
            float* fout = nx->dst ? reinterpret_cast<float*>(nx->dst) : pong[which];
            void* tables = s.p0 ? const_cast<void*>(s.p0) : scratch;
            const int ready = (s.p0 && (s.flags & 2)) ? 1 : 0;
            rc = otf_resize_gauss_f32(cur, B, C, h, w, fout, s.oh, s.ow, s.mode, s.flags & 1, tables,
                                      otf_resize_workspace_bytes(h, w, s.oh, s.ow, s.mode), ready, (const float*)nx->p0,
                                      (const float*)nx->p1, nx->seed, nx->offset, (const uint64_t*)nx->p4, nx->flags, stream);
            if (rc != OTF_OK) return rc;
            launches += ready ? 1 : 2;
            h = s.oh;
            w = s.ow;
            cur = fout;
            if (!nx->dst) which ^= 1;
            ++i;
            continue;
        }
        if (fuse && nx && nx->op == OTF_OP_CROP_PAIR && !s.dst && C == 3 && s.op == OTF_OP_JPEG && ((s.flags >> 3) & 1) &&
            (nx->n * nx->mode) % 4 == 0 && (reinterpret_cast<uintptr_t>(nx->p1) & 15) == 0) {
            // the last codec pass stores only the crop window; the GT window rides in the same launch
            rc = otf_diffjpeg_crop_pair_f32(cur, B, h, w, (const float*)s.p0, s.f0, s.flags & 1, (s.flags >> 1) & 1, (s.flags >> 2) & 1,
                                            (const float*)nx->p0, H, W, nx->oh, nx->ow, (const int32_t*)nx->p4, nx->n, nx->mode,
                                            (float*)const_cast<void*>(nx->p1), (float*)const_cast<void*>(nx->p2), stream);
            if (rc != OTF_OK) return rc;
            launches += 1;
            ++i;
            continue;
        }
        if (fuse && nx && nx->op == OTF_OP_CROP_PAIR && !s.dst && s.op == OTF_OP_CLAMP_ROUND) {
            rc = otf_crop_pair_f32((const float*)nx->p0, B * C, H, W, cur, h, w, nx->oh, nx->ow, (const int32_t*)nx->p4, nx->n, nx->mode, 1,
                                   (float*)const_cast<void*>(nx->p1), (float*)const_cast<void*>(nx->p2), stream);
            if (rc != OTF_OK) return rc;
            launches += 1;
            ++i;
            continue;
        }
        switch (s.op) {
            case OTF_OP_ANALYSE: {
                const float* sets[4] = {(const float*)s.p0, (const float*)s.p1, (const float*)s.p2, (const float*)s.p3};
                rc = otf_filter2d_analyse_f32(sets, s.n, s.kb, s.K, analysis, stream);
                launches += 1;
                analysed = true;
                analysed_kb = s.kb;
                analysed_sets = s.n;
                makes_image = false;
                break;
            }
            case OTF_OP_FILTER2D:
                if (s.n >= 0) {
                    OTF_REQUIRE(analysed && s.n < analysed_sets && s.kb == analysed_kb, OTF_ERR_BAD_ARG,
                                "run_stages[%d]: filter2d refers to analysis set %d that no earlier analyse stage produced", i, s.n);
                    rc = otf_filter2d_f32(cur, B, C, h, w, (const float*)s.p0, s.kb, s.K,
                                          analysis + (int64_t)s.n * otf_filter2d_scratch_words(s.kb), 1, out, stream);
                    launches += 1;
                } else {
                    rc = otf_filter2d_f32(cur, B, C, h, w, (const float*)s.p0, s.kb, s.K, (int32_t*)scratch, 0, out, stream);
                    launches += s.K > 21 ? 1 : 2;
                }
                break;
            case OTF_OP_USM:
                rc = otf_usm_sharp_f32(cur, B * C, h, w, (const float*)s.p0, s.n, s.f0, s.f1, scratch,
                                       otf_usm_workspace_bytes(B * C, h, w), out, stream);
                launches += otf_usm_launch_count(B * C, h, w);
                break;
            case OTF_OP_SEPCONV:
                rc = otf_sepconv_reflect_f32(cur, B * C, h, w, (const float*)s.p0, s.n, s.mode, out, stream);
                launches += 1;
                break;
            case OTF_OP_RESIZE: {
                void* tables = s.p0 ? const_cast<void*>(s.p0) : scratch;
                rc = otf_resize_f32(cur, B * C, h, w, out, s.oh, s.ow, s.mode, s.flags & 1, tables,
                                    otf_resize_workspace_bytes(h, w, s.oh, s.ow, s.mode), (s.p0 && (s.flags & 2)) ? 1 : 0, stream);
                launches += (s.p0 && (s.flags & 2)) ? 1 : 2;
                h = s.oh;
                w = s.ow;
                break;
            }
            case OTF_OP_GAUSS:
                rc = otf_gaussian_noise_f32(cur, B, C, h, w, (const float*)s.p0, (const float*)s.p1, (const float*)s.p2,
                                            (const float*)s.p3, s.seed, s.offset, (const uint64_t*)s.p4, s.flags, out, stream);
                launches += 1;
                break;
            case OTF_OP_POISSON:
                // flags bit 3: p2 carries the universal CDF tables (otf_poisson_build_tables) instead of injected counts
                rc = otf_poisson_noise_f32(cur, B, C, h, w, (const float*)s.p0, (const float*)s.p1,
                                           (s.flags & 8) ? nullptr : (const float*)s.p2, (s.flags & 8) ? nullptr : (const float*)s.p3,
                                           s.seed, s.offset, (const uint64_t*)s.p4, s.flags & 7, (uint32_t*)scratch,
                                           (s.flags & 8) ? s.p2 : nullptr,
                                           nullptr, nullptr, nullptr, out, stream);
                launches += 2;
                break;
            case OTF_OP_JPEG:
                OTF_REQUIRE(C == 3, OTF_ERR_BAD_ARG, "run_stages[%d]: DiffJPEG needs 3 channels", i);
                rc = otf_diffjpeg_f32(cur, B, h, w, (const float*)s.p0, s.f0, s.flags & 1, (s.flags >> 1) & 1, (s.flags >> 2) & 1,
                                      (s.flags >> 3) & 1, out, stream);
                launches += 1;
                break;
            case OTF_OP_CLAMP_ROUND:
                rc = otf_clamp_round_f32(cur, (int64_t)B * C * h * w, out, stream);
                launches += 1;
                break;
            case OTF_OP_LIBJPEG:
                rc = otf_libjpeg_roundtrip_f32(cur, B, h, w, s.n, scratch, otf_libjpeg_workspace_bytes(B, h, w), out, stream);
                launches += 2;
                break;
            case OTF_OP_WARP:
                rc = otf_warp_f32(cur, B, C, h, w, s.mode, s.f0, out, stream);
                launches += 1;
                break;
            case OTF_OP_TAPS_ZERO: {
                rc = otf_taps_zero_f32(cur, B * C, h, w, s.K, (const float*)s.p0, s.flags, s.f0, out, stream);
                const int pad = s.K / 2;
                h = h + 2 * pad - s.K + 1;
                w = w + 2 * pad - s.K + 1;
                launches += 1;
                break;
            }
            case OTF_OP_GAIN:
                rc = otf_channel_gain_f32(cur, B, C, (int64_t)h * w, s.f0, s.f1, s.f2, s.flags & 1, out, stream);
                launches += 1;
                break;
            case OTF_OP_SENSOR:
                rc = otf_sensor_noise_f32(cur, (int64_t)B * C * h * w, s.f0, (const float*)s.p0, s.seed, s.offset, out, stream);
                launches += 1;
                break;
            case OTF_OP_DEMOSAIC:
                OTF_REQUIRE(C == 3, OTF_ERR_BAD_ARG, "run_stages[%d]: demosaic needs 3 channels", i);
                rc = otf_demosaic_f32(cur, B, h, w, out, stream);
                launches += 1;
                break;
            case OTF_OP_TRUNC8:
                rc = otf_trunc8_f32(cur, (int64_t)B * C * h * w, out, stream);
                launches += 1;
                break;
            case OTF_OP_CROP_PAIR:
                rc = otf_crop_pair_f32((const float*)s.p0, B * C, H, W, cur, h, w, s.oh, s.ow, (const int32_t*)s.p4, s.n, s.mode, 0,
                                       (float*)const_cast<void*>(s.p1), (float*)const_cast<void*>(s.p2), stream);
                launches += 1;
                makes_image = false;
                break;
        }
        if (rc != OTF_OK) return rc;  // the entry point has set the message
        g_last_launches = launches;
        if (makes_image) {
            cur = out;
            if (!s.dst) which ^= 1;
        }
    }
    g_last_launches = launches;
    if (final_h) *final_h = h;
    if (final_w) *final_w = w;
    return OTF_OK;
}

extern "C" int otf_run_stages_launches(void) { return otf::g_last_launches; }

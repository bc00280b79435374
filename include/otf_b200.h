/*
 * otf_b200.h — C ABI of libotf_b200.so: B200 (sm_100a) kernels for the on-the-fly
 * Real-ESRGAN-style second-order degradation path of traiNNer-redux.
 *
 * This is the drop-in boundary.  The reference has no FFI for
 * this path — it is pure Python calling ATen — so each entry point below names
 * the reference *function* it replaces (file:line relative to the reference
 * tree).  A maintainer binds these with ctypes; the shipped
 * binding is trainner_redux_b200/_lib.py.
 *
 * Conventions
 *  - every pointer named *_dev / img / out is a DEVICE pointer to fp32 data in
 *    dense NCHW order unless stated; the caller owns all memory;
 *  - nothing here allocates persistent device memory, synchronises the device
 *    or the stream, or reads results back to the host;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL =
 *    the legacy default stream);
 *  - return value: 0 (OTF_OK) or a negative OTF_ERR_* code; the message for the
 *    last error on the calling thread is available from otf_last_error();
 *  - the library is reentrant; it keeps no mutable global state except the
 *    per-thread error string and a lazily resolved driver entry point.
 */
#ifndef OTF_B200_H_
#define OTF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OTF_ABI_VERSION 6  /* 6: otf_f32_to_u8; 5: otf_libjpeg_roundtrip_f32 + OTF_OP_LIBJPEG; 4: prefetcher upload step (otf_upload_async + events); 3: OtfStage.f2, fork-extra ops of the stage executor, otf_usm_launch_count */

enum {
    OTF_OK = 0,
    OTF_ERR_BAD_ARG = -1,     /* null pointer, non-positive extent, even kernel size ... */
    OTF_ERR_UNSUPPORTED = -2, /* shape outside what the kernels are built for          */
    OTF_ERR_CUDA = -3,        /* a CUDA runtime call or launch failed                   */
    OTF_ERR_WORKSPACE = -4    /* caller-provided workspace too small                    */
};

/* resize modes — traiNNer/data/degradations.py:1004-1021 (`resize_pt`) */
enum {
    OTF_RESIZE_BILINEAR_AA = 0, /* F.interpolate(mode="bilinear", antialias=True)      */
    OTF_RESIZE_BICUBIC_AA = 1,  /* F.interpolate(mode="bicubic",  antialias=True), a=-0.5 */
    OTF_RESIZE_AREA = 2,        /* adaptive average pooling                             */
    OTF_RESIZE_NEAREST_EXACT = 3,
    OTF_RESIZE_BICUBIC = 4,     /* non-antialiased bicubic, a=-0.75 (tail of "lanczos") */
    OTF_RESIZE_NEAREST = 5,     /* legacy F.interpolate(mode="nearest"): floor(o * in/out) — the fork's aliasing stage */
    OTF_RESIZE_LANCZOS = 6      /* the whole "lanczos" mode (:961-1001) as one pass: Lanczos-3 prefilter (reflect padding) on the
                                   shrinking axes composed with the plain bicubic sample; prefilter radius <= 62 */
};

/* fork extras — traiNNer/models/paragon_otf_degradations.py */
enum {
    OTF_WARP_LENS = 0,    /* apply_lens_distortion  :297-342  (p0 = strength)               */
    OTF_WARP_SHUTTER = 1, /* apply_rolling_shutter  :417-455  (p0 = strength*H/W)           */
    OTF_WARP_CHROMA = 2   /* apply_chromatic_aberration :485-523 (R x1.001, B x0.999, clamp) */
};
enum { OTF_TAPS_NONE = 0, OTF_TAPS_OVERSHARPEN = 1 };
#define OTF_MAX_TAPS 160

/* flags for the noise tails: degradations.py:626-632 */
enum {
    OTF_NOISE_CLIP = 1,
    OTF_NOISE_ROUNDS = 2,
    OTF_NOISE_FIELD_ONLY = 4, /* write the noise field itself (generate_*_noise_pt), no add/tail */
    OTF_NOISE_RAW_FIELD = 16  /* Gaussian entry only: noise_color_dev holds the FINISHED noise field (what generate_*_noise_pt
                                 returned); out = tail(img + field) — degradations.py:625 / :833 */
};

int otf_abi_version(void);
const char* otf_last_error(void);
/* Compute capability of the current device as major*10+minor (100 on B200), or <0. */
int otf_device_cc(void);

/* ---- a1: filter2d — traiNNer/utils/img_process_util.py:8-32 -------------------
 * out[b,c,y,x] = sum_{i,j<K} reflect_pad(img)[b,c,y+i,x+j] * kernel[kb,i,j],
 * kb = b (kernel_batch == B) or 0 (kernel_batch == 1).  K odd, K//2 < min(H,W).
 * `scratch_dev` (otf_filter2d_scratch_words(kernel_batch) 4-byte words, 16-byte aligned;
 * may be NULL = analyse nothing, run every kernel at full K) holds the per-kernel analysis
 * the main kernel consumes without any host round trip: true half-width (largest |offset|
 * with a non-zero tap), launch order of the samples (most expensive first, one packed
 * record per position), rank-1 / mirror-symmetry flags, and each sample's taps laid out
 * the way the main kernel keeps them in shared memory (rank-1 factors, or folded / dense
 * paired taps), fetched by one bulk copy per CTA.  With scratch_ready == 0 the call runs
 * the analysis itself (1 small launch); otf_filter2d_analyse_f32 fills the scratch of up to
 * 4 kernel tensors (kernel1, kernel2, sinc_kernel) with ONE launch, laid out
 * back to back (set i at scratch_dev + i * otf_filter2d_scratch_words(kb)), after which
 * the filter calls pass scratch_ready == 1.  K <= 21 runs the register-blocked path
 * specialised on the true support (rank-1 kernels as K + K taps, halo tiles staged by
 * TMA); larger odd K runs the generic path.  `img` and `out` must not alias. */
int64_t otf_filter2d_scratch_words(int kernel_batch);
int otf_filter2d_analyse_f32(const float* const* kernels_host_array, int nsets, int kernel_batch, int K,
                             int32_t* scratch_dev, void* stream);
int otf_filter2d_f32(const float* img, int B, int C, int H, int W,
                     const float* kernel, int kernel_batch, int K,
                     int32_t* scratch_dev, int scratch_ready, float* out, void* stream);

/* ---- 1-D correlation with reflect padding along one axis ----------------------
 * Building block of USMSharp (exactly separable 51x51 Gaussian) and of the
 * Lanczos prefilter in degradations.py:982-998.  `taps_host` is a HOST array of
 * `ntaps` (odd, <= 1023) fp32 taps; axis 0 = vertical (H), 1 = horizontal (W). */
int otf_sepconv_reflect_f32(const float* img, int planes, int H, int W,
                            const float* taps_host, int ntaps, int axis,
                            float* out, void* stream);

/* ---- a2: USMSharp.forward — traiNNer/utils/img_process_util.py:45-55 ----------
 * taps_host: the 1-D Gaussian whose outer product is the module's `kernel`
 * buffer (cv2.getGaussianKernel(radius, sigma) as fp32).  workspace_dev must
 * hold otf_usm_workspace_bytes(planes,H,W) bytes. */
int64_t otf_usm_workspace_bytes(int planes, int H, int W);
int otf_usm_sharp_f32(const float* img, int planes, int H, int W,
                      const float* taps_host, int ntaps, float weight, float threshold,
                      void* workspace_dev, int64_t workspace_bytes, float* out, void* stream);
/* kernels one otf_usm_sharp_f32 call launches: 4 per L2-sized chunk of planes (the passes run chunk by chunk so that
 * the three intermediates stay in L2) */
int otf_usm_launch_count(int planes, int H, int W);

/* ---- a3: resize_pt — traiNNer/data/degradations.py:1004-1021 ------------------
 * Separable resampling with ATen's index/weight rules (SURVEY.md §8a "R"),
 * followed by clamp(0,1) when `clamp01` != 0 (resize_pt always clamps).  Two
 * launches: the per-axis (first index, count, weights) tables into workspace_dev
 * (otf_resize_workspace_bytes bytes, a few KB), then the tiled resampler.  The tables
 * depend on (H, W, OH, OW, mode) only; a caller that kept a workspace filled by an earlier
 * call with the same five values passes tables_ready != 0 and skips the first launch. */
int64_t otf_resize_workspace_bytes(int H, int W, int OH, int OW, int mode);
/* Fill a workspace with the weight tables only (what otf_resize_f32 does first when tables_ready == 0): lets a caller
 * build the tables of a chain before capturing it into a CUDA graph. */
int otf_resize_tables_f32(int H, int W, int OH, int OW, int mode, void* workspace_dev, int64_t workspace_bytes, void* stream);
int otf_resize_f32(const float* img, int planes, int H, int W,
                   float* out, int OH, int OW, int mode, int clamp01,
                   void* workspace_dev, int64_t workspace_bytes, int tables_ready, void* stream);
/* a3 + a4 in ONE launch: the resampler with the Gaussian-noise stage that follows it in the chain
 * (realesrgan_model.py:531-546, :553-562: resize_pt then add_gaussian_noise_pt) as its epilogue — the resampled,
 * clamped pixel gets its noise and the noise tail before it is stored, so the intermediate image never reaches HBM.
 * The Philox positions are otf_gaussian_noise_f32's: the result is BIT-IDENTICAL to otf_resize_f32 followed by
 * otf_gaussian_noise_f32(sigma_dev, gray_dev, seed, offset, offset_dev, noise_flags).  Only the plain / clip tails
 * (noise_flags = 0 or OTF_NOISE_CLIP) and generated fields are fused. */
int otf_resize_gauss_f32(const float* img, int B, int C, int H, int W,
                         float* out, int OH, int OW, int mode, int clamp01,
                         void* workspace_dev, int64_t workspace_bytes, int tables_ready,
                         const float* sigma_dev, const float* gray_dev, uint64_t seed, uint64_t offset,
                         const uint64_t* offset_dev, int noise_flags, void* stream);

/* ---- a4: Gaussian noise — degradations.py:569-633 -----------------------------
 * out = tail(img + mix(N*sigma[b]/255, G*sigma[b]/255, gray[b])).
 * sigma_dev, gray_dev: fp32[B] (gray_dev may be NULL = colour noise only).
 * Either inject the standard-normal fields (noise_color_dev fp32[B,C,H,W],
 * noise_gray_dev fp32[H,W] — ONE field shared by the batch) or pass NULL for
 * both and give a Philox4x32-10 (seed, offset) pair. flags: OTF_NOISE_*.
 * offset_dev (uint64 on the device, may be NULL) is added to `offset` by the kernel: a launch captured in a
 * CUDA graph draws a fresh field on every replay once the caller advances that word (feed_data keeps it in
 * the per-step parameter block it uploads anyway). */
int otf_gaussian_noise_f32(const float* img, int B, int C, int H, int W,
                           const float* sigma_dev, const float* gray_dev,
                           const float* noise_color_dev, const float* noise_gray_dev,
                           uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int flags,
                           float* out, void* stream);
/* Fill out[n] with Philox standard normals / uniforms (distribution tests). */
int otf_philox_normal_f32(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream);
int otf_philox_uniform_f32(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream);

/* ---- a5: Poisson noise — degradations.py:762-842 ------------------------------
 * Two launches behind one call: (1) per-sample 256-bit presence masks of the
 * 8-bit-quantised colour and gray images -> vals = 2^ceil(log2(#distinct));
 * (2) sampling + mixing + tail.  C must be 3.  masks_dev: uint32[B*16] scratch
 * (zeroed by the call).
 * tables_dev (may be NULL): the universal alias tables filled once per device by
 * otf_poisson_build_tables (otf_poisson_tables_bytes() bytes, 16-byte aligned, read-only
 * afterwards).  lambda = (level/255) * vals with vals = 2^v, v <= 8, so only 2304
 * lambdas can ever occur; with the tables and production arguments (no injected
 * counts, no exports) every count is drawn from one Philox uniform and ONE 8-byte
 * table read (Walker / Vose alias method, probabilities right to 2^-24, the
 * resolution of the uniform).  Without them (or with injected counts, exports,
 * fractional gray flags) counts come from a rejection sampler (sequential inversion
 * below lambda 10, PTRS above).
 * vals_out_dev (fp32[B*2]: colour, gray; may be NULL) and lambda_*_dev (may be
 * NULL) export the deterministic half for parity tests.  counts_*_dev (may be
 * NULL) inject pre-drawn Poisson counts. */
int64_t otf_poisson_tables_bytes(void);
int otf_poisson_build_tables(void* tables_dev, void* stream);
int otf_poisson_noise_f32(const float* img, int B, int C, int H, int W,
                          const float* scale_dev, const float* gray_dev,
                          const float* counts_color_dev, const float* counts_gray_dev,
                          uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int flags,
                          uint32_t* masks_dev, const void* tables_dev, float* vals_out_dev,
                          float* lambda_color_dev, float* lambda_gray_dev,
                          float* out, void* stream);
/* out[i] ~ Poisson(lambda[i]) with the library's sampler (distribution tests). */
int otf_philox_poisson_f32(const float* lambda_dev, float* out, int64_t n,
                           uint64_t seed, uint64_t offset, void* stream);

/* ---- a6: DiffJPEG.forward — traiNNer/utils/diffjpeg.py:503-527 ----------------
 * One fused kernel: x255, RGB->YCbCr, 4:2:0, 8x8 DCT, quantise by table*factor[b]
 * with round-half-even (differentiable=0) or round(x)+(x-round(x))^3, dequantise,
 * IDCT, chroma x2, YCbCr->RGB, clamp, /255, crop of the x16 zero padding.
 * factor_dev fp32[B] or NULL to use factor_scalar; the value is a compression factor
 * (already quality_to_factor'ed) or, with factor_is_quality != 0, a raw JPEG quality that the
 * kernel converts itself (saves the separate launch when the caller's tensor need not be mutated).
 * clamp_in != 0 first clamps the input to [0,1] (the call form used by the chain).
 * round8_out != 0 additionally applies clamp(round(x*255),0,255)/255 (a7). */
int otf_quality_to_factor_f32(float* quality_dev, int B, void* stream); /* diffjpeg.py:48-61, in place */
int otf_diffjpeg_f32(const float* img, int B, int H, int W,
                     const float* factor_dev, float factor_scalar, int factor_is_quality,
                     int differentiable, int clamp_in, int round8_out, float* out, void* stream);

/* The chain's usual last three steps in ONE launch: DiffJPEG with the 8-bit lattice on its output (a6 + a7), the LQ crop
 * window stored straight into the dense (B,3,p,p) `lq_out`, and the GT crop window (a8) copied by the CTAs behind the
 * codec's.  Arguments as otf_diffjpeg_f32 (round8_out implied) + otf_crop_pair_f32.  The GT patch must be a multiple of 4
 * pixels wide.  gt_out == NULL: no GT copy (the caller keeps the GT window as a strided view of `gt`, which is what the
 * reference's paired_random_crop returns: transforms.py:124-129). */
int otf_diffjpeg_crop_pair_f32(const float* img, int B, int H, int W,
                               const float* factor_dev, float factor_scalar, int factor_is_quality,
                               int differentiable, int clamp_in,
                               const float* gt, int Hg, int Wg, int top, int left, const int32_t* top_left_dev,
                               int lq_patch, int scale, float* gt_out, float* lq_out, void* stream);

/* ---- a7: clamp/round — traiNNer/models/realesrgan_model.py:616 ----------------
 * out = clamp(round(x*255),0,255)/255, round half to even. In place allowed. */
int otf_clamp_round_f32(const float* x, int64_t n, float* out, void* stream);

/* ---- a8: paired crop — traiNNer/data/transforms.py:124-135 + .contiguous() ----
 * Copies the LQ window (top,left,p,p) and the GT window (top*scale,left*scale,
 * p*scale...) into dense outputs in one launch.  top_left_dev (int32[2] on the device, may be NULL)
 * overrides (top, left): a captured launch then follows the offsets the caller uploads per step (they are
 * clamped to the valid range on the device; the host validates them when it draws them).  lq_round8 != 0 applies
 * the 8-bit lattice of a7 to the LQ window on the way (clamp/round + crop in one launch).  gt_out == NULL skips
 * the GT window (the caller keeps it as a strided view, as the reference's crop does). */
int otf_crop_pair_f32(const float* gt, int planes, int Hg, int Wg,
                      const float* lq, int Hl, int Wl,
                      int top, int left, const int32_t* top_left_dev, int lq_patch, int scale, int lq_round8,
                      float* gt_out, float* lq_out, void* stream);

/* uint8 image -> fp32 / 255 (the host-side normalisation of traiNNer/utils/img_util.py:65-109 `img2tensor`,
 * moved behind a 4x smaller upload). */
int otf_u8_to_f32(const uint8_t* src, int64_t n, float* dst, void* stream);

/* ---- the way back — traiNNer/utils/img_util.py:112-181 (`tensor2img`: `(img * 255.0).round()` as uint8, done on the
 * host after an fp32 read-back) ----  dst[i] = clamp(round(src[i] * 255), 0, 255) (round half to even).  A finished LQ batch
 * lies on the 8-bit lattice (realesrgan_model.py:616), so its read-back moves one byte per value and
 * `u8.float() / 255` restores the fp32 tensor bit for bit.  src 16-byte aligned, dst 4-byte aligned. */
int otf_f32_to_u8(const float* src, int64_t n, uint8_t* dst, void* stream);

/* ---- f3: the "jpeg" round of the fork's unified compression stage — traiNNer/models/paragon_otf_degradations.py:95-158
 * (`_compress_with_format`): `(img.clamp(0,1) * 255).astype(uint8)`, `PIL.Image.save(format="JPEG", quality=int(q))`,
 * `Image.open(...).convert("RGB")`, `/ 255`.  The decoded pixels are those of libjpeg(-turbo)'s baseline round trip with its
 * defaults (4:2:0, Annex-K tables scaled by `quality`, integer "slow" DCT, fancy chroma up-sampling), reproduced BIT FOR
 * BIT on the device (entropy coding is lossless and skipped).  img / out: (B,3,H,W) fp32 dense, any H, W >= 1 (libjpeg's
 * edge expansion included); quality is clamped to 1..100 as jpeg_set_quality does.  workspace: device bytes
 * (otf_libjpeg_workspace_bytes, 16-byte aligned) holding the decoded Y / Cb / Cr planes between the two launches. */
int64_t otf_libjpeg_workspace_bytes(int B, int H, int W);
int otf_libjpeg_roundtrip_f32(const float* img, int B, int H, int W, int quality,
                              void* workspace, int64_t workspace_bytes, float* out, void* stream);

/* ---- side-stream prefetcher — traiNNer/data/prefetch_dataloader.py:418-499 (`CUDAPrefetcher.preload` / `.next`) ----
 * One upload step: [ev_consumed is recorded on consumer_stream and copy_stream waits for it — the destination slots
 * may still be read by what the consumer has issued so far (pass NULL to skip)]; n plain cudaMemcpyAsync host->device
 * on copy_stream (n <= 64; pinned sources for a truly asynchronous copy); ev_ready recorded on copy_stream.  The
 * consumer later orders itself behind the upload with otf_stream_wait_event(consumer_stream, ev_ready) — the
 * `wait_stream` of prefetch_dataloader.py:488-493.  Events are plain CUDA events with timing disabled, owned by the
 * caller; otf_event_query sets *done to 1 when everything recorded before the event has completed, else 0. */
int otf_event_create(void** event);
int otf_event_destroy(void* event);
int otf_event_query(void* event, int* done);
int otf_stream_wait_event(void* stream, void* event);
int otf_upload_async(int n, void* const* dst_dev, const void* const* src_host, const uint64_t* bytes,
                     void* copy_stream, void* consumer_stream, void* ev_consumed, void* ev_ready);
/* The mirror image for a step's result (the `.cpu()` / `.item()` reads of a training loop, e.g. the images
 * traiNNer/models/sr_model.py `get_current_visuals` pulls back): ev_produced is recorded on producer_stream, copy_stream
 * waits for it, one cudaMemcpyAsync device->host (pinned destination), ev_done recorded on copy_stream.
 * otf_event_synchronize blocks the calling host thread until the event has completed. */
int otf_download_async(void* dst_host, const void* src_dev, uint64_t bytes,
                       void* copy_stream, void* producer_stream, void* ev_produced, void* ev_done);
int otf_event_synchronize(void* event);

/* Strided (e.g. channels_last) -> dense NCHW copy; strides in elements. */
int otf_copy_strided_f32(const float* src, const int64_t strides[4],
                         int B, int C, int H, int W, float* dst, void* stream);

/* ---- f2: blur / sinc kernel synthesis — traiNNer/data/degradations.py:22-212, :472-507 ----------
 * params_dev: double[B][8] = [type, ksize, sig_x, sig_y, theta, beta, omega_c, pad_to], type 0 iso,
 * 1 aniso, 2 generalized_iso, 3 generalized_aniso, 4 plateau_iso, 5 plateau_aniso, 6 sinc, 7 pulse
 * (the host draws them in the order of traiNNer/data/realesrgan_dataset.py:149-206).  out: fp32[B][21][21],
 * normalised, zero-padded to 21x21 as the dataset does. */
int otf_synth_kernels_f32(const double* params_dev, int B, float* out, void* stream);

/* ---- a9: pair pool without bulk copies ----------------------
 * Gathers `n` slots: dst[i] = src[idx_host[i]] for slot_elems floats each, and
 * scatters likewise (dst[idx_host[i]] = src[i]).  n <= 512. */
int otf_gather_slots_f32(const float* src, const int32_t* idx_host, int n,
                         int64_t slot_elems, float* dst, void* stream);
int otf_scatter_slots_f32(const float* src, const int32_t* idx_host, int n,
                          int64_t slot_elems, float* dst, void* stream);
/* One pool step (traiNNer/models/realesrgan_model.py:430-447) in one launch: for slot idx_host[i] of BOTH queues the
 * stored pair goes to out_*[i] (dequeue; pass NULL for both while the pool is still filling) and in_*[i] takes its
 * place (enqueue).  lq_elems / gt_elems: floats per slot. */
int otf_pool_exchange_f32(float* queue_lq, float* queue_gt, const int32_t* idx_host, int n,
                          int64_t lq_elems, int64_t gt_elems, const float* in_lq, const float* in_gt,
                          float* out_lq, float* out_gt, void* stream);

/* ---- f4: MoA batch augment on the finished pair — traiNNer/ops/batchaug.py:21-509 ----------------
 * The host draws the augmentation, ratio, permutation and box exactly as the reference does; the
 * resizes inside resizemix / cutblur / downup / up go through otf_resize_f32.
 *
 * otf_mixup_f32 (batchaug.py:150-158): out[b] = lam * img[b] + one_minus_lam * img[perm_host[b]], each
 * product and the sum rounded to fp32 separately as torch does. Not in place. B <= 512.
 *
 * otf_copy_box_f32: dst[b, :, dy:dy+bh, dx:dx+bw] = src[perm_host ? perm_host[b] : b, :, sy:sy+bh, sx:sx+bw]
 * for dense (B, planes_per_sample, Hs, Ws) / (B, planes_per_sample, Hd, Wd) tensors: the box paste of
 * cutmix (:222-227), resizemix (:318-319) and cutblur (:394-401), and the crops of `up` (:476-477).
 * A permuted copy must not be in place (stage the boxes in a workspace first); empty boxes are a no-op. */
int otf_mixup_f32(const float* img, const int32_t* perm_host, int B, int64_t sample_elems,
                  float lam, float one_minus_lam, float* out, void* stream);
int otf_copy_box_f32(const float* src, int Hs, int Ws, int sy, int sx,
                     float* dst, int Hd, int Wd, int dy, int dx, int bh, int bw,
                     int B, int planes_per_sample, const int32_t* perm_host, void* stream);

/* ---- native stage executor: one call launches a whole degradation chain ---------------------------
 * feed_data (traiNNer/models/realesrgan_model.py:455-650) issues its stages one Python call at a time;
 * with a freshly drawn plan per iteration that is ~20 launches of interpreter + binding overhead, more
 * than the GPU time of the chain at the reference's usual batch sizes.  The host side still decides
 * WHAT runs (it mirrors the reference's branches and random draws) and describes it as an array of
 * stages; this entry point launches them back to back on `stream`, threading the image through two
 * ping-pong buffers inside `workspace_dev`.  Every stage is the entry point documented above with the
 * same arithmetic, so a chain run here is bit-identical to the same stages called one by one.
 *
 * Stage i reads the previous stage's output (stage 0 reads `img`, which is never written) and writes
 * either `dst` (caller-owned, dense) or an executor-owned buffer.  Ops and the fields they read:
 *   OTF_OP_ANALYSE      p0..p3 kernel tensors (n = number of them), kb, K     -> shared analysis scratch
 *   OTF_OP_FILTER2D     p0 kernel, kb, K, n = set index in the shared analysis or -1 (analyse alone)
 *   OTF_OP_USM          p0 HOST taps, n = ntaps, f0 weight, f1 threshold
 *   OTF_OP_SEPCONV      p0 HOST taps, n = ntaps, mode = axis (0 vertical, 1 horizontal)
 *   OTF_OP_RESIZE       mode, oh, ow, flags&1 = clamp01, p0 = prebuilt tables or NULL (flags&2 = ready)
 *   OTF_OP_GAUSS        p0 sigma, p1 gray|NULL, p2/p3 injected fields|NULL, seed, offset, p4 = device offset word|NULL,
 *                       flags = OTF_NOISE_*
 *   OTF_OP_POISSON      p0 scale, p1 gray|NULL, p2/p3 injected counts|NULL, seed, offset, p4 as above, flags = OTF_NOISE_*;
 *                       with flags bit 3 (8) set, p2 is the universal CDF table block of otf_poisson_build_tables instead
 *   OTF_OP_JPEG         p0 per-sample factor/quality or NULL (f0 scalar), flags: 1 is_quality, 2 differentiable,
 *                       4 clamp_in, 8 round8_out
 *   OTF_OP_CLAMP_ROUND  -
 *   OTF_OP_CROP_PAIR    terminal: p0 = GT (B,C,H0,W0 of the chain input), oh = top, ow = left, n = lq_patch,
 *                       mode = scale; writes p1 = gt_out, p2 = lq_out; p4 = device int32[2] (top, left) override|NULL
 *   the fork's extra stages (row f3; the entry points documented under "f3: fork extras" below):
 *   OTF_OP_WARP         mode = OTF_WARP_*, f0 = parameter                                     (otf_warp_f32)
 *   OTF_OP_TAPS_ZERO    p0 HOST K x K kernel, K, flags = OTF_TAPS_* epilogue, f0 = strength   (otf_taps_zero_f32; an even K
 *                       grows the image by one row and column)
 *   OTF_OP_GAIN         f0, f1, f2 = per-channel gains, flags&1 = clamp01                      (otf_channel_gain_f32)
 *   OTF_OP_SENSOR       f0 = std, p0 injected N(0,1) field|NULL, seed, offset                  (otf_sensor_noise_f32)
 *   OTF_OP_DEMOSAIC     -  (C must be 3)                                                       (otf_demosaic_f32)
 *   OTF_OP_TRUNC8       -                                                                      (otf_trunc8_f32)
 *   OTF_OP_LIBJPEG      n = quality (C must be 3; two launches)                                (otf_libjpeg_roundtrip_f32)
 * `final_h/final_w` (host, may be NULL) receive the extent of the last image-producing stage. */
enum {
    OTF_OP_ANALYSE = 0, OTF_OP_FILTER2D = 1, OTF_OP_USM = 2, OTF_OP_SEPCONV = 3, OTF_OP_RESIZE = 4,
    OTF_OP_GAUSS = 5, OTF_OP_POISSON = 6, OTF_OP_JPEG = 7, OTF_OP_CLAMP_ROUND = 8, OTF_OP_CROP_PAIR = 9,
    OTF_OP_WARP = 10, OTF_OP_TAPS_ZERO = 11, OTF_OP_GAIN = 12, OTF_OP_SENSOR = 13, OTF_OP_DEMOSAIC = 14, OTF_OP_TRUNC8 = 15,
    OTF_OP_LIBJPEG = 16
};
typedef struct OtfStage {
    int32_t op, mode, oh, ow, n, kb, K, flags;
    float f0, f1;
    uint64_t seed, offset;
    const void* p0;
    const void* p1;
    const void* p2;
    const void* p3;
    void* dst;
    const void* p4; /* per-step device parameters of a captured chain (see the op table) */
    float f2;       /* ABI 3: third scalar (OTF_OP_GAIN) */
    int32_t reserved;
} OtfStage;
int64_t otf_run_stages_workspace_bytes(int B, int C, int H, int W, const OtfStage* stages, int nstages);
int otf_run_stages_f32(const float* img, int B, int C, int H, int W, const OtfStage* stages, int nstages,
                       void* workspace_dev, int64_t workspace_bytes, int* final_h, int* final_w, void* stream);
/* Kernels the calling thread's last otf_run_stages_f32 launched.  The executor fuses adjacent stages where a fused
 * kernel exists (resize + Gaussian noise; DiffJPEG + clamp/round + both crops; clamp/round + both crops; a same-size
 * resize of a clamped image is dropped): fewer launches, bit-identical results.  OTF_FUSE=0 in the environment turns
 * the fusions off (the unfused path the tests compare against). */
int otf_run_stages_launches(void);

/* ---- f3: fork extras — traiNNer/models/paragon_otf_degradations.py:251-572 ----------------
 * otf_warp_f32: analytic sampling grid + F.grid_sample(bilinear, align_corners=False) in one pass.
 *   LENS (:297-342, padding reflection), SHUTTER (:417-455, reflection), CHROMA (:485-523 and
 *   realesrgan_model.py:244-310: per-channel affine scale, zeros padding, clamp(0,1); needs C == 3
 *   to move anything).  ATen's fp32 arithmetic is reproduced operation by operation.
 * otf_taps_zero_f32: F.conv2d(img, k.repeat(C,1,1,1), padding=K//2, groups=C) for a small K x K
 *   kernel given on the HOST (zeros are skipped; at most OTF_MAX_TAPS non-zero taps): motion blur
 *   (:251-273; an even K grows the image by one row/column exactly as the reference does) and, with
 *   epilogue OTF_TAPS_OVERSHARPEN, clamp(img + (img - blur) * strength, 0, 1) (:458-482).
 *   out holds planes x (H + 2*(K/2) - K + 1) x (W + 2*(K/2) - K + 1).
 * otf_channel_gain_f32: out = img * g[c] (+ clamp(0,1)): exposure (:345-362), colour temperature
 *   (:365-394), editing exposure (realesrgan_model.py:596-603).
 * otf_sensor_noise_f32: clamp(img + N * std, 0, 1) (:397-414); N injected (noise_dev) or Philox. */
int otf_warp_f32(const float* img, int B, int C, int H, int W, int mode, float p0, float* out, void* stream);
int otf_taps_zero_f32(const float* img, int planes, int H, int W, int K, const float* kernel_host,
                      int epilogue, float strength, float* out, void* stream);
int otf_channel_gain_f32(const float* img, int B, int C, int64_t hw, float g0, float g1, float g2,
                         int clamp01_out, float* out, void* stream);
int otf_sensor_noise_f32(const float* img, int64_t n, float std, const float* noise_dev,
                         uint64_t seed, uint64_t offset, float* out, void* stream);
/* Bayer mosaic + cv2.demosaicing(COLOR_BAYER_BG2BGR) of apply_demosaicing_artifacts (:526-552), bit for bit:
 * uint8 truncation, one channel per pixel, OpenCV's integer bilinear demosaic with its border rule, / 255.
 * img / out: B x 3 x H x W. */
int otf_demosaic_f32(const float* img, int B, int H, int W, float* out, void* stream);
/* floor(clamp(img,0,1) * 255) / 255: the uint8 truncation in front of every codec round (:114-115). */
int otf_trunc8_f32(const float* img, int64_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OTF_B200_H_ */

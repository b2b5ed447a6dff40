/*
 * gtsfm_b200 — C ABI of the B200-native pairwise deep front-end (SuperPoint -> LightGlue/SuperGlue -> RANSAC).
 *
 * This is the drop-in boundary.  Every entry point is plain C: opaque handle, raw pointers, sizes,
 * an `int` status (0 = ok, <0 = error; text via b2_last_error).  No torch / C++ types cross it.  `*_dev` entry points
 * take DEVICE pointers plus a CUDA stream (passed as void*, i.e. a cudaStream_t / CUstream; NULL = legacy default
 * stream) so PyTorch tensors go in and out through `tensor.data_ptr()`; the `*_host` entry points take HOST pointers
 * and do the H2D / D2H copies themselves (this is what the per-call GTSfM plugins use, mirroring the reference's
 * `.to(device)` / `.cpu().numpy()` inside each call).
 *
 * Reference interface each group replaces (paths relative to the reference repo):
 *   b2_superpoint_*  : gtsfm/frontend/detector_descriptor/superpoint.py:63-93 (detect_and_describe) and the model
 *                      thirdparty/SuperGluePretrainedNetwork/models/superpoint.py:145-202
 *   b2_lightglue_*   : gtsfm/frontend/matcher/lightglue_matcher.py:43-112 and
 *                      thirdparty/LightGlue/lightglue/lightglue.py:474-629
 *   b2_superglue_*   : gtsfm/frontend/matcher/superglue_matcher.py:47-115 and
 *                      thirdparty/SuperGluePretrainedNetwork/models/superglue.py:226-283
 *   b2_ransac_*      : gtsfm/frontend/verifier/ransac.py:52-111 (cv2.findEssentialMat / findFundamentalMat) and
 *                      gtsfm/utils/verification.py:54-96 (cv2.recoverPose)
 *
 * Threading: a handle serialises its own calls with an internal mutex (Dask workers run one thread per process,
 * gtsfm/runner.py:153-155); use one handle per thread/stream for concurrency.
 */
#ifndef GTSFM_B200_H
#define GTSFM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_context b2_context;

/* ---- lifecycle -------------------------------------------------------------------------------------------------- */
int b2_version(void);
/* Creates a context on CUDA device `device`.  Fails (returns <0, *out = NULL) when no sm_100 device is present. */
int b2_create(int device, b2_context** out);
void b2_destroy(b2_context* ctx);
const char* b2_last_error(const b2_context* ctx);
/* Number of kernels this library has launched through `ctx` since creation (bench.py's "gpu_launches"). */
uint64_t b2_launch_count(const b2_context* ctx);
/* Host-to-device bytes actually copied so far by the entry points that keep device copies of their host inputs
 * (b2_lightglue_match_host: feature arrays already uploaded for an earlier pair are not sent again). */
uint64_t b2_h2d_bytes(const b2_context* ctx);
/* Tuning knobs.  "reserve_sms" = n: the persistent kernels (attention, GEMM) launch sm_count - n CTAs, leaving n SMs to
 * kernels of OTHER contexts / streams running concurrently (the batched front-end overlaps pair k's RANSAC with pair
 * k+1's matching; a one-CTA-per-SM kernel that finds an SM busy would otherwise wait for a whole CTA lifetime).
 * "lightglue_batch" = 0..8: pairs per lock-step batch of b2_lightglue_match_batched_dev (0 = 8, the maximum).
 * "force_simt" = 0 | 1: models whose weights are set afterwards run the exact-fp32 SIMT kernels instead of the tcgen05
 * split-fp16 ones (the on-device cross-check of the tensor-core path; tests only).
 * "superpoint_graph" = 0 (default) | 1: launch every kernel of the SuperPoint network directly / replay the ~21 launches as one
 * CUDA graph per (image shape, parameters, buffers) key (measured slower on B200: the path is GPU-bound, not launch-bound).
 * "feature_cache" = 0 | 1: drop every cached device copy of host feature arrays and (0, default) copy on every call like the
 * reference / (1) keep device copies keyed by (host pointer, size) and validated by a hash of the FULL contents, so arrays
 * edited in place are re-sent.  Only pays off for callers that pass the same numpy buffers repeatedly (it does nothing for
 * Dask tasks, which unpickle fresh arrays per call). */
int b2_set_option(b2_context* ctx, const char* name, int64_t value);
/* Live kernel timing for roofline reporting: CUDA events on the launching stream around every launch whose kernel name
 * starts with `kernel_prefix` (e.g. "k_flash_attn"), until b2_profile_stop, which returns the summed device time, the
 * number of such launches and the algorithmic work (FLOP) they performed. */
int b2_profile_start(b2_context* ctx, const char* kernel_prefix);
int b2_profile_stop(b2_context* ctx, double* total_ms, uint64_t* launches, double* work);
/* Copies a named intermediate device buffer of the last call to host (tests only). Returns #floats written or <0. */
int64_t b2_debug_fetch(b2_context* ctx, const char* name, float* host_out, int64_t max_floats);

/* Test-only: C[M,N] = A[M,K] * B[N,K]^T (+ bias) on HOST fp32 buffers.  mode 0 = SIMT fp32 kernel, 1 = tcgen05 split-fp16
 * kernel with fp32 B converted in-kernel, 2 = tcgen05 split-fp16 kernel with pre-split fp16 B.  K must be a multiple of 64. */
int b2_debug_gemm_host(b2_context* ctx, int mode, const float* A, const float* B, const float* bias, float* C, int M, int N,
                       int K);

/* ---- SuperPoint -------------------------------------------------------------------------------------------------- */
/* `blob`: the 24 state-dict tensors in reference order (conv1a.weight, conv1a.bias, conv1b.weight, ... convDb.bias;
 * SURVEY.md Appendix A), OIHW fp32, concatenated.  n_floats must equal 1300865. */
int b2_superpoint_set_weights(b2_context* ctx, const float* host_blob, size_t n_floats);

/* Detect stage on a device image.  `image`: uint8, `channels` in {1,3,4} interleaved (RGB/RGBA are converted with
 * cv2's fixed-point COLOR_RGB2GRAY), row pitch in bytes.  Runs encoder + both heads + NMS + threshold + border
 * removal + ordered compaction.  Keypoints come out in row-major (y, then x) order like torch.nonzero.
 * out_xy: [cap][2] float (x, y); out_score: [cap] float; *out_n (HOST int) = number found (may exceed cap: only the
 * first cap are written).  *out_map_token (HOST, may be NULL) identifies the dense descriptor map this call left in the
 * context.  Synchronises `stream` before returning (the count is data dependent). */
int b2_superpoint_detect_dev(b2_context* ctx, const uint8_t* image, int height, int width, int channels, size_t pitch,
                             float keypoint_threshold, int nms_radius, int border, float* out_xy, float* out_score,
                             int cap, int* out_n, uint64_t* out_map_token, void* stream);
/* Describe stage: bilinear-samples the dense descriptor map identified by `map_token` at `n` (x, y) positions and
 * L2-normalises.  Fails with a "stale feature-map token" error if another detect has run on the context since the token
 * was issued (two images interleaved on one handle can therefore never get each other's descriptors).
 * out_desc: [n][256] float row-major.  Asynchronous on `stream`. */
int b2_superpoint_describe_dev(b2_context* ctx, uint64_t map_token, const float* xy, int n, float* out_desc, void* stream);
/* Fused, self-contained extraction for the batched path: detect -> device top-k (the `max_keypoints` largest responses,
 * ties by lower index, row-major order kept) -> describe, under one lock, one host synchronisation.  out_xy [max_keypoints][2],
 * out_score [max_keypoints], out_desc [max_keypoints][256] are DEVICE buffers; *out_n (HOST) = keypoints written. */
int b2_superpoint_extract_dev(b2_context* ctx, const uint8_t* image, int height, int width, int channels, size_t pitch,
                              float keypoint_threshold, int nms_radius, int border, int max_keypoints, float* out_xy,
                              float* out_score, float* out_desc, int* out_n, void* stream);
/* The same with NO host synchronisation, for callers that keep many images in flight: `out_n_pinned` (page-locked host int,
 * one per image in flight) receives the keypoint count when `stream` gets there; the library's work buffers are reused in
 * stream order.  Call b2_superpoint_finish_dev (stream synchronisation + tensor-core pipeline fault check) before reading the
 * counts or the outputs on the host. */
int b2_superpoint_extract_async_dev(b2_context* ctx, const uint8_t* image, int height, int width, int channels, size_t pitch,
                                    float threshold, int nms_radius, int border, int max_keypoints, float* out_xy,
                                    float* out_score, float* out_desc, int* out_n_pinned, void* stream);
int b2_superpoint_finish_dev(b2_context* ctx, void* stream);
/* Device top-k by score (k largest, ties broken by lower index), result kept in row-major order; writes the selected
 * indices (int32, ascending) and returns count in *out_k (HOST).  For the batched path; the per-call plugin uses the
 * reference's host argpartition to keep its index order. */
int b2_topk_indices_dev(b2_context* ctx, const float* scores, int n, int k, int32_t* out_idx, int* out_k, void* stream);

/* Host-pointer variants (H2D / D2H inside). */
int b2_superpoint_detect_host(b2_context* ctx, const uint8_t* image, int height, int width, int channels,
                              float keypoint_threshold, int nms_radius, int border, float* out_xy, float* out_score,
                              int cap, int* out_n, uint64_t* out_map_token);
int b2_superpoint_describe_host(b2_context* ctx, uint64_t map_token, const float* xy, int n, float* out_desc);

/* ---- image ingest ------------------------------------------------------------------- */
/* The loader's cubic resize (gtsfm/utils/images.py:102-129: cv2.resize(INTER_CUBIC) to the size picked by
 * get_downsampling_factor_per_axis, :150-220) on the device: src / dst are DEVICE uint8 images, `channels` interleaved,
 * src row pitch in bytes, dst dense.  OpenCV's 11-bit fixed-point arithmetic as restated in oracle/images_ref.py.
 * Asynchronous on `stream`; feed dst straight into b2_superpoint_detect_dev / b2_superpoint_extract_dev. */
int b2_image_resize_dev(b2_context* ctx, const uint8_t* src, int height, int width, int channels, size_t pitch, uint8_t* dst,
                        int new_height, int new_width, void* stream);

/* ---- LightGlue --------------------------------------------------------------------------------------------------- */
/* `blob`: packed fp32 tensors in the order documented in gtsfm_b200/weights.py::LIGHTGLUE_ORDER (nn.Linear layout
 * (out,in) row-major as in the checkpoint).  n_floats must equal 11851601 (the 251 parameter tensors; the
 * confidence_thresholds buffer is recomputed). */
int b2_lightglue_set_weights(b2_context* ctx, const float* host_blob, size_t n_floats);

typedef struct b2_lightglue_params {
  /* doubles on purpose: the reference holds these as Python floats and rounds the DERIVED value to float32 at the
   * comparison (e.g. scores > float32(1 - 0.99)), which a float32 field could not reproduce bit-exactly. */
  double depth_confidence; /* 0.95; <= 0 disables early exit   (lightglue.py:330)            */
  double width_confidence; /* 0.99; <= 0 disables pruning      (lightglue.py:331)            */
  double filter_threshold; /* 0.1                              (lightglue.py:332)            */
  int prune_min_kpts;      /* -1 = reference CPU semantics (prune at every layer); 1536 = its CUDA+flash value */
  int fp16_attention;      /* 0 (default): fp32-equivalent attention, what the reference's CPU front-end computes and the fixtures
                            * pin.  1: the reference's CUDA numerics (lightglue.py:116-121: q, k, v cast to half, fp16 flash SDPA,
                            * half result cast back) - one tensor-core product instead of three; descriptors then agree with the
                            * fp32 path to ~1e-3 and match indices are no longer guaranteed bit-identical to the CPU reference. */
} b2_lightglue_params;

/* kp: [n][2] float (x, y) pixels; desc: [n][256] float.  out_matches: [min(n0,n1)][2] int64 rows (idx0, idx1)
 * ascending in idx0; out_scores: [min(n0,n1)] float (exp of the log assignment) or NULL.  *out_k, *out_stop_layer are
 * HOST ints (stop layer is 1-based like the reference's "stop").  Synchronises `stream`. */
int b2_lightglue_match_dev(b2_context* ctx, const float* kp0, const float* desc0, int n0, const float* kp1,
                           const float* desc1, int n1, const b2_lightglue_params* params, int64_t* out_matches,
                           float* out_scores, int* out_k, int* out_stop_layer, void* stream);
int b2_lightglue_match_host(b2_context* ctx, const float* kp0, const float* desc0, int n0, const float* kp1,
                            const float* desc1, int n1, const b2_lightglue_params* params, int64_t* out_matches,
                            float* out_scores, int* out_k, int* out_stop_layer);

/* Batched variant (SURVEY.md 8b "_batched variants taking arrays of pairs"; the seam is
 * gtsfm/frontend/correspondence_generator/det_desc_correspondence_generator.py:65-85, one matcher task per pair).  Up to 8
 * pairs at a time are walked in lock-step: every linear layer, attention call and pruning step is ONE launch over all
 * images of the batch, the per-layer early-exit / pruning counters of all pairs come back in one 128-byte read, pairs
 * that stop early drop out of the later launches.  Results are identical to n_pairs calls of b2_lightglue_match_dev.
 * All pointers inside `pairs` are DEVICE pointers; out_k / out_stop_layer are written on the HOST struct. */
typedef struct b2_lightglue_pair {
  const float* kp0;   /* [n0][2] (x, y) pixels */
  const float* desc0; /* [n0][256] */
  int n0;
  const float* kp1;
  const float* desc1;
  int n1;
  int64_t* out_matches; /* [min(n0, n1)][2] rows (idx0, idx1) ascending in idx0 */
  float* out_scores;    /* [min(n0, n1)] or NULL */
  int out_k;            /* written: number of matches */
  int out_stop_layer;   /* written: 1-based stopping layer */
} b2_lightglue_pair;
int b2_lightglue_match_batched_dev(b2_context* ctx, b2_lightglue_pair* pairs, int n_pairs, const b2_lightglue_params* params,
                                   void* stream);

/* ---- SuperGlue --------------------------------------------------------------------------------------------------- */
/* `blob`: packed fp32 tensors in gtsfm_b200/weights.py::SUPERGLUE_ORDER with eval-mode BatchNorm already folded
 * into the preceding Conv1d by the host loader. */
int b2_superglue_set_weights(b2_context* ctx, const float* host_blob, size_t n_floats);
/* score: [n] keypoint responses; (h, w): image sizes used for keypoint normalisation (superglue.py:63-70).
 * out_matches: [min(n0,n1)][2] uint32 rows (i, matches0[i]) ascending in i. */
int b2_superglue_match_dev(b2_context* ctx, const float* kp0, const float* score0, const float* desc0, int n0, int h0,
                           int w0, const float* kp1, const float* score1, const float* desc1, int n1, int h1, int w1,
                           int sinkhorn_iters, float match_threshold, uint32_t* out_matches, float* out_scores,
                           int* out_k, void* stream);
int b2_superglue_match_host(b2_context* ctx, const float* kp0, const float* score0, const float* desc0, int n0, int h0,
                            int w0, const float* kp1, const float* score1, const float* desc1, int n1, int h1, int w1,
                            int sinkhorn_iters, float match_threshold, uint32_t* out_matches, float* out_scores,
                            int* out_k);

/* ---- RANSAC verifier --------------------------------------------------------------------------------------------- */
typedef struct b2_ransac_params {
  double threshold;   /* inlier threshold: max point-to-epipolar-line style distance (same unit as the points):      */
                      /* thr_px / fx for E on normalised points, thr_px for F on pixels (ransac.py:79,107)           */
  double confidence;  /* 0.999999 (ransac.py:22)                                                                     */
  int max_iters;      /* hypothesis budget (cv2 maxIters): E 1000 (cv2 default, ransac.py:74-81), clamped to 65 536; F: the       */
                      /* reference passes 10^6, the library clamps to 262 144 and stops earlier through the standard        */
                      /* confidence bound.  E only: when the bound says max_iters (<= 4096) samples were NOT enough for      */
                      /* `confidence`, up to 4 x max_iters further samples are drawn (decided on the device).                */
  uint64_t seed;      /* fixed per call => run-to-run identical results (repro test)           */
} b2_ransac_params;

/* x1, x2: [k][2] double matched points (normalised coordinates for E, pixels for F).
 * out_model: 9 doubles row-major (E or F, x2^T M x1 = 0); out_mask: [k] uint8; *out_num_inliers HOST int.
 * For E also recovers the pose by the cheirality vote (cv2.recoverPose semantics): out_R 9 doubles row-major (i2Ri1),
 * out_t 3 doubles (unit i2ti1); may be NULL.  Returns 1 when no model could be estimated (mask all zero). */
int b2_ransac_essential_host(b2_context* ctx, const double* x1, const double* x2, int k, const b2_ransac_params* params,
                             double* out_model, uint8_t* out_mask, int* out_num_inliers, double* out_R, double* out_t);
int b2_ransac_fundamental_host(b2_context* ctx, const double* x1, const double* x2, int k,
                               const b2_ransac_params* params, double* out_model, uint8_t* out_mask,
                               int* out_num_inliers);
/* Device-resident variant for the batched path: kp1 / kp2 are DEVICE [n][2] float pixel coordinates, matches DEVICE
 * [k][2] int64 rows; cal1 / cal2 are HOST {f, u0, v0} of distortion-free pinhole cameras (Cal3Bundler with k1 = k2 = 0,
 * gtsfm/utils/features.py:41-51).  The threshold in `params` is in calibrated units (thr_px / max(f)).  out_mask_dev is a
 * DEVICE [k] uint8 buffer (or NULL); the small outputs are HOST.  Synchronises `stream`. */
int b2_ransac_essential_dev(b2_context* ctx, const float* kp1, const float* kp2, const int64_t* matches, int k,
                            const double* cal1, const double* cal2, const b2_ransac_params* params, double* out_model,
                            uint8_t* out_mask_dev, int* out_num_inliers, double* out_R, double* out_t, void* stream);
/* cv2.recoverPose restated: decompose E, pick (R, t) with most points in front of both cameras. */
int b2_recover_pose_host(b2_context* ctx, const double* E, const double* x1, const double* x2, int k, double* out_R,
                         double* out_t, int* out_num_good);

/* ---- NetVLAD global descriptor (gtsfm/frontend/global_descriptor/netvlad_global_descriptor.py:53-71,
 * thirdparty/hloc/netvlad.py:52-75,163-193) ------------------------------------------------------------------------------------- */
/* blob = 13 x (conv weight OIHW, bias) of VGG16 features[:-2], score_proj [64][512], centers [512][64], whitening weight
 * [4096][32768] and bias [4096], mean [3] (the checkpoint's averageImage): b2_netvlad_blob_floats() floats. */
size_t b2_netvlad_blob_floats(void);
int b2_netvlad_set_weights(b2_context* ctx, const float* blob, size_t n_floats);
/* images: [B][3][H][W] fp32 in [0, 1] (what the reference's batch transform produces), H, W >= 16; out: [B][4096] unit-norm
 * descriptors.  _dev: device pointers, synchronises `stream` before returning; _host: host pointers. */
int b2_netvlad_describe_dev(b2_context* ctx, const float* images, int batch, int height, int width, float* out, void* stream);
int b2_netvlad_describe_host(b2_context* ctx, const float* images, int batch, int height, int width, float* out);

/* ---- retrieval front ---------------------------------------------------------------- */
/* gtsfm/retriever/similarity_retriever.py:86-260: sim = G G^T of the global image descriptors (desc: HOST [n][dim] fp32,
 * dim a multiple of 64), then per query image i its `num_matched` best partners among j > i with sim >= min_score, best
 * first.  out_partners: HOST [n][min(num_matched, n)] int32, -1 = no (further) partner; out_sim: HOST [n][n] or NULL.
 * (Descriptors: b2_netvlad_describe_* above, or any other unit-norm global descriptor.) */
int b2_similarity_pairs_host(b2_context* ctx, const float* desc, int n, int dim, int num_matched, float min_score,
                             int32_t* out_partners, float* out_sim);

#ifdef __cplusplus
}
#endif
#endif /* GTSFM_B200_H */

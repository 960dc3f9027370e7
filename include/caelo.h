/* caelo.h — C ABI of libcaelo_b200.so: the CAE-LO odometry hot path on one B200 (sm_100a).
 *
 * The reference (SRainGit/CAE-LO) has no FFI layer: its boundary is a handful of Python
 * functions plus duck-typed Keras ``model.predict``.  Each entry point
 * below replaces one of those call sites; the reference file:line it stands in for is
 * given with it.  The Python side (caelo_b200/api.py) keeps the reference's names and
 * binds these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C types only; every ``dev`` pointer is a DEVICE pointer (e.g. torch
 *     ``Tensor.data_ptr()``), every ``host`` pointer is host memory;
 *   - the last argument is the CUDA stream (``cudaStream_t`` passed as void*; NULL = default);
 *     calls are asynchronous on that stream unless stated otherwise;
 *   - return value: 0 = CAELO_OK, negative = error (caelo_error_string); never throws,
 *     never allocates or frees caller memory; scratch lives in the ctx and grows on demand;
 *   - one ctx per GPU; a ctx is not thread-safe; distinct ctxs are independent.
 *   - there is NO CPU fallback: without a CUDA device caelo_create fails.
 */
#ifndef CAELO_H
#define CAELO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct caelo_ctx caelo_ctx;

enum {
    CAELO_OK = 0,
    CAELO_ERR_CUDA = -1,           /* a CUDA runtime call failed (see caelo_last_cuda_error) */
    CAELO_ERR_ARG = -2,            /* bad argument (null pointer, size out of range) */
    CAELO_ERR_NO_WEIGHTS = -3,     /* the network weights were not set on this ctx */
    CAELO_ERR_NONBINARY_PATCH = -4,/* encoder input has a value other than 0 or 1 */
    CAELO_ERR_TOO_FEW_VOXELS = -5, /* a voxel list has < 496 entries (sklearn raises too) */
    CAELO_ERR_NO_DEVICE = -6,      /* no usable CUDA device */
    CAELO_ERR_UNSUPPORTED = -7
};

#define CAELO_COUNTER_I8 0
#define CAELO_COUNTER_I32 1
#define CAELO_MAX_TRIALS 500       /* RANSAC4RT maxTrails, Match.py:168 */

int caelo_version(void);
const char *caelo_error_string(int code);
const char *caelo_last_cuda_error(const caelo_ctx *ctx);

int caelo_create(int device_id, caelo_ctx **out);
int caelo_destroy(caelo_ctx *ctx);
int caelo_num_sms(const caelo_ctx *ctx);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches) */
int64_t caelo_launch_count(const caelo_ctx *ctx);

/* Per-launch device timing: while enabled every kernel launch is bracketed by CUDA events on
 * its stream.  caelo_profile_fetch synchronises, writes one "name count total_ms" line per
 * kernel into buf and clears the records (bench.py's roofline numbers come from here). */
int caelo_profile_enable(caelo_ctx *ctx, int on);
int caelo_profile_fetch(caelo_ctx *ctx, char *buf, int buflen);

/* Weights of SphericalRingPCRespondLayer.h5 (HOST pointers, Keras layouts):
 * w1 (3,3,3,32) HWIO, b1 (32), w2 (1,1,32,8), b2 (8).  Replaces keras load_model at
 * Match.py:324, BatchPreprocess.py:168.  Synchronous. */
int caelo_set_respond_weights(caelo_ctx *ctx, const float *w1, const float *b1, const float *w2,
                              const float *b2);

/* Weights of EncoderModel4VoxelPatch.h5 (HOST pointers, Keras layouts): conv kernels DHWIO
 * (3,3,3,1,8) (3,3,3,8,16) (3,3,3,16,32), dense (2048,200) (200,20), all activations tanh.
 * Replaces keras load_model at Match.py:313, PoseEstimation.py:73.  Synchronous. */
int caelo_set_encoder_weights(caelo_ctx *ctx, const float *k1, const float *b1, const float *k2,
                              const float *b2, const float *k3, const float *b3, const float *d1,
                              const float *bd1, const float *d2, const float *bd2);

/* a1 — RespondLayer.predict (SphericalRing.py:405-408, BatchPreprocess.py:110):
 * relu(conv1x1(relu(conv3x3_same(x)+b1))+b2).  ring: dev [B,H,W,3] f32 NHWC; resp: dev [B,H,W,8]. */
int caelo_respond_forward(caelo_ctx *ctx, const float *ring, int B, int H, int W, float *resp,
                          void *stream);

/* a2 — GetKeyPtsByAE (SphericalRing.py:113-291).  resp: dev [B,H,W,8]; ring: dev
 * [B,ring_H,ring_W,ring_C] (ring_C = 3 or 5; the range test norms ALL channels, quirk 3);
 * counter: dev [B,cnt_H,cnt_W] int8 or int32.  All three share the pixel origin.
 * Outputs: kpts dev [B,max_kpts,3] f32 (ascending score), kpix dev [B,max_kpts,2] int64
 * (row,col), n_kpts dev [B] int32 (<= max_kpts; rows beyond it are zero).  max_kpts <= 4095. */
int caelo_select_keypoints(caelo_ctx *ctx, const float *resp, int H, int W, const float *ring,
                           int ring_C, int ring_H, int ring_W, const void *counter,
                           int counter_dtype, int cnt_H, int cnt_W, int B, int max_kpts,
                           float *kpts, int64_t *kpix, int32_t *n_kpts, void *stream);

/* a1+a2 fused — GetKeyPtsFromRawFileName (SphericalRing.py:389-416) without the file I/O:
 * the response image is computed tile by tile in shared memory and never written to HBM.
 * The CNN input is ring[:, 0:H, 0:W, 0:3].  resp_out may be NULL. */
int caelo_respond_select(caelo_ctx *ctx, const float *ring, int ring_C, int ring_H, int ring_W,
                         const void *counter, int counter_dtype, int cnt_H, int cnt_W, int H, int W,
                         int B, int max_kpts, float *kpts, int64_t *kpix, int32_t *n_kpts,
                         float *resp_out, void *stream);

/* a6 — GetPatchesList (Voxel.py:177-216) for F frames at once.  kpts: dev [F,K,3] f32 or
 * f64 (kpts_f64 != 0); n_kpts: dev [F] int32 or NULL (= K everywhere); vox: dev, the three
 * int16 (V,3) lists of every frame concatenated; vox_offsets: HOST [F*3+1] int64 element
 * (row) offsets into vox, frame-major then scale.  Outputs: packed dev [F,3,K,128] uint32
 * (bit i of the patch = flattened index (x*16+y)*16+z, stored rolled by 8 exactly as the
 * reference's negative-index scatter does); patches_f32 dev [F,3,K,16,16,16] or NULL;
 * trunc dev [F,3,K] uint8 or NULL (1 where the 496-neighbour cut removed a voxel). */
int caelo_gather_patches(caelo_ctx *ctx, const void *kpts, int kpts_f64, const int32_t *n_kpts,
                         int F, int K, const int16_t *vox, const int64_t *vox_offsets,
                         uint32_t *packed, float *patches_f32, uint8_t *trunc, void *stream);

/* a3 — PatchEncoder.predict (Match.py:131-133): dev [P,16,16,16(,1)] f32 in {0,1} -> dev [P,20].
 * status: dev int32[1] or NULL; set to CAELO_ERR_NONBINARY_PATCH if an input is not 0/1. */
int caelo_encode_patches(caelo_ctx *ctx, const float *patches, int P, float *feat, int32_t *status,
                         void *stream);
/* a3 on the bit-packed form: packed dev [P,128] uint32 -> feat dev, row p written at
 * feat + p*feat_stride (+ feat_col0), 20 floats.  GetFeaturesFromPatches (Match.py:130-135)
 * is three such calls with feat_stride=60, col0=0/20/40, or one call over [3K] with
 * scale-major packing via caelo_encode_frames. */
int caelo_encode_packed(caelo_ctx *ctx, const uint32_t *packed, int P, float *feat,
                        int feat_stride, int feat_col0, void *stream);
/* packed dev [F,3,K,128] -> feat dev [F,K,60] (np.c_[f0,f1,f2], Match.py:134) */
int caelo_encode_frames(caelo_ctx *ctx, const uint32_t *packed, int F, int K, float *feat,
                        void *stream);

/* a4 — cdist(Codes0,Codes1,'euclidean') + argmin(axis=0) (Match.py:257-258) for P pairs:
 * codes0 dev [P,N,D], codes1 dev [P,M,D] -> pair_idx dev [P,M] int64 (ties -> lowest row,
 * decided on float64 distances as scipy computes them). */
int caelo_nn_match(caelo_ctx *ctx, const float *codes0, const float *codes1, int P, int N, int M,
                   int D, int64_t *pair_idx, void *stream);

/* a5 — one threshold round of RANSAC4RT (Match.py:181-206) for P pairs, all hypotheses in
 * parallel, then the sequential accept/stop rule replayed on the device.
 * pc0 dev [P,N0,3], pc1 dev [P,N,3], pair_idx dev [P,N] int64 (Pairs0 = pc0[pair_idx]) or
 * NULL (then N0 == N and Pairs0 = pc0); sample_idx dev [P,T,4] int32 drawn by the caller
 * from np.random exactly as Match.py:182-184 does; thr dev [P] f32; best_n_in dev [P]
 * int32 (curNumInliers carried across ladder rounds) or NULL (= 0); skip_if_ok dev [P,16] f32 or
 * NULL: pairs whose row there has isSuccess != 0 are skipped and their outputs left untouched (lets
 * the 0.8 / 1.6 ladder rounds be queued without a host round trip).
 * Outputs (dev): result [P,16] f32 = R(9) T(3) isSuccess trials nInliers bestTrial — R,T of
 * the accepted hypothesis (identity/0 if none); inlier_mask [P,N] uint8 of that hypothesis;
 * counts [P,T] int32 per-hypothesis inlier counts or NULL. */
int caelo_ransac_round(caelo_ctx *ctx, const float *pc0, int N0, const float *pc1, int N,
                       const int64_t *pair_idx, const int32_t *sample_idx, int T, const float *thr,
                       const int32_t *best_n_in, const float *skip_if_ok, int P, float *result,
                       uint8_t *inlier_mask, int32_t *counts, void *stream);

/* a5 — the whole RANSAC4RT threshold ladder (Match.py:207-214) + SolveRelativePose's refit (Match.py:273-283) for
 * P pairs in one call, no host round trip: round r (threshold thr_ladder[r], HOST [rounds]) only runs for the pairs
 * that have no model yet.  sample_idx dev [rounds,P,T,4] (caelo_ransac_draw_samples).  Outputs (dev): result
 * [P,16] as in caelo_ransac_round, inlier_mask [P,N] of the accepted hypothesis (all 0 if none), Rt [P,12] refit
 * over the inliers, thr_used [P] threshold of the round that produced the model (last one if none), credible [P]. */
int caelo_ransac_ladder(caelo_ctx *ctx, const float *pc0, int N0, const float *pc1, int N,
                        const int64_t *pair_idx, const int32_t *sample_idx, int T, int rounds,
                        const float *thr_ladder, int P, float *result, uint8_t *inlier_mask, float *Rt,
                        float *thr_used, int32_t *credible, void *stream);

/* a5 — SolveRT (Match.py:138-158) for P independent problems: p0 = pc0[pair_idx] (or pc0),
 * p1 = pc1, restricted to mask != 0 (mask dev [P,N] or NULL = all).  Rt dev [P,12] (R row-major
 * then T), credible dev [P] int32 (+1, -1 = reflection quirk applied, 0 = no points).
 * skip_if_ok as in caelo_ransac_round. */
int caelo_kabsch(caelo_ctx *ctx, const float *pc0, int N0, const float *pc1, int N,
                 const int64_t *pair_idx, const uint8_t *mask, const float *skip_if_ok, int P,
                 float *Rt, int32_t *credible, void *stream);

/* a5 (sample indices) — the index stream RANSAC4RT draws after np.random.seed(seed) (Match.py:182-184:
 * np.random.random((4,)) per trial, int32(u*N)), generated on the device: numpy's legacy MT19937 seeding and
 * random_sample arithmetic, bit for bit.  seeds: HOST int64 [P], each in [0, 2^32) (numpy raises otherwise);
 * rounds_done: threshold-ladder rounds whose T*4 doubles are skipped first (a failed round consumes all T
 * trials); samples: dev int32 [rounds,P,T,4].  Used by the batched pipeline, where pair i of a sequence is
 * seeded with i; the per-pair API (SolveRelativePose) keeps drawing from the caller's global numpy stream. */
int caelo_ransac_draw_samples(caelo_ctx *ctx, const int64_t *seeds, int P, int n_points, int T, int rounds,
                              int rounds_done, int32_t *samples, void *stream);

/* f1 — ProjectPC2SphericalRing (SphericalRing.py:72-94) for F scans at once.  pts: dev float32 rows
 * (x,y,z,intensity), all scans concatenated; pts_offsets: HOST [F+1] int64 row offsets into pts.
 * Outputs, each dev or NULL (at least one): ring5 [F,69,1800,5] f32 and counter_i32 [F,69,1800] — the
 * function's two return values; ring3 [F,64,1792,3] f32 = ring5[:, 0:64, 0:1792, 0:3] and counter_i8
 * [F,69,1800] — what GetAllRespondImgs (BatchPreprocess.py:97-108) builds from them for the CNN (the
 * int8 copy wraps like numpy's assignment).  Every element of a requested output is written.
 * status: dev int32 [F] or NULL — number of points whose column index equals ImgW (a point exactly on
 * the -x axis with y = -0.0; the reference raises IndexError there, here the point is skipped). */
int caelo_project_ring(caelo_ctx *ctx, const float *pts, const int64_t *pts_offsets, int F, float *ring5,
                       int32_t *counter_i32, float *ring3, int8_t *counter_i8, int32_t *status,
                       void *stream);

/* f2 — Voxelization (Voxel.py:100-173) for F scans at once: the ordered occupied-voxel lists that
 * BatchVoxelization.py:61 stores.  pts / pts_offsets as above; cap = rows reserved per list (>= the
 * largest scan).  Outputs (dev): vox int16 [F,3,cap,3] — list s of frame f starts at row (f*3+s)*cap
 * (AllVoxels0 grouped by 1.28 m block in block-first-seen order, AllVoxels1/2 in first-seen order);
 * counts int32 [F,4] = len(AllVoxels0), len(AllVoxels1), len(AllVoxels2), len(avlBlocksList);
 * optional local0 int16 [F,cap,3] (AllVoxels), blocks int16 [F,cap,3] (avlBlocksList), cnt int32
 * [F,cap+1] (cntVoxelsLength); status int32 [F] or NULL — points that index outside the 156x156x23
 * block grid (the reference raises IndexError; here they are skipped). */
int caelo_voxelize(caelo_ctx *ctx, const float *pts, const int64_t *pts_offsets, int F, int cap,
                   int16_t *vox, int32_t *counts, int16_t *local0, int16_t *blocks, int32_t *cnt,
                   int32_t *status, void *stream);

/* f2+a6 fused — GetPatchesList(Pts, *Voxelization(scan)[6:9]) (Voxel.py:100-216) for F scans without
 * materialising the voxel lists: the occupancy bricks are built straight from the points with
 * Voxelization's arithmetic (the list ORDER never reaches GetPatchesList).  kpts / n_kpts / packed /
 * patches_f32 / trunc as in caelo_gather_patches; pts / pts_offsets as in caelo_voxelize.
 * nvox: dev int32 [F,3] — distinct voxels per scale (= len(AllVoxels_s)); status: dev int32 [F] — low 30
 * bits: points outside the block grid (reference: IndexError), bit 30: a list has < 496 voxels
 * (reference: sklearn ValueError); patches of such a frame are all-zero. */
int caelo_gather_patches_scans(caelo_ctx *ctx, const void *kpts, int kpts_f64, const int32_t *n_kpts,
                               int F, int K, const float *pts, const int64_t *pts_offsets,
                               uint32_t *packed, float *patches_f32, uint8_t *trunc, int32_t *nvox,
                               int32_t *status, void *stream);

/* a6 / f2+a6 in two steps.  The occupancy index (bricks) of a batch does not depend on its key points, so it can
 * be built on a second stream while the key points are still being selected: caelo_bricks_build[_scans] (arguments
 * as in caelo_gather_patches[_scans]) then caelo_bricks_gather for the same F.  The index lives in the ctx (one at a
 * time); nvox / status of the scans variant are passed on to the gather (NULL, NULL after caelo_bricks_build). */
int caelo_bricks_build(caelo_ctx *ctx, const int16_t *vox, const int64_t *vox_offsets, int F, void *stream);
int caelo_bricks_build_scans(caelo_ctx *ctx, const float *pts, const int64_t *pts_offsets, int F, int32_t *nvox,
                             int32_t *status, void *stream);
int caelo_bricks_gather(caelo_ctx *ctx, const void *kpts, int kpts_f64, const int32_t *n_kpts, int F, int K,
                        uint32_t *packed, float *patches_f32, uint8_t *trunc, const int32_t *nvox,
                        int32_t *status, void *stream);

/* f4 (front end) — ExtendKeyPtsInShpericalRing (SphericalRing.py:294-317) for B frames: all occupied pixels
 * of each key pixel's 13x13 window, first key pixel wins a pixel, output in key-pixel order then row-major
 * inside the window.  ring / counter as in caelo_select_keypoints (shared pixel origin); kpix dev
 * [B,max_kpts,2] int64 (row,col), n_kpts dev [B] or NULL.  ext: dev [B,ext_cap,3] f32 (ext_cap <=
 * max_kpts*169 always suffices), n_ext dev [B].  zero_counter != 0 reproduces the reference's in-place side
 * effect (the windows of `counter` are zeroed, :307). */
int caelo_extend_keypoints(caelo_ctx *ctx, const float *ring, int ring_C, int ring_H, int ring_W,
                           void *counter, int counter_dtype, int cnt_H, int cnt_W, const int64_t *kpix,
                           const int32_t *n_kpts, int B, int max_kpts, float *ext, int ext_cap,
                           int32_t *n_ext, int zero_counter, void *stream);

/* f4 — the device side of MyICP.py:28-73 `ICP` (one iteration = caelo_nn3, caelo_kabsch on the inliers,
 * caelo_transform_points; the loop control stays on the host as in the reference).
 * caelo_nn3: exact 1-nearest neighbour of every pc1 point (dev [M,3] f32) among pc0 (dev [N,3] f32) — what the
 * reference asks sklearn for (MyICP.py:33-34, :77-78): idx dev int64 [M], dist dev float64 [M] (float64 Euclidean
 * distance, ties -> lowest index); mask dev uint8 [M] or NULL = dist < thr; count dev int32 [1] or NULL = number of
 * inliers.  caelo_transform_points: pc <- R pc + T in place (MyICP.py:51), Rt dev [12] (R row-major, T). */
int caelo_nn3(caelo_ctx *ctx, const float *pc0, int N, const float *pc1, int M, int64_t *idx, double *dist,
              double thr, uint8_t *mask, int32_t *count, void *stream);
int caelo_transform_points(caelo_ctx *ctx, const float *Rt, float *pc, int M, void *stream);

/* f4, batched — WHOLE ICPs (MyICP.py:28-73) for B frame pairs in one call, no host round trip per iteration (what
 * RefinePoses.py:273-334 `RefinementCore` runs per key-frame pair; configs[4] shards the pairs of seq 00-10 over ranks).
 * pc0 dev [S0,3] f32 = the B target clouds concatenated, off0 HOST int64 [B+1] row offsets; pc1 dev [S1,3] f32 = the B
 * source clouds (already moved by the odometry pose), off1 likewise — pc1 is the work copy and is updated in place.
 * Per iteration and pair: exact nearest neighbour within the current inlier threshold (uniform grid over pc0, same
 * pairs as the brute-force search of caelo_nn3), SolveRT on them (caelo_kabsch's arithmetic), pc1 <- R pc1 + T, then
 * the reference's loop control on the device: fewer than min_inliers -> failure; Euler-angle / translation norms below
 * ep after min_iter iterations -> converged; both below small_shift -> thr *= decay.
 * Out: hist dev [B,max_iter,12] f32 = [R|T] of every iteration run (zeros beyond), hist_n dev int32 [B,max_iter] = its
 * inlier count, state dev float64 [B,4] = success flag, iterations run, last inlier count, final threshold.  The host
 * accumulates R*, T* from hist with the reference's own numpy expressions (api.icp_batch). */
int caelo_icp_batch(caelo_ctx *ctx, const float *pc0, const int64_t *off0, float *pc1, const int64_t *off1, int B,
                    double thr0, double decay, double small_shift, double ep, int max_iter, int min_iter,
                    int min_inliers, float *hist, int32_t *hist_n, double *state, void *stream);

/* Test / measurement hooks of the nn match (tools/nn_margin.py): the first pass's approximate per-column results of the LAST
 * caelo_nn_match call (dev float [P,M] best and second-best d^2, dev int32 [P,M] best row, dev int32 [1] number of columns
 * that went to the exact re-scan; any may be NULL), and the error margin E = margin * |a|max |b_j| assumed for that pass. */
int caelo_debug_nn_last(caelo_ctx *ctx, int P, int N, int M, float *best_d, float *second_d, int32_t *best_i,
                        int32_t *n_undecided, void *stream);
int caelo_debug_set_nn_margin(caelo_ctx *ctx, float margin);

/* Debug: device buffer [grid][64][16] int64 (+ [1024][8] for dense) receiving clock64 stamps of the encoder's per-patch
 * phases (NULL disables).  Used by tools/encoder_timeline.py. */
int caelo_debug_set_timeline(caelo_ctx *ctx, long long *buf);

/* Test hook (tests/test_gpu_umma.py): one tcgen05 GEMM D = A*B^T (f16 in, f32 accumulate in
 * TMEM) with operands staged under caller-chosen SBO/LBO byte strides; dumps the raw 128-lane x
 * 64-column accumulator.  Pins the descriptor and TMEM-layout facts the encoder relies on. */
int caelo_debug_umma(caelo_ctx *ctx, const void *A_f16, const void *B_f16, int M, int N, int K,
                     int sbo_a, int lbo_a, int sbo_b, int lbo_b, int d_lane_off, float *dump,
                     void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CAELO_H */

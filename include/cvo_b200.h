/*
 * include/cvo_b200.h -- C ABI of libcvo_b200.so: the B200-native (sm_100a) replacement for the
 * RKHS SE(3) registration inner loop of MaaniGhaffari/cvo-rgbd.
 *
 * The reference has no plugin / FFI layer: the hot path is four PRIVATE member functions
 *     transform_pcd -> se_kernel -> compute_flow -> compute_step_size
 * called only from align() of cvo::cvo / acvo::acvo
 * (cpp/rkhs_registration/src/cvo.cpp:361-420, src/adaptive_cvo.cpp:490-555).  This ABI sits directly
 * underneath align(): the C++ frontends in include/cvo_b200_frontend.hpp keep the reference's public
 * surface (init, iter, transform, prev_transform, accum_transform, set_pcd, align, run_cvo,
 * function_inner_product; inc/cvo.hpp:101-107,171-192, inc/adaptive_cvo.hpp:108-114,169-195) and
 * forward to these entry points.  Plain pointers and sizes only; no C++ / torch types; nothing throws.
 *
 * All pointers are HOST pointers unless a name ends in _dev.  Matrices are row-major.
 * One ctx = one GPU = one caller thread (like one cvo object, which is not thread-safe either).
 * There is no CPU fallback: every entry point fails with CVO_B200_ERR_CUDA if the device is unusable.
 */
#ifndef CVO_B200_H
#define CVO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cvo_b200_ctx cvo_b200_ctx;

enum { CVO_B200_OK = 0,
       CVO_B200_ERR_ARG = -1,     /* bad slot / size / NULL pointer                     */
       CVO_B200_ERR_CUDA = -2,    /* CUDA runtime error; see cvo_b200_last_error()      */
       CVO_B200_ERR_EMPTY = -3,   /* empty cloud (the reference asserts in nanoflann,
                                     thirdparty/KDTreeVectorOfVectorsAdaptor.h:61)      */
       CVO_B200_ERR_UNSUPPORTED = -4 }; /* a branch of the reference that is not built  */

enum { CVO_B200_MODE_CVO = 0,     /* cvo::compute_flow   (src/cvo.cpp:164-210)          */
       CVO_B200_MODE_ACVO = 1 };  /* acvo::compute_flow  (src/adaptive_cvo.cpp:154-272) */

enum { CVO_B200_ELL_SCHEDULE = 0, /* src/cvo.cpp:408-410                                */
       CVO_B200_ELL_ADAPTIVE = 1, /* src/adaptive_cvo.cpp:538-545                       */
       CVO_B200_ELL_FIXED = 2 };  /* benchmark configs 2 and 5                          */

/* per-pair exit status of align() */
enum { CVO_B200_STATUS_MAX_ITER = 0,        /* loop cap hit (src/cvo.cpp:366)           */
       CVO_B200_STATUS_CONVERGED_TWIST = 1, /* |omega|<eps && |v|<eps (src/cvo.cpp:380) */
       CVO_B200_STATUS_CONVERGED_UPDATE = 2,/* dist_se3(dR,dT)<eps_2 (src/cvo.cpp:402)  */
       CVO_B200_STATUS_NAN = 3 };           /* non-finite twist                         */

/* Every constructor initialiser of the two reference classes
 * (src/cvo.cpp:18-48, src/adaptive_cvo.cpp:18-50). */
typedef struct cvo_b200_params {
    int    mode;        /* CVO_B200_MODE_*                                       */
    int    ell_policy;  /* CVO_B200_ELL_*                                        */
    float  ell_init;    /* cvo 0.15 (src/cvo.cpp:25), acvo 0.1 (adaptive:25)     */
    float  ell_min;     /* acvo 0.0391                                           */
    float  ell_max;     /* acvo 0.15, re-armed per pair (adaptive:477)           */
    double dl_step;     /* acvo 0.3                                              */
    float  sigma;       /* 0.1                                                   */
    float  sp_thres;    /* cvo 8e-3, acvo 8.315e-3                               */
    float  c;           /* 7                                                     */
    float  d;           /* 7                                                     */
    float  c_ell;       /* cvo 200, acvo 0.5                                     */
    float  c_sigma;     /* 1                                                     */
    float  c_sp_thres;  /* acvo 8.315e-3 (cvo gates colour with sp_thres)        */
    int    max_iter;    /* 2000                                                  */
    float  min_step;    /* 0.2                                                   */
    float  max_step;    /* 0.8                                                   */
    float  eps;         /* 5e-5                                                  */
    float  eps_2;       /* 1e-5                                                  */
    int    fixed_iters; /* >0: run exactly this many iterations, stop tests off  */
} cvo_b200_params;

/* One outer iteration's observables (level-1 / level-2 parity hooks).  The reference keeps
 * these in private members: omega, v (inc/cvo.hpp:93-94), step (:88), B..E (src/cvo.cpp:242-245),
 * A.nonZeros(), dl (inc/adaptive_cvo.hpp:77). */
typedef struct cvo_b200_iter_rec {
    float     ell;
    float     step;
    float     omega[3];
    float     v[3];
    double    B, C, D, E;
    double    sum_a;
    double    dl;
    long long nnz, nnz_xx, nnz_yy;
    float     R[9];   /* state after this iteration's update */
    float     T[3];
} cvo_b200_iter_rec;

void cvo_b200_default_params_cvo(cvo_b200_params* p);   /* == cvo::cvo()   src/cvo.cpp:18-48          */
void cvo_b200_default_params_acvo(cvo_b200_params* p);  /* == acvo::acvo() src/adaptive_cvo.cpp:18-50 */

/* Creates a context on `device` with `max_slots` frame-pair slots of up to `max_points` points per
 * cloud.  All device buffers are owned by the ctx. */
int  cvo_b200_create(cvo_b200_ctx** out, int device, int max_points, int max_slots);
void cvo_b200_destroy(cvo_b200_ctx* ctx);
const char* cvo_b200_last_error(const cvo_b200_ctx* ctx);

/* Replaces the tail of set_pcd() (src/cvo.cpp:343-356): binds the two clouds of a frame pair to a slot.
 * xyz: n x 3 f32 (cloud_t, inc/data_type.h:30); feat: n x 5 f32 ROW-major (the reference's
 * point_cloud::features is column-major, inc/data_type.h:64 -- callers pass .transpose() or rows).
 * The data is copied (caller keeps ownership), spatially sorted and packed on the device.
 * Enqueues on the ctx stream and returns; pageable host buffers may be reused on return, pinned ones
 * after cvo_b200_sync(). */
int cvo_b200_set_pair(cvo_b200_ctx* ctx, int slot,
                      const float* fixed_xyz, const float* fixed_feat, int n_fixed,
                      const float* moving_xyz, const float* moving_feat, int n_moving);

/* cvo_b200_set_pair for `n_pairs` independent frame pairs at once (the batch / multi-GPU driver of
 * BASELINE config 4; the reference binds one pair per object, src/cvo.cpp:343-356): four host->device
 * copies and ONE pack launch for the whole batch.  Cloud i of every array starts at element
 * i * stride_points * 3 (xyz) or i * stride_points * 5 (feat) and holds n_fixed[i] / n_moving[i]
 * (<= stride_points) points.  Same copy / ownership rules as cvo_b200_set_pair.
 * Pipelining: the copies are enqueued on a copy stream of their own and the call returns at once; the pack
 * launch follows them on that stream (or, while a cvo_b200_align_begin is in flight, is enqueued when something
 * first needs these slots).  A driver that binds batch k+1 to a second set of slots BEFORE calling
 * cvo_b200_align on batch k therefore overlaps the upload of k+1 with the align kernel of k, and the pack of
 * k+1 fills the SMs that kernel's last wave leaves idle (two batches can be in flight; host buffers must stay
 * untouched until the batch has been aligned or cvo_b200_sync() returned). */
int cvo_b200_set_pairs(cvo_b200_ctx* ctx, const int* slots, int n_pairs,
                       const float* fixed_xyz, const float* fixed_feat, const int* n_fixed,
                       const float* moving_xyz, const float* moving_feat, const int* n_moving,
                       int stride_points);

/* The multi-GPU batch driver (BASELINE config 4; the reference's driver loop src/cvo_main.cpp:36-66 over INDEPENDENT
 * pairs, each started from the identity): `n_pairs` frame pairs laid out as for cvo_b200_set_pairs are dealt round-robin
 * to `n_ctx` contexts -- one per GPU, pair q -> ctxs[q mod n_ctx] -- and every context uploads and aligns its share on a
 * host thread of its own, in chunks of its slot count.  No data-path collective: the pairs are independent; the 4x4
 * `transform`s (n_pairs x 16), `iters` and `status` (n_pairs, may be NULL) are written at index q, which is the gather.
 * kernel_ms (n_ctx, may be NULL) receives each context's summed align-kernel time.  One process, all devices; a
 * one-process-per-GPU job shards with the same rule and gathers with one NCCL all-gather (cvo_rgbd_b200/sharding.py). */
int cvo_b200_align_multi(cvo_b200_ctx* const* ctxs, int n_ctx, int n_pairs,
                         const float* fixed_xyz, const float* fixed_feat, const int* n_fixed,
                         const float* moving_xyz, const float* moving_feat, const int* n_moving,
                         int stride_points, const cvo_b200_params* p,
                         float* transform, int* iters, int* status, float* kernel_ms);

/* Replaces `ptr_fixed_pcd = std::move(ptr_moving_pcd)` (src/cvo.cpp:417) + the next set_pcd():
 * the slot's moving cloud becomes its fixed cloud (pointer swap on the device) and a new moving cloud
 * is uploaded, so a sequence uploads each frame once. */
int cvo_b200_push_frame(cvo_b200_ctx* ctx, int slot, const float* xyz, const float* feat, int n);
/* Replaces the slot's MOVING cloud only (no promotion): what a second set_pcd() without an align() in between does in
 * the reference, where the promotion sits at the end of align() (src/cvo.cpp:336-351 vs :417).  The frontends call
 * this instead of cvo_b200_push_frame when no align() has run since the last set_pcd(). */
int cvo_b200_replace_moving(cvo_b200_ctx* ctx, int slot, const float* xyz, const float* feat, int n);

/* Image front door.  Replaces pcd_generator::load_image + create_pointcloud
 * (src/pcd_generator.cpp:384-420: cv::cvtColor to gray / HSV, the 3-level gradient pyramid, DSO's PixelSelector2
 * with num_want = 3000, pinhole back-projection with the intrinsics of `dataset_seq` (:241-302), the 5-D features
 * of `feature_type` 0 = HSV/[180,255,255] + gradient*2/255 (acvo) or 1 = raw channels + raw gradient (cvo)) and the
 * tail of set_pcd() (src/cvo.cpp:319-357), entirely on the device: the image and depth are copied once and the
 * cloud is written straight into the slot's packed planes.
 *  img3   height x width x 3, 8-bit, as cv::imread returns it;  depth  height x width, 16-bit.
 *  width and height must be multiples of 32 (TUM: 640 x 480).
 * The first call on a fresh slot creates the FIXED cloud (src/cvo.cpp:326-334); every later call makes the slot's
 * moving cloud its fixed cloud (:417) and creates the new moving cloud.  *num_points receives the cloud size.
 * A low-texture frame (fewer than num_want/3 pixels selected) gets the reference's Canny top-up
 * (src/pcd_generator.cpp:135-163: cv::blur 3x3, cv::Canny(0, 25, 3), one extra pixel per 8 x 8 block), also on the
 * device.  Synchronous.  On an error the slot is left unchanged. */
int cvo_b200_push_frame_images(cvo_b200_ctx* ctx, int slot, const unsigned char* img3, const unsigned short* depth,
                               int width, int height, int dataset_seq, int feature_type, int* num_points);
/* cvo_b200_push_frame_images without the promotion: the new cloud replaces the slot's moving cloud (see
 * cvo_b200_replace_moving).  On a slot without a bound pair it behaves like cvo_b200_push_frame_images. */
int cvo_b200_replace_moving_images(cvo_b200_ctx* ctx, int slot, const unsigned char* img3, const unsigned short* depth,
                                   int width, int height, int dataset_seq, int feature_type, int* num_points);
/* cvo_b200_align in two halves, for a driver that has something else to enqueue while the kernel runs (the front end of
 * the next frame: cvo_b200_prefetch_frame_images).  _begin validates, writes the pair records and launches; _finish waits
 * and hands the results out exactly as cvo_b200_align does.  One align in flight per context: a second _begin, or any
 * other align / eval call, before _finish is an error.  Uploads enqueued in between are ordered behind the kernel. */
int cvo_b200_align_begin(cvo_b200_ctx* ctx, const int* slots, int n_pairs, const cvo_b200_params* params,
                         const float* RT_in, const float* ell_in);
int cvo_b200_align_finish(cvo_b200_ctx* ctx, float* RT_out, float* ell_out, float* transform, float* prev_transform,
                          int* iters, int* status);
/* Look-ahead for a sequence driver (src/cvo_main.cpp:36-66 reads frame k + 1 from disk before it needs it): starts the
 * image front end for the NEXT frame on the context's copy stream and returns at once -- image upload, the 23 launches
 * and the packed cloud (into a context-level buffer) overlap the cvo_b200_align() of the current pair, which occupies a
 * fraction of the SMs.  cvo_b200_push_prefetched_frame then gives the cloud to a slot exactly as
 * cvo_b200_push_frame_images (promote != 0) / cvo_b200_replace_moving_images (promote == 0) would have: it waits for the
 * prefetch (normally long finished), checks the point count and copies the cloud device-to-device.  The image buffers
 * must stay valid until cvo_b200_push_prefetched_frame returns.  One prefetched frame per context; a
 * cvo_b200_push_frame_images in between discards it.  The resulting cloud is bit-identical to the synchronous path's. */
int cvo_b200_prefetch_frame_images(cvo_b200_ctx* ctx, const unsigned char* img3, const unsigned short* depth, int width,
                                   int height, int dataset_seq, int feature_type);
int cvo_b200_push_prefetched_frame(cvo_b200_ctx* ctx, int slot, int promote, int* num_points);
/* The cloud the last cvo_b200_push_frame_images generated, in the reference's (raster) order: xyz n x 3,
 * feat n x 5 row-major (parity hook; valid until the next upload of any kind). */
int cvo_b200_last_generated_cloud(cvo_b200_ctx* ctx, float* xyz, float* feat, int capacity, int* n);
/* Whether the last cvo_b200_push_frame_images took the Canny top-up branch. */
int cvo_b200_last_frame_used_canny(const cvo_b200_ctx* ctx);
/* Forgets the clouds of a slot (start of a new sequence). */
int cvo_b200_reset_slot(cvo_b200_ctx* ctx, int slot);
/* Test hook: the byte sequence `rand() & 0xFF` of glibc after srand(seed), as used for the selector's
 * randomPattern (thirdparty/PixelSelector2.cpp:36-38), produced without touching the C library's state. */
int cvo_b200_selftest_rand_bytes(unsigned seed, int n, unsigned char* out);
/* Test hook: the device's line search on n coefficient sets {B, C, D, E} (poly_solver + root selection,
 * src/cvo.cpp:53-69,291-307), one result per set. */
int cvo_b200_selftest_step_size(cvo_b200_ctx* ctx, const double* bcde, int n, float min_step, float max_step, float* out);
/* Test hook: the device's Exp_SEK3 (src/LieGroup.cpp:159-186, incl. the small-angle quirk) on n rows
 * {omega[3], v[3], dt}; out: n rows {dR row-major [9], dT[3]}. */
int cvo_b200_selftest_exp_sek3(cvo_b200_ctx* ctx, const float* omega_v_dt, int n, float* dR_dT);

/* One pass of transform_pcd + se_kernel + compute_flow + compute_step_size at a given state
 * (src/cvo.cpp:368-377) without updating anything: fills one record (R,T echo the input). */
int cvo_b200_eval(cvo_b200_ctx* ctx, int slot, const float* R, const float* T, float ell,
                  const cvo_b200_params* p, cvo_b200_iter_rec* out);

/* Replaces align() (src/cvo.cpp:361-420, src/adaptive_cvo.cpp:490-555) for `n_pairs` slots at once; the
 * whole iteration (flow, step size, cubic, Exp_SEK3, stop tests, ell policy) runs on the device.
 *  RT_io        n_pairs x 12 (R row-major, then T): in = carried-in state (quirk Q4 warm start),
 *               out = state at loop exit.  NULL: start from identity, result not returned.
 *  ell_io       n_pairs: in/out length-scale (cvo never re-arms ell between pairs); NULL: params->ell_init.
 *  transform    n_pairs x 16: `transform` after align(), i.e. [R^T, -R^T T] of the final state (src/cvo.cpp:415)
 *  prev_transform  n_pairs x 16 or NULL: the one-update-stale transform that the reference multiplies
 *               into accum_transform (quirk Q3, src/cvo.cpp:413-414)
 *  iters        n_pairs: k at exit (max_iter when the cap was hit);  status: CVO_B200_STATUS_*
 * Every array is indexed like `slots`.  (Inside, the free clusters pull the pairs from a queue ordered largest
 * N x M first, so that a ragged batch ends on its cheap pairs; a pair's result does not depend on its place.)
 * Synchronous: returns when results are in the host arrays. */
int cvo_b200_align(cvo_b200_ctx* ctx, const int* slots, int n_pairs, const cvo_b200_params* p,
                   float* RT_io, float* ell_io, float* transform, float* prev_transform,
                   int* iters, int* status);

/* Same as cvo_b200_align for ONE slot, additionally returning the first `trace_cap` iteration records
 * (level-2 parity).  *trace_len receives the number of iterations executed. */
int cvo_b200_align_trace(cvo_b200_ctx* ctx, int slot, const cvo_b200_params* p,
                         float* RT_io, float* ell_io, float* transform, float* prev_transform,
                         int* iters, int* status,
                         cvo_b200_iter_rec* trace, int trace_cap, int* trace_len);

/* Replaces acvo::function_inner_product(cloud_a, cloud_b) (src/adaptive_cvo.cpp:385-439) with
 * cloud_a = the slot's fixed cloud and cloud_b = its moving cloud, both UNtransformed, at length-scale
 * `ell`.  *value = float(sum_a/nnz) as the reference returns; the parts are returned too. */
int cvo_b200_inner_product(cvo_b200_ctx* ctx, int slot, float ell, const cvo_b200_params* p,
                           float* value, double* sum_a, long long* nnz);

/* Blocks until everything enqueued on the ctx stream has finished. */
int cvo_b200_sync(cvo_b200_ctx* ctx);

/* Measurement hooks (bench.py): device time of the kernels of the last align / eval call measured
 * with CUDA events on the ctx stream, kernels launched so far, and the launch geometry last used. */
float     cvo_b200_last_kernel_ms(const cvo_b200_ctx* ctx);
long long cvo_b200_kernel_launches(const cvo_b200_ctx* ctx);
int       cvo_b200_last_cluster_size(const cvo_b200_ctx* ctx);
int       cvo_b200_last_num_clusters(const cvo_b200_ctx* ctx);
/* Overrides the CTAs-per-pair choice (1..16); 0 = automatic: the size that minimises waves x per-pair time for
 * the batch at hand, a single wave counted as its slowest pair (a single pair: 16 CTAs; 20 pairs: 6; 63 .. 160
 * pairs: 2; 2 x #SMs pairs: 1 -- measured table: profiles/r02_cfg4_cluster_sweep.txt). */
int       cvo_b200_set_cluster_size(cvo_b200_ctx* ctx, int ctas_per_pair);
/* Whole-GPU mode for one large pair (the reference's own use: one warm-started pair at a time, src/cvo_main.cpp:36-52;
 * BASELINE config 5): the pairs of an align call are taken one after the other and EVERY cluster of the launch works
 * on the current one; the clusters' partial sums meet in global memory (two grid-wide barriers per iteration).
 * clusters_per_pair: 0 = automatic (a single-pair call with at least eight row tiles per CTA of one cluster, i.e.
 * more than 4096 points at 16 CTAs, takes every cluster the device holds; smaller pairs stay on one cluster: their
 * iteration is dominated by the part that does not split), 1 = off (one cluster per pair), n > 1 = n clusters
 * (clamped to what the device holds).  Results are bit-deterministic for a given (ctas_per_pair, clusters_per_pair). */
int       cvo_b200_set_group_clusters(cvo_b200_ctx* ctx, int clusters_per_pair);
int       cvo_b200_last_group_clusters(const cvo_b200_ctx* ctx);
/* Sum over the pairs of the last align call of iterations executed (work accounting for the roofline). */
long long cvo_b200_last_total_iterations(const cvo_b200_ctx* ctx);
/* Neighbour candidate lists: the device-side replacement of the kd-tree the reference rebuilds in every
 * se_kernel call (src/cvo.cpp:106-113, thirdparty/nanoflann.hpp).  Per frame pair the (row, col) index pairs
 * inside a ball of radius (1 + skin) * r are kept in a per-CTA scratch area and re-used by the passes of the
 * following iterations until the pose has moved the moving cloud by more than skin * r (or ell changed r);
 * the strict ell-ball test and the kernel values are still evaluated on the fly, A is never stored.
 * enable = 0: every pass tests all tile pairs on the fly (also the automatic fallback if a list overflows
 * its scratch).  Default: enabled, skin = 0.10.  Results agree between the two modes up to f32 summation order. */
int       cvo_b200_set_neighbor_lists(cvo_b200_ctx* ctx, int enable, float skin);
/* Sum over the pairs of the last align call of (x, y) list builds. */
long long cvo_b200_last_list_builds(const cvo_b200_ctx* ctx);
/* ... and of the times one of the pair's lists ((x, y), or acvo's (x, x) / (y, y)) was narrowed in place after ell
 * shrank: a filter of the old list instead of an all-pairs sweep. */
long long cvo_b200_last_list_refines(const cvo_b200_ctx* ctx);
/* Work accounting of the (x, y) lists of the last align call, summed over its pairs and their list builds (rank-0 CTA
 * of every cluster): candidates kept, and the quad slots that hold them (four per quad, padding included). */
int       cvo_b200_last_list_fill(const cvo_b200_ctx* ctx, long long* entries, long long* slots);
int       cvo_b200_num_sms(const cvo_b200_ctx* ctx);
/* 1 if the last align used (or the next will use) neighbour lists; 0 if they are disabled or their scratch could not
 * be allocated (every pass then runs on the fly: same results, several times slower). */
int       cvo_b200_neighbor_lists_active(const cvo_b200_ctx* ctx);
/* Bytes of HBM scratch the neighbour lists currently occupy: CTAs of the largest launch so far x 5 areas x capacity,
 * the capacity following the largest cloud aligned so far (not max_points). */
long long cvo_b200_list_scratch_bytes(const cvo_b200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif

// cvo_api.cu -- host side of libcvo_b200.so: context, buffers, launches, the C ABI of include/cvo_b200.h.
// No torch, no CPU fallback: every entry point needs a CUDA device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <algorithm>
#include <vector>

#include "cvo_kernels.cuh"
#include "pcd_kernels.cuh"

using namespace cvo_b200;

struct cvo_b200_ctx {
    int device = 0;
    int max_points = 0;
    int max_slots = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // raw staging (stream-ordered reuse): two clouds of xyz (n x 3) and feat (n x 5)
    float* d_raw_xyz = nullptr;
    float* d_raw_feat = nullptr;
    // packed clouds: [slot][2][max_points]
    float4* d_pk_g = nullptr;
    float4* d_pk_f = nullptr;
    float* d_pk_f4 = nullptr;

    struct Slot {
        int n[2] = {0, 0};
        int fixed_buf = 0;  // which of the two buffers currently holds the fixed cloud
        bool bound = false;
        bool have_fixed = false;  // a fixed cloud is in place (image front door: the first frame of a sequence)
        int pending_batch = -1;  // batched upload whose pack launch has not been enqueued yet
    };
    std::vector<Slot> slots;

    PairDev* h_pairs = nullptr;     // pinned, mapped: the kernels read the descriptors from here
    PairState* h_states = nullptr;  // pinned, mapped: carried-in state in, result out
    PairDev* h_pairs_dev = nullptr;     // the device's view of the two
    PairState* h_states_dev = nullptr;
    int* d_counter = nullptr;
    double* d_group_xchg = nullptr;  // whole-GPU mode (group_allreduce): cluster records, and d_counter[1] counts the arrivals
    cvo_b200_iter_rec* d_trace = nullptr;
    cvo_b200_iter_rec* h_trace = nullptr;  // pinned
    int trace_cap = 0;
    double* d_inner = nullptr;
    PackJob* d_jobs = nullptr;  // descriptors of the pack launch in flight (stream-ordered reuse)

    // Batched upload (cvo_b200_set_pairs): two staging areas so that the host->device copies of one batch run on
    // `copy_stream` while the align kernel of the previous batch runs on `stream`.  The pack launch of a batch is
    // deferred until something needs its slots (it could not overlap the persistent align kernel anyway and must
    // not be queued in front of it) -- unless no align is in flight when the batch is uploaded: then the pack goes
    // onto `copy_stream` right behind the copies (packed_eager).  In a pipelined driver (upload of batch k + 1, then
    // align of batch k) its CTAs fill the SMs the align kernel's last wave leaves idle instead of standing in front of the
    // next align.
    struct Batch {
        float* d_raw = nullptr;       // [fixed xyz | fixed feat | moving xyz | moving feat]
        size_t raw_floats = 0;
        PackJob* d_jobs = nullptr;    // 2 * max_slots
        PackJob* h_jobs = nullptr;    // pinned
        cudaEvent_t copied = nullptr; // recorded on copy_stream after the batch's copies
        cudaEvent_t packed = nullptr; // recorded on stream after the pack that read d_raw
        bool pending = false;         // copies enqueued, pack not yet (or enqueued on copy_stream: packed_eager)
        bool packed_eager = false;
        bool used = false;
        int njobs = 0;
        std::vector<int> slots;
    } batch[2];
    int next_batch = 0;
    cudaStream_t copy_stream = nullptr;

    // image front end (cvo_b200_push_frame_images), allocated on first use for one image size
    struct ImagePipe {
        int w = 0, h = 0;
        uint8_t* d_img3 = nullptr;
        uint16_t* d_depth = nullptr;
        float* d_pyr = nullptr;      // I, dx, dy, g2 of the three levels, one allocation
        float* d_ths = nullptr;      // ths | thsSmoothed
        uint8_t* d_map = nullptr;
        uint8_t* d_rnd = nullptr;    // the selector's randomPattern
        int* d_blockcnt = nullptr;   // flagged pixels per 1024-pixel block (raster-order ranks)
        SelCtl* d_ctl = nullptr;
        SelCtl* h_ctl = nullptr;     // pinned
        bool tables = false;
        // cvo_b200_prefetch_frame_images: the NEXT frame's cloud, generated on copy_stream while an align() runs
        float* d_next_raw_xyz = nullptr;
        float* d_next_raw_feat = nullptr;
        float4* d_next_g = nullptr;
        float4* d_next_f = nullptr;
        float* d_next_f4 = nullptr;
        PackJob* d_next_job = nullptr;
        SelCtl* h_ctl_next = nullptr;    // pinned
        cudaEvent_t ev_next = nullptr;   // recorded on copy_stream behind the prefetched frame
        cudaEvent_t ev_taken = nullptr;  // recorded on stream behind the copy of a prefetched cloud into its slot
        bool next_pending = false;
    } pipe;
    int last_gen_n = 0;
    int last_gen_canny = 0;
    double* h_inner = nullptr;  // pinned

    // neighbour-list scratch (allocated on the first align): [num_sms][LIST_KINDS] areas
    uint2* d_list_entries = nullptr;
    unsigned list_cap = 0;
    int list_ctas = 0;  // CTAs the scratch currently has areas for
    bool lists_enabled = true;
    bool lists_alloc_failed = false;
    float list_skin = 0.10f;  // measured optimum (cfg2 and the stock schedules, profiles/r02_skin_sweep.txt)
    // Extra slack W / r of the WIDE (x, y) list (WideState in cvo_common.cuh); 0 = off, < 0 = automatic: 0.5 for acvo, whose
    // length-scale moves every iteration (12.8 rebuilds per pair: stock acvo +5 %), off for cvo (a filter of the wide list costs
    // 0.6 of a sweep and the wide sweep 1.6: cfg 2 -2 %, stock cvo -8 %; profiles/r02_wide_list.txt)
    float list_wide = -1.0f;
    float list_ahead = 0.5f;  // the (x, y) list is built this many skins ahead of the motion (list_policy; measured optimum)
    float list_skin_min = 0.003f;  // absolute floor of the skin [m]: stock cvo +4 % (short lists at small ell, fewer rebuilds)
    float list_shrink = 0.7f;
    float list_refine_min = 1.0f;
    long long last_list_builds = 0, last_list_refines = 0, last_xy_entries = 0, last_xy_slots = 0;

    float last_ms = 0.f;
    long long launches = 0;
    int last_G = 0, last_nclusters = 0, force_G = 0;
    int force_group = 0, last_group = 1;
    // an align launched by cvo_b200_align_begin and not yet collected by cvo_b200_align_finish
    std::vector<int> order;  // queue position -> index into the caller's arrays (largest pairs first, see run_align_begin)
    struct PendingAlign { bool active = false; int n_pairs = 0, G = 0, ncl = 0, group = 1, trace_cap = 0; bool trace = false; } pending;  // clusters per pair: 0 = automatic, 1 = one cluster per pair, n = whole-GPU mode
    int max_clusters[17] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};  // per cluster size, -1 = not asked yet
    long long last_total_iters = 0;
    int sort_points = 1;
    size_t pack_smem_max = 0;
    std::string err;
};

namespace {

constexpr int kTraceCap = 4096;

bool cuda_ok(cvo_b200_ctx* ctx, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    if (ctx) ctx->err = buf;
    return false;
}
#define CK(call)                                                        \
    do {                                                                \
        if (!cuda_ok(ctx, (call), #call)) return CVO_B200_ERR_CUDA;     \
    } while (0)

int fail_arg(cvo_b200_ctx* ctx, const char* msg) {
    if (ctx) ctx->err = msg;
    return CVO_B200_ERR_ARG;
}

float4* slot_g(cvo_b200_ctx* ctx, int slot, int buf) {
    return ctx->d_pk_g + ((size_t)slot * 2 + buf) * ctx->max_points;
}
float4* slot_f(cvo_b200_ctx* ctx, int slot, int buf) {
    return ctx->d_pk_f + ((size_t)slot * 2 + buf) * ctx->max_points;
}
float* slot_f4(cvo_b200_ctx* ctx, int slot, int buf) {
    return ctx->d_pk_f4 + ((size_t)slot * 2 + buf) * ctx->max_points;
}

// Host-side constants in the reference's own arithmetic (src/cvo.cpp:102-103, src/adaptive_cvo.cpp:100-101).
KParams make_kparams(const cvo_b200_params* p, bool for_inner_product) {
    KParams k;
    memset(&k, 0, sizeof(k));
    k.mode = p->mode;
    k.ell_policy = p->ell_policy;
    k.max_iter = p->max_iter;
    k.fixed_iters = p->fixed_iters;
    k.ell_min = p->ell_min;
    k.s2 = p->sigma * p->sigma;
    k.cs2 = p->c_sigma * p->c_sigma;
    k.sp_thres = p->sp_thres;
    float c_gate;
    if (for_inner_product) {  // src/adaptive_cvo.cpp:391-392
        k.log_ratio = logf(p->sp_thres / p->sigma / p->sigma);
        c_gate = p->sp_thres;
    } else {
        k.log_ratio = logf(p->sp_thres / k.s2);
        c_gate = (p->mode == CVO_B200_MODE_ACVO) ? p->c_sp_thres : p->sp_thres;
    }
    k.d2c_thres = (float)(-2.0 * p->c_ell * p->c_ell * logf(c_gate / p->c_sigma / p->c_sigma));
    k.inv2cl2 = (float)(1.0 / (2.0 * (double)p->c_ell * (double)p->c_ell));
    k.c2 = (float)(1.4426950408889634 / (2.0 * (double)p->c_ell * (double)p->c_ell));
    k.s2cs2 = (float)((double)k.s2 * (double)k.cs2);
    k.c_ell = p->c_ell;
    k.sp_band = 2.0e-6f * fabsf(p->sp_thres);
    k.t_lim = (float)(log2((double)k.s2cs2 / (double)p->sp_thres) + 1.0e-5);
    k.inv_c = 1 / p->c;
    k.inv_d = 1 / p->d;
    k.min_step = p->min_step;
    k.max_step = p->max_step;
    k.eps = p->eps;
    k.eps_2 = p->eps_2;
    k.dl_step = p->dl_step;
    return k;
}

int upload_cloud(cvo_b200_ctx* ctx, int which, const float* xyz, const float* feat, int n) {
    float* dx = ctx->d_raw_xyz + (size_t)which * ctx->max_points * 3;
    float* df = ctx->d_raw_feat + (size_t)which * ctx->max_points * 5;
    CK(cudaMemcpyAsync(dx, xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(df, feat, sizeof(float) * 5 * n, cudaMemcpyHostToDevice, ctx->stream));
    return CVO_B200_OK;
}

int launch_pack(cvo_b200_ctx* ctx, const PackJob* jobs_host, int njobs, PackJob* d_jobs_scratch) {
    int nmax = 0;
    for (int i = 0; i < njobs; ++i) nmax = jobs_host[i].n > nmax ? jobs_host[i].n : nmax;
    int npad = 1;
    while (npad < nmax) npad <<= 1;
    const size_t smem = (size_t)npad * sizeof(unsigned long long);
    if (smem > ctx->pack_smem_max) return fail_arg(ctx, "cloud too large for the single-CTA sort");
    CK(cudaMemcpyAsync(d_jobs_scratch, jobs_host, sizeof(PackJob) * njobs, cudaMemcpyHostToDevice, ctx->stream));
    pack_sort_kernel<<<njobs, kPackThreads, smem, ctx->stream>>>(d_jobs_scratch, ctx->sort_points);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return CVO_B200_OK;
}

// Enqueues the deferred pack launch of a batched upload on the main stream (after its copies).
int launch_pack(cvo_b200_ctx* ctx, cvo_b200_ctx::Batch& B, cudaStream_t st) {
    int nmax = 0;
    for (int i = 0; i < B.njobs; ++i) nmax = B.h_jobs[i].n > nmax ? B.h_jobs[i].n : nmax;
    int npad = 1;
    while (npad < nmax) npad <<= 1;
    const size_t smem = (size_t)npad * sizeof(unsigned long long);
    pack_sort_kernel<<<B.njobs, kPackThreads, smem, st>>>(B.d_jobs, ctx->sort_points);
    CK(cudaGetLastError());
    CK(cudaEventRecord(B.packed, st));
    ctx->launches += 1;
    return CVO_B200_OK;
}

int flush_batch(cvo_b200_ctx* ctx, int b) {
    cvo_b200_ctx::Batch& B = ctx->batch[b];
    if (!B.pending) return CVO_B200_OK;
    B.pending = false;
    for (int s : B.slots)
        if (ctx->slots[s].pending_batch == b) ctx->slots[s].pending_batch = -1;
    if (B.packed_eager) {  // already packed on copy_stream: whatever uses the slots next waits for that
        B.packed_eager = false;
        CK(cudaStreamWaitEvent(ctx->stream, B.packed, 0));
        return CVO_B200_OK;
    }
    CK(cudaStreamWaitEvent(ctx->stream, B.copied, 0));
    return launch_pack(ctx, B, ctx->stream);
}
int flush_all_batches(cvo_b200_ctx* ctx) {
    for (int b = 0; b < 2; ++b) {
        const int rc = flush_batch(ctx, b);
        if (rc) return rc;
    }
    return CVO_B200_OK;
}

// How many clusters of G CTAs of align_kernel the device can hold at once (one CTA per SM; the GPCs' SM counts decide
// how many clusters of a given size fit).  Cached per context.
int max_resident_clusters(cvo_b200_ctx* ctx, int G) {
    if (ctx->max_clusters[G] >= 0) return ctx->max_clusters[G];
    int n = 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = sizeof(Smem);
    cfg.gridDim = dim3(G, 1, 1);
    bool ok = cudaFuncSetAttribute(align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) == cudaSuccess;
    if (ok && G > 8) ok = cudaFuncSetAttribute(align_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (!ok || cudaOccupancyMaxActiveClusters(&n, align_kernel, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    if (n * G > ctx->num_sms) n = ctx->num_sms / G;
    ctx->max_clusters[G] = n;
    return n;
}

// CTAs per pair.  Measured on BASELINE config 4 (profiles/r02_cfg4_cluster_sweep.txt): a pair on a cluster of G CTAs takes
// kPairTime[G] of its one-CTA time -- the serial section on one warp, the cluster barriers and the column staging every
// CTA repeats do not split.  With `fit` clusters resident, a batch of P <= fit pairs lasts as long as its SLOWEST pair:
// with the stop tests on the pairs' iteration counts differ by a factor of three (59 in the mean, 100 at the end of the
// tail: 1.5 mean pair times for a wave of 60 - 150 pairs), with a fixed iteration count only the list rebuilds differ
// (1.2).  A larger batch is pulled from the shared counter by whichever cluster is free: P / fit pair times plus about
// half a pair time of tail, and never less than the single wave.  The G with the smallest estimate wins, a larger G
// only if it is at least 3 % better; any size 1..16 is allowed (the row tiles are dealt rank * tiles / G).  E.g. the
// per-GPU shares of config 4: 63 pairs (8 GPUs) run as 63 clusters of 2 (3.7 ms; one CTA each: 6.8), 125 pairs (4 GPUs)
// as 74 clusters of 2 pulling from the queue (5.5 ms; one wave of 125 CTAs: 7.7), 250 and 500 pairs on one CTA each.
int choose_cluster(cvo_b200_ctx* ctx, int n_pairs, bool stop_tests) {
    if (ctx->force_G > 0) return ctx->force_G;
    static const double kPairTime[kMaxCluster + 1] = {1.0,  1.0,  0.57, 0.50, 0.45, 0.41, 0.38, 0.37, 0.36,
                                                      0.34, 0.32, 0.31, 0.29, 0.28, 0.27, 0.26, 0.25};
    const double slowest = n_pairs < 4 ? 1.0 : (stop_tests ? 1.5 : 1.2);
    int best_G = 1;
    double best = 1.0e30;
    for (int G = 1; G <= kMaxCluster; ++G) {
        const int fit = max_resident_clusters(ctx, G);
        if (fit < 1) continue;
        double waves = slowest;
        if (n_pairs > fit && (double)n_pairs / fit + 0.5 > waves) waves = (double)n_pairs / fit + 0.5;
        const double cost = waves * kPairTime[G];
        if (cost < best * 0.97) {
            best = cost;
            best_G = G;
        }
    }
    return best_G;
}

template <typename KernelT, typename ArgT>
int launch_cluster_kernel(cvo_b200_ctx* ctx, KernelT kernel, const ArgT& args, int G, int want_clusters,
                          int* nclusters_out) {
    const size_t smem = sizeof(Smem);
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (G > 8) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cfg.gridDim = dim3(G, 1, 1);
    int max_clusters = 0;
    CK(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
    if (max_clusters < 1) {
        ctx->err = "cluster size not schedulable on this device";
        return CVO_B200_ERR_CUDA;
    }
    int ncl = want_clusters < max_clusters ? want_clusters : max_clusters;
    if (ncl * G > ctx->num_sms) ncl = ctx->num_sms / G;  // one CTA per SM: per-CTA scratch is sized by it
    if (ncl < 1) ncl = 1;
    cfg.gridDim = dim3(ncl * G, 1, 1);
    CK(cudaLaunchKernelEx(&cfg, kernel, args));
    ctx->launches += 1;
    if (nclusters_out) *nclusters_out = ncl;
    return CVO_B200_OK;
}

// The neighbour-list scratch is only ever touched by the CTA it belongs to: [CTAs of the launch][3 lists + build
// staging] areas of `list_cap` entries.  It is sized by the launch at hand -- the CTAs actually launched and the
// largest cloud among the pairs being aligned, not max_points -- and only ever grows (a frontend object with
// max_points = 16384 aligning 3000-point pairs on one 16-CTA cluster holds 0.6 GB, not the 9.9 GB of a full-machine
// launch of maximum-size clouds).  Failing to allocate it is not an error: the passes then run on the fly
// (cvo_b200_neighbor_lists_active reports it).
void ensure_list_scratch(cvo_b200_ctx* ctx, int n_ctas, int max_n) {
    if (ctx->lists_alloc_failed || !ctx->lists_enabled) return;
    // (the cloud size in steps of 512 points: a sequence whose frames differ by a few points must not re-allocate -- a
    // re-allocation of the scratch costs about 17 ms, twenty alignments)
    auto cap_for = [](int n) {
        unsigned long long c = (unsigned long long)n * n / 8;
        if (c < (1ull << 18)) c = 1ull << 18;
        if (c > (1ull << 21)) c = 1ull << 21;
        return (c + 1023ull) & ~1023ull;
    };
    max_n = (max_n + 511) & ~511;
    unsigned long long cap = cap_for(max_n);
    if (!(ctx->d_list_entries && ctx->list_cap >= cap)) {  // (re)allocating: with a quarter of headroom, within max_points
        int roomy = (max_n + max_n / 4 + 511) & ~511;
        const int limit = (ctx->max_points + 511) & ~511;
        if (roomy > limit) roomy = limit > max_n ? limit : max_n;
        cap = cap_for(roomy);
    }
    if (const char* env = getenv("CVO_B200_LIST_CAP")) {  // test hook: a small area forces the overflow fallback
        const long long v = atoll(env);
        if (v >= 1024) cap = (unsigned long long)v & ~1023ull;
    }
    if (ctx->d_list_entries && ctx->list_ctas >= n_ctas && ctx->list_cap >= cap) return;
    if (ctx->d_list_entries) {  // grow: nothing is in flight on this stream between align calls, but be explicit
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->d_list_entries);
        ctx->d_list_entries = nullptr;
        if (ctx->list_cap > cap) cap = ctx->list_cap;
        if (ctx->list_ctas > n_ctas) n_ctas = ctx->list_ctas;
    }
    const size_t areas = (size_t)n_ctas * kListAreas;
    if (cudaMalloc(&ctx->d_list_entries, areas * cap * sizeof(uint2) + kListSlackBytes) != cudaSuccess) {
        cudaGetLastError();
        ctx->d_list_entries = nullptr;
        ctx->lists_alloc_failed = true;
        ctx->list_cap = 0;
        ctx->list_ctas = 0;
        return;
    }
    // The quad passes load two trips ahead without clamping to the end of a list (cvo_quads.cuh): what they read past it
    // is never used, but it should be defined memory (compute-sanitizer initcheck stays meaningful).  Once per allocation.
    cudaMemsetAsync(ctx->d_list_entries, 0, areas * cap * sizeof(uint2) + kListSlackBytes, ctx->stream);
    ctx->list_cap = (unsigned)cap;
    ctx->list_ctas = n_ctas;
}


// glibc's rand() after srand(seed) (TYPE_3 additive feedback generator, r[i] = r[i-3] + r[i-31]), restated so that
// the selector's randomPattern (thirdparty/PixelSelector2.cpp:36-38: srand(3141592); rand() & 0xFF) can be produced
// without touching the process-wide C random state.  tests/test_pcd_frontend.py checks it against libc.
void glibc_rand_bytes(unsigned seed, size_t n, uint8_t* out) {
    std::vector<int32_t> r(344 + n);
    r[0] = (int32_t)(seed == 0 ? 1u : seed);  // srandom: "we must make sure the seed is not 0"
    for (int i = 1; i < 31; ++i) {
        const int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        r[i] = (int32_t)word;
    }
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    for (size_t i = 34; i < 344 + n; ++i) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
    for (size_t k = 0; k < n; ++k) out[k] = (uint8_t)(((uint32_t)r[k + 344] >> 1) & 0xFF);
}

CamInfo camera_info(int dataset_seq) {  // src/pcd_generator.cpp:241-302
    switch (dataset_seq) {
        case 1: return {5000.0f, 517.3f, 516.5f, 318.6f, 255.3f};
        case 2: return {5000.0f, 520.9f, 521.0f, 325.1f, 249.7f};
        case 3: return {5000.0f, 535.4f, 539.2f, 320.1f, 247.6f};
        case 4: return {2000.0f, 718.856f, 718.856f, 607.1928f, 185.2157f};
        case 5: return {2000.0f, 707.0912f, 707.0912f, 601.8873f, 183.1104f};
        default: return {1000.0f, 616.368f, 616.745f, 319.935f, 243.639f};
    }
}

void free_image_pipe(cvo_b200_ctx* ctx) {
    cvo_b200_ctx::ImagePipe& P = ctx->pipe;
    cudaFree(P.d_img3); cudaFree(P.d_depth); cudaFree(P.d_pyr); cudaFree(P.d_ths); cudaFree(P.d_map); cudaFree(P.d_rnd); cudaFree(P.d_blockcnt);
    cudaFree(P.d_ctl);
    if (P.h_ctl) cudaFreeHost(P.h_ctl);
    cudaFree(P.d_next_raw_xyz); cudaFree(P.d_next_raw_feat); cudaFree(P.d_next_g); cudaFree(P.d_next_f); cudaFree(P.d_next_f4);
    cudaFree(P.d_next_job);
    if (P.h_ctl_next) cudaFreeHost(P.h_ctl_next);
    if (P.ev_next) cudaEventDestroy(P.ev_next);
    if (P.ev_taken) cudaEventDestroy(P.ev_taken);
    P = cvo_b200_ctx::ImagePipe();
}

size_t pyr_floats(int w, int h) {
    size_t n = 0;
    for (int l = 0; l < 3; ++l) {
        n += (size_t)w * h;
        w /= 2;
        h /= 2;
    }
    return n;
}

int ensure_image_pipe(cvo_b200_ctx* ctx, int w, int h) {
    cvo_b200_ctx::ImagePipe& P = ctx->pipe;
    if (P.w == w && P.h == h) return CVO_B200_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
    free_image_pipe(ctx);
    const size_t wh = (size_t)w * h;
    const size_t mp = (size_t)ctx->max_points;
    CK(cudaMalloc(&P.d_next_raw_xyz, mp * 3 * sizeof(float)));
    CK(cudaMalloc(&P.d_next_raw_feat, mp * 5 * sizeof(float)));
    CK(cudaMalloc(&P.d_next_g, mp * sizeof(float4)));
    CK(cudaMalloc(&P.d_next_f, mp * sizeof(float4)));
    CK(cudaMalloc(&P.d_next_f4, mp * sizeof(float)));
    CK(cudaMalloc(&P.d_next_job, sizeof(PackJob)));
    CK(cudaMallocHost(&P.h_ctl_next, sizeof(SelCtl)));
    CK(cudaEventCreateWithFlags(&P.ev_next, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&P.ev_taken, cudaEventDisableTiming));
    CK(cudaMalloc(&P.d_img3, wh * 3));
    CK(cudaMalloc(&P.d_depth, wh * sizeof(uint16_t)));
    CK(cudaMalloc(&P.d_pyr, pyr_floats(w, h) * 4 * sizeof(float)));
    CK(cudaMalloc(&P.d_ths, (size_t)(w / 32) * (h / 32) * 2 * sizeof(float)));
    CK(cudaMalloc(&P.d_map, wh));
    CK(cudaMalloc(&P.d_rnd, wh));
    CK(cudaMalloc(&P.d_blockcnt, ((wh + 1023) / 1024) * sizeof(int)));
    CK(cudaMalloc(&P.d_ctl, sizeof(SelCtl)));
    CK(cudaMallocHost(&P.h_ctl, sizeof(SelCtl)));
    std::vector<uint8_t> rnd(wh);
    glibc_rand_bytes(3141592u, wh, rnd.data());
    CK(cudaMemcpy(P.d_rnd, rnd.data(), wh, cudaMemcpyHostToDevice));
    int sdiv[256], hdiv[256];  // OpenCV's 8-bit HSV tables (hsv_shift = 12), round half to even like cv::saturate_cast
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
        sdiv[i] = (int)lrint((255 << 12) / (1. * i));
        hdiv[i] = (int)lrint((180 << 12) / (6. * i));
    }
    CK(cudaMemcpyToSymbol(c_sdiv, sdiv, sizeof(sdiv)));
    CK(cudaMemcpyToSymbol(c_hdiv, hdiv, sizeof(hdiv)));
    P.w = w;
    P.h = h;
    return CVO_B200_OK;
}

PairDev make_pair_dev(cvo_b200_ctx* ctx, int slot) {
    const cvo_b200_ctx::Slot& s = ctx->slots[slot];
    PairDev pd;
    const int fb = s.fixed_buf, mb = 1 - s.fixed_buf;
    pd.x.g = slot_g(ctx, slot, fb);
    pd.x.f = slot_f(ctx, slot, fb);
    pd.x.f4 = slot_f4(ctx, slot, fb);
    pd.x.n = s.n[fb];
    pd.x.pad = 0;
    pd.y.g = slot_g(ctx, slot, mb);
    pd.y.f = slot_f(ctx, slot, mb);
    pd.y.f4 = slot_f4(ctx, slot, mb);
    pd.y.n = s.n[mb];
    pd.y.pad = 0;
    return pd;
}

// The launch half of an align: pair descriptors and states into the pinned arrays the kernel reads, scratch, launch.
int run_align_begin(cvo_b200_ctx* ctx, const int* slots, int n_pairs, const cvo_b200_params* p, const float* RT_io,
                    const float* ell_io, bool trace, int trace_cap) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (ctx->pending.active) return fail_arg(ctx, "an align is in flight (cvo_b200_align_finish first)");
    if (!slots || !p || n_pairs <= 0 || n_pairs > ctx->max_slots) return fail_arg(ctx, "bad align arguments");
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < n_pairs; ++i) {
        const int s = slots[i];
        if (s < 0 || s >= ctx->max_slots || !ctx->slots[s].bound) return fail_arg(ctx, "slot not bound");
        if (ctx->slots[s].pending_batch >= 0) {
            const int rc = flush_batch(ctx, ctx->slots[s].pending_batch);
            if (rc) return rc;
        }
    }
    // The clusters pull pairs from a queue: the largest pairs go first (N x M, stable), so that a ragged batch ends on its
    // cheap pairs -- the last wave is what a batch of a few pairs per cluster waits for.  Results go back by `order`.
    // (A traced align keeps the caller's pair 0 at the head of the queue: the kernel traces queue position 0.)
    ctx->order.resize(n_pairs);
    for (int i = 0; i < n_pairs; ++i) ctx->order[i] = i;
    std::stable_sort(ctx->order.begin() + ((trace && trace_cap > 0) ? 1 : 0), ctx->order.end(), [&](int a, int b) {
        const cvo_b200_ctx::Slot& sa = ctx->slots[slots[a]];
        const cvo_b200_ctx::Slot& sb = ctx->slots[slots[b]];
        return (long long)sa.n[0] * sa.n[1] > (long long)sb.n[0] * sb.n[1];
    });
    for (int q = 0; q < n_pairs; ++q) {
        const int i = ctx->order[q];
        ctx->h_pairs[q] = make_pair_dev(ctx, slots[i]);
        PairState& st = ctx->h_states[q];
        memset(&st, 0, sizeof(st));
        if (RT_io) {
            memcpy(st.R, RT_io + (size_t)i * 12, sizeof(float) * 9);
            memcpy(st.T, RT_io + (size_t)i * 12 + 9, sizeof(float) * 3);
        } else {
            st.R[0] = st.R[4] = st.R[8] = 1.f;
        }
        st.ell = ell_io ? ell_io[i] : p->ell_init;
        st.ell_max = p->ell_max;
    }
    AlignArgs args;
    args.pairs = ctx->h_pairs_dev;
    args.states = ctx->h_states_dev;
    args.n_pairs = n_pairs;
    args.counter = ctx->d_counter;
    args.trace = nullptr;
    args.trace_cap = 0;
    if (trace && trace_cap > 0) {
        args.trace = ctx->d_trace;
        args.trace_cap = trace_cap < ctx->trace_cap ? trace_cap : ctx->trace_cap;
    }
    args.kp = make_kparams(p, false);
    int G = choose_cluster(ctx, n_pairs, p->fixed_iters <= 0);
    int group = 1;
    {
        int fit = max_resident_clusters(ctx, G);
        if (fit < 1 && G > 8 && ctx->force_G == 0) {  // 16-CTA clusters are opt-in; fall back to portable 8
            G = 8;
            fit = max_resident_clusters(ctx, G);
        }
        if (fit < 1) {
            ctx->err = "cluster size not schedulable on this device";
            return CVO_B200_ERR_CUDA;
        }
        int max_n = 0;
        for (int i = 0; i < n_pairs; ++i) {
            max_n = ctx->h_pairs[i].x.n > max_n ? ctx->h_pairs[i].x.n : max_n;
            max_n = ctx->h_pairs[i].y.n > max_n ? ctx->h_pairs[i].y.n : max_n;
        }
        // Whole-GPU mode (group_allreduce in cvo_kernels.cuh): a single pair whose clouds give every CTA of one cluster
        // eight or more row tiles is spread over every cluster the device holds.  Below that the iteration is dominated
        // by what does not split -- the serial section, the column staging every CTA repeats, the barriers -- and the
        // two extra global-memory barriers cost more than the shorter passes save (measured, profiles/r02_group_mode.txt:
        // 3000 points 20.5 -> 21.8 us per iteration, 10000 points 75.9 -> 46.2).
        if (ctx->force_group > 1) group = ctx->force_group;
        else if (ctx->force_group == 0 && n_pairs == 1 && max_n / kTile >= 8 * G) group = fit;
        if (group > fit) group = fit;
        if (group > kMaxGroupClusters) group = kMaxGroupClusters;
        if (group < 1) group = 1;
        ensure_list_scratch(ctx, (group > 1 ? group : (n_pairs < fit ? n_pairs : fit)) * G, max_n);
    }
    args.group_clusters = group;
    args.group_xchg = ctx->d_group_xchg;
    args.group_count = reinterpret_cast<unsigned*>(ctx->d_counter + 1);
    // the list passes rely on a > sp_thres implying the ell-ball test, which holds for c_sigma^2 <= 1 (cvo_quads.cuh)
    const bool lists = ctx->lists_enabled && ctx->d_list_entries != nullptr && args.kp.cs2 <= 1.0f;
    args.list_entries = lists ? ctx->d_list_entries : nullptr;
    args.list_cap = ctx->list_cap;
    args.list_skin = ctx->list_skin;
    args.list_skin_min = ctx->list_skin_min;
    args.list_ahead = ctx->list_ahead;
    args.list_wide = ctx->list_wide >= 0.f ? ctx->list_wide : (args.kp.mode == CVO_B200_MODE_ACVO ? 0.5f : 0.f);
    args.list_shrink = ctx->list_shrink;
    args.list_refine_min = ctx->list_refine_min;
    // The pair descriptors and states live in pinned host memory that the kernel reads and writes directly (unified
    // addressing: 272 B per pair, once at its start and once at its end).  No host->device copy sits in the launch
    // path: a small copy would queue behind the upload of the NEXT batch on the copy engine (measured: 0.8 ms).
    CK(cudaMemsetAsync(ctx->d_counter, 0, 2 * sizeof(int), ctx->stream));
    int ncl = 0;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = launch_cluster_kernel(ctx, align_kernel, args, G, group > 1 ? group : n_pairs, &ncl);
    if (rc != CVO_B200_OK) return rc;
    if (lists && ncl * G > ctx->list_ctas) {  // cannot happen: both sides use the same occupancy query
        ctx->err = "internal: launch larger than the neighbour-list scratch";
        return CVO_B200_ERR_CUDA;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    if (args.trace)
        CK(cudaMemcpyAsync(ctx->h_trace, ctx->d_trace, sizeof(cvo_b200_iter_rec) * args.trace_cap,
                           cudaMemcpyDeviceToHost, ctx->stream));
    ctx->pending.active = true;
    ctx->pending.n_pairs = n_pairs;
    ctx->pending.G = G;
    ctx->pending.ncl = ncl;
    ctx->pending.group = group;
    ctx->pending.trace = args.trace != nullptr;
    ctx->pending.trace_cap = args.trace_cap;
    return CVO_B200_OK;
}

// The collecting half: waits for the kernel, hands the results out.
int run_align_finish(cvo_b200_ctx* ctx, float* RT_io, float* ell_io, float* transform, float* prev_transform, int* iters,
                     int* status, cvo_b200_iter_rec* trace, int* trace_len) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (!ctx->pending.active) return fail_arg(ctx, "no align in flight (cvo_b200_align_begin)");
    CK(cudaSetDevice(ctx->device));
    const int n_pairs = ctx->pending.n_pairs, G = ctx->pending.G, ncl = ctx->pending.ncl, group = ctx->pending.group;
    ctx->pending.active = false;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    ctx->last_G = G;
    ctx->last_nclusters = ncl;
    ctx->last_group = group;
    long long total = 0, builds = 0, refines = 0, xy_entries = 0, xy_slots = 0;
    for (int q = 0; q < n_pairs; ++q) {
        const PairState& st = ctx->h_states[q];
        const int i = ctx->order[q];
        total += st.n_run;
        builds += st.n_builds;
        refines += st.n_refines;
        xy_entries += st.xy_entries;
        xy_slots += st.xy_slots;
        if (RT_io) {
            memcpy(RT_io + (size_t)i * 12, st.R, sizeof(float) * 9);
            memcpy(RT_io + (size_t)i * 12 + 9, st.T, sizeof(float) * 3);
        }
        if (ell_io) ell_io[i] = st.ell;
        if (transform) memcpy(transform + (size_t)i * 16, st.tf, sizeof(float) * 16);
        if (prev_transform) memcpy(prev_transform + (size_t)i * 16, st.prev_tf, sizeof(float) * 16);
        if (iters) iters[i] = st.iters;
        if (status) status[i] = st.status;
    }
    ctx->last_total_iters = total;
    ctx->last_list_builds = builds;
    ctx->last_list_refines = refines;
    ctx->last_xy_entries = xy_entries;
    ctx->last_xy_slots = xy_slots;
    if (ctx->pending.trace && trace) {
        const int n = ctx->h_states[0].n_run < ctx->pending.trace_cap ? ctx->h_states[0].n_run : ctx->pending.trace_cap;
        memcpy(trace, ctx->h_trace, sizeof(cvo_b200_iter_rec) * n);
    }
    if (trace_len) *trace_len = ctx->h_states[0].n_run;
    return CVO_B200_OK;
}

int run_align(cvo_b200_ctx* ctx, const int* slots, int n_pairs, const cvo_b200_params* p, float* RT_io,
              float* ell_io, float* transform, float* prev_transform, int* iters, int* status,
              cvo_b200_iter_rec* trace, int trace_cap, int* trace_len) {
    const int rc = run_align_begin(ctx, slots, n_pairs, p, RT_io, ell_io, trace != nullptr, trace_cap);
    if (rc) return rc;
    return run_align_finish(ctx, RT_io, ell_io, transform, prev_transform, iters, status, trace, trace_len);
}

}  // namespace

extern "C" {

void cvo_b200_default_params_cvo(cvo_b200_params* p) {  // src/cvo.cpp:18-48
    memset(p, 0, sizeof(*p));
    p->mode = CVO_B200_MODE_CVO;
    p->ell_policy = CVO_B200_ELL_SCHEDULE;
    p->ell_init = 0.15f;
    p->ell_min = 0.0391f;
    p->ell_max = 0.15f;
    p->dl_step = 0.3;
    p->sigma = 0.1f;
    p->sp_thres = 8e-3f;
    p->c = 7.0f;
    p->d = 7.0f;
    p->c_ell = 200.f;
    p->c_sigma = 1.f;
    p->c_sp_thres = 8e-3f;
    p->max_iter = 2000;
    p->min_step = (float)(2 * 1.0e-1);
    p->max_step = 0.8f;
    p->eps = (float)(5 * 1.0e-5);
    p->eps_2 = 1.0e-5f;
    p->fixed_iters = 0;
}

void cvo_b200_default_params_acvo(cvo_b200_params* p) {  // src/adaptive_cvo.cpp:18-50
    cvo_b200_default_params_cvo(p);
    p->mode = CVO_B200_MODE_ACVO;
    p->ell_policy = CVO_B200_ELL_ADAPTIVE;
    p->ell_init = 0.1f;
    p->sp_thres = 8.315e-3f;
    p->c_ell = 0.5f;
    p->c_sp_thres = 8.315e-3f;
}

int cvo_b200_create(cvo_b200_ctx** out, int device, int max_points, int max_slots) {
    if (!out || max_points <= 0 || max_slots <= 0) return CVO_B200_ERR_ARG;
    *out = nullptr;
    cvo_b200_ctx* ctx = new cvo_b200_ctx();
    ctx->device = device;
    ctx->max_points = (max_points + 31) / 32 * 32;
    ctx->max_slots = max_slots;
    ctx->slots.resize(max_slots);
    ctx->trace_cap = kTraceCap;
    cudaError_t e;
#define CKC(call)                                                                   \
    do {                                                                            \
        e = (call);                                                                 \
        if (e != cudaSuccess) {                                                     \
            fprintf(stderr, "cvo_b200_create: %s: %s\n", #call, cudaGetErrorString(e)); \
            cvo_b200_destroy(ctx);                                                  \
            return CVO_B200_ERR_CUDA;                                               \
        }                                                                           \
    } while (0)
    CKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) {
        fprintf(stderr, "cvo_b200_create: device %d is sm_%d%d; this library is built for sm_100a only\n", device,
                prop.major, prop.minor);
        cvo_b200_destroy(ctx);
        return CVO_B200_ERR_CUDA;
    }
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CKC(cudaEventCreate(&ctx->ev0));
    CKC(cudaEventCreate(&ctx->ev1));
    const size_t mp = ctx->max_points;
    CKC(cudaMalloc(&ctx->d_raw_xyz, sizeof(float) * 3 * mp * 2));
    CKC(cudaMalloc(&ctx->d_raw_feat, sizeof(float) * 5 * mp * 2));
    CKC(cudaMalloc(&ctx->d_pk_g, sizeof(float4) * mp * 2 * max_slots));
    CKC(cudaMalloc(&ctx->d_pk_f, sizeof(float4) * mp * 2 * max_slots));
    CKC(cudaMalloc(&ctx->d_pk_f4, sizeof(float) * mp * 2 * max_slots));
    // the TMA copies move whole 32-point tiles: the padding behind a cloud's last point must be readable, finite data
    CKC(cudaMemset(ctx->d_pk_g, 0, sizeof(float4) * mp * 2 * max_slots));
    CKC(cudaMemset(ctx->d_pk_f, 0, sizeof(float4) * mp * 2 * max_slots));
    CKC(cudaMemset(ctx->d_pk_f4, 0, sizeof(float) * mp * 2 * max_slots));
    CKC(cudaHostAlloc(&ctx->h_pairs, sizeof(PairDev) * max_slots, cudaHostAllocMapped));
    CKC(cudaHostAlloc(&ctx->h_states, sizeof(PairState) * max_slots, cudaHostAllocMapped));
    CKC(cudaHostGetDevicePointer(&ctx->h_pairs_dev, ctx->h_pairs, 0));
    CKC(cudaHostGetDevicePointer(&ctx->h_states_dev, ctx->h_states, 0));
    CKC(cudaMalloc(&ctx->d_counter, 2 * sizeof(int)));
    CKC(cudaMalloc(&ctx->d_group_xchg, sizeof(double) * 2 * kMaxGroupClusters * kNumAcc));
    CKC(cudaMalloc(&ctx->d_trace, sizeof(cvo_b200_iter_rec) * kTraceCap));
    CKC(cudaMallocHost(&ctx->h_trace, sizeof(cvo_b200_iter_rec) * kTraceCap));
    CKC(cudaMalloc(&ctx->d_inner, sizeof(double) * 2));
    CKC(cudaMalloc(&ctx->d_jobs, sizeof(PackJob) * 2));
    CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        CKC(cudaMalloc(&ctx->batch[b].d_jobs, sizeof(PackJob) * 2 * max_slots));
        CKC(cudaMallocHost(&ctx->batch[b].h_jobs, sizeof(PackJob) * 2 * max_slots));
        CKC(cudaEventCreateWithFlags(&ctx->batch[b].copied, cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&ctx->batch[b].packed, cudaEventDisableTiming));
    }
    CKC(cudaMallocHost(&ctx->h_inner, sizeof(double) * 2));
    ctx->pack_smem_max = 128 * 1024;
    CKC(cudaFuncSetAttribute(pack_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->pack_smem_max));
    if ((size_t)ctx->max_points * sizeof(unsigned long long) > ctx->pack_smem_max) {
        fprintf(stderr, "cvo_b200_create: max_points %d exceeds the single-CTA sort capacity (16384)\n", max_points);
        cvo_b200_destroy(ctx);
        return CVO_B200_ERR_ARG;
    }
    const char* env = getenv("CVO_B200_NO_SORT");
    if (env && env[0] == '1') ctx->sort_points = 0;
    env = getenv("CVO_B200_NO_LISTS");  // tuning / A-B switch; cvo_b200_set_neighbor_lists is the API
    if (env && env[0] == '1') ctx->lists_enabled = false;
    env = getenv("CVO_B200_LIST_SHRINK");  // tuning knob
    if (env && atof(env) > 0.0 && atof(env) <= 1.0) ctx->list_shrink = (float)atof(env);
    env = getenv("CVO_B200_LIST_REFINE_MIN");
    if (env && atof(env) > 0.0) ctx->list_refine_min = (float)atof(env);
    env = getenv("CVO_B200_LIST_WIDE");
    if (env && atof(env) >= 0.0 && atof(env) <= 4.0) ctx->list_wide = (float)atof(env);
    env = getenv("CVO_B200_LIST_AHEAD");
    if (env && atof(env) >= 0.0 && atof(env) < 1.0) ctx->list_ahead = (float)atof(env);
    env = getenv("CVO_B200_LIST_SKIN_MIN");
    if (env && atof(env) >= 0.0 && atof(env) <= 1.0) ctx->list_skin_min = (float)atof(env);
    env = getenv("CVO_B200_LIST_SKIN");
    if (env && atof(env) >= 0.0 && atof(env) <= 1.0) ctx->list_skin = (float)atof(env);
#undef CKC
    *out = ctx;
    return CVO_B200_OK;
}

void cvo_b200_destroy(cvo_b200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_raw_xyz);
    cudaFree(ctx->d_raw_feat);
    cudaFree(ctx->d_pk_g);
    cudaFree(ctx->d_pk_f);
    cudaFree(ctx->d_pk_f4);
    cudaFreeHost(ctx->h_pairs);
    cudaFreeHost(ctx->h_states);
    cudaFree(ctx->d_counter);
    cudaFree(ctx->d_group_xchg);
    cudaFree(ctx->d_trace);
    cudaFreeHost(ctx->h_trace);
    cudaFree(ctx->d_inner);
    cudaFree(ctx->d_jobs);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int b = 0; b < 2; ++b) {
        cudaFree(ctx->batch[b].d_raw);
        cudaFree(ctx->batch[b].d_jobs);
        cudaFreeHost(ctx->batch[b].h_jobs);
        if (ctx->batch[b].copied) cudaEventDestroy(ctx->batch[b].copied);
        if (ctx->batch[b].packed) cudaEventDestroy(ctx->batch[b].packed);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    free_image_pipe(ctx);
    cudaFree(ctx->d_list_entries);
    cudaFreeHost(ctx->h_inner);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* cvo_b200_last_error(const cvo_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int cvo_b200_set_pair(cvo_b200_ctx* ctx, int slot, const float* fixed_xyz, const float* fixed_feat, int n_fixed,
                      const float* moving_xyz, const float* moving_feat, int n_moving) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (slot < 0 || slot >= ctx->max_slots) return fail_arg(ctx, "slot out of range");
    if (!fixed_xyz || !fixed_feat || !moving_xyz || !moving_feat) return fail_arg(ctx, "null cloud pointer");
    if (n_fixed <= 0 || n_moving <= 0) {
        ctx->err = "empty cloud";
        return CVO_B200_ERR_EMPTY;
    }
    if (n_fixed > ctx->max_points || n_moving > ctx->max_points) return fail_arg(ctx, "cloud larger than max_points");
    CK(cudaSetDevice(ctx->device));
    int rc = flush_all_batches(ctx);
    if (rc) return rc;
    rc = upload_cloud(ctx, 0, fixed_xyz, fixed_feat, n_fixed);
    if (rc) return rc;
    rc = upload_cloud(ctx, 1, moving_xyz, moving_feat, n_moving);
    if (rc) return rc;
    cvo_b200_ctx::Slot& s = ctx->slots[slot];
    s.fixed_buf = 0;
    s.n[0] = n_fixed;
    s.n[1] = n_moving;
    s.bound = true;
    s.have_fixed = true;
    PackJob jobs[2];
    const size_t mp = ctx->max_points;
    jobs[0] = {ctx->d_raw_xyz, ctx->d_raw_feat, slot_g(ctx, slot, 0), slot_f(ctx, slot, 0), slot_f4(ctx, slot, 0), n_fixed, 0};
    jobs[1] = {ctx->d_raw_xyz + mp * 3, ctx->d_raw_feat + mp * 5, slot_g(ctx, slot, 1), slot_f(ctx, slot, 1), slot_f4(ctx, slot, 1), n_moving, 0};
    return launch_pack(ctx, jobs, 2, ctx->d_jobs);
}

// cvo_b200_set_pairs with the batch's pairs `pair_step` pairs apart in the caller's arrays (1: contiguous; W: every
// W-th pair, the share of one of W devices in cvo_b200_align_multi).  The counts are indexed the same way.
static int set_pairs_strided(cvo_b200_ctx* ctx, const int* slots, int n_pairs, const float* fixed_xyz,
                             const float* fixed_feat, const int* n_fixed, const float* moving_xyz, const float* moving_feat,
                             const int* n_moving, int stride_points, int pair_step) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (!slots || !fixed_xyz || !fixed_feat || !moving_xyz || !moving_feat || !n_fixed || !n_moving)
        return fail_arg(ctx, "null pointer");
    if (n_pairs <= 0 || n_pairs > ctx->max_slots) return fail_arg(ctx, "bad pair count");
    if (stride_points <= 0) return fail_arg(ctx, "bad stride");
    if (pair_step < 1) return fail_arg(ctx, "bad pair step");
    for (int i = 0; i < n_pairs; ++i) {
        const int nf = n_fixed[(size_t)i * pair_step], nm = n_moving[(size_t)i * pair_step];
        if (slots[i] < 0 || slots[i] >= ctx->max_slots) return fail_arg(ctx, "slot out of range");
        if (nf <= 0 || nm <= 0) {
            ctx->err = "empty cloud";
            return CVO_B200_ERR_EMPTY;
        }
        if (nf > ctx->max_points || nm > ctx->max_points || nf > stride_points || nm > stride_points)
            return fail_arg(ctx, "cloud larger than max_points / stride");
    }
    CK(cudaSetDevice(ctx->device));
    const int b = ctx->next_batch;
    ctx->next_batch ^= 1;
    cvo_b200_ctx::Batch& B = ctx->batch[b];
    if (B.pending) {  // a batch nobody consumed: its pack still has to run before its staging area is reused
        const int rc = flush_batch(ctx, b);
        if (rc) return rc;
    }
    if (B.used) {
        CK(cudaEventSynchronize(B.copied));                    // the pinned job array may be rewritten
        CK(cudaStreamWaitEvent(ctx->copy_stream, B.packed, 0));  // the staging area may be overwritten
    }
    // one staging area for the whole batch: [fixed xyz | fixed feat | moving xyz | moving feat]
    const size_t cloud3 = (size_t)stride_points * 3, cloud5 = (size_t)stride_points * 5;
    const size_t need = (size_t)n_pairs * 2 * (cloud3 + cloud5);
    if (need > B.raw_floats) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->copy_stream));
        cudaFree(B.d_raw);
        B.d_raw = nullptr;
        B.raw_floats = 0;
        CK(cudaMalloc(&B.d_raw, need * sizeof(float)));
        B.raw_floats = need;
    }
    float* d_fx = B.d_raw;
    float* d_ff = d_fx + (size_t)n_pairs * cloud3;
    float* d_mx = d_ff + (size_t)n_pairs * cloud5;
    float* d_mf = d_mx + (size_t)n_pairs * cloud3;
    B.slots.assign(slots, slots + n_pairs);
    for (int i = 0; i < n_pairs; ++i) {
        const int slot = slots[i];
        cvo_b200_ctx::Slot& s = ctx->slots[slot];
        if (s.pending_batch >= 0 && s.pending_batch != b) {  // an older, unconsumed upload of this slot goes first
            const int rc = flush_batch(ctx, s.pending_batch);
            if (rc) return rc;
        }
        const int nf = n_fixed[(size_t)i * pair_step], nm = n_moving[(size_t)i * pair_step];
        s.fixed_buf = 0;
        s.n[0] = nf;
        s.n[1] = nm;
        s.bound = true;
        s.have_fixed = true;
        s.pending_batch = b;
        B.h_jobs[2 * i] = {d_fx + i * cloud3, d_ff + i * cloud5, slot_g(ctx, slot, 0), slot_f(ctx, slot, 0),
                           slot_f4(ctx, slot, 0), nf, 0};
        B.h_jobs[2 * i + 1] = {d_mx + i * cloud3, d_mf + i * cloud5, slot_g(ctx, slot, 1), slot_f(ctx, slot, 1),
                               slot_f4(ctx, slot, 1), nm, 0};
        if ((size_t)nf * sizeof(unsigned long long) > ctx->pack_smem_max ||
            (size_t)nm * sizeof(unsigned long long) > ctx->pack_smem_max)
            return fail_arg(ctx, "cloud too large for the single-CTA sort");
    }
    B.njobs = 2 * n_pairs;
    cudaStream_t cs = ctx->copy_stream;
    if (pair_step == 1) {
        CK(cudaMemcpyAsync(d_fx, fixed_xyz, sizeof(float) * n_pairs * cloud3, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpyAsync(d_ff, fixed_feat, sizeof(float) * n_pairs * cloud5, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpyAsync(d_mx, moving_xyz, sizeof(float) * n_pairs * cloud3, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpyAsync(d_mf, moving_feat, sizeof(float) * n_pairs * cloud5, cudaMemcpyHostToDevice, cs));
    } else {  // every pair_step-th cloud: a 2-D copy, one row per pair
        const size_t w3 = sizeof(float) * cloud3, w5 = sizeof(float) * cloud5;
        CK(cudaMemcpy2DAsync(d_fx, w3, fixed_xyz, w3 * pair_step, w3, n_pairs, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpy2DAsync(d_ff, w5, fixed_feat, w5 * pair_step, w5, n_pairs, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpy2DAsync(d_mx, w3, moving_xyz, w3 * pair_step, w3, n_pairs, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpy2DAsync(d_mf, w5, moving_feat, w5 * pair_step, w5, n_pairs, cudaMemcpyHostToDevice, cs));
    }
    CK(cudaMemcpyAsync(B.d_jobs, B.h_jobs, sizeof(PackJob) * B.njobs, cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(B.copied, cs));
    B.pending = true;
    B.used = true;
    B.packed_eager = false;
    if (!ctx->pending.active) {  // no align in flight that could be reading these slots: pack right behind the copies
        const int rc = launch_pack(ctx, B, cs);
        if (rc) return rc;
        B.packed_eager = true;
    }
    return CVO_B200_OK;
}

static int push_cloud(cvo_b200_ctx* ctx, int slot, const float* xyz, const float* feat, int n, bool promote) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (slot < 0 || slot >= ctx->max_slots || !ctx->slots[slot].bound) return fail_arg(ctx, "slot not bound");
    if (!xyz || !feat) return fail_arg(ctx, "null cloud pointer");
    if (n <= 0) {
        ctx->err = "empty cloud";
        return CVO_B200_ERR_EMPTY;
    }
    if (n > ctx->max_points) return fail_arg(ctx, "cloud larger than max_points");
    CK(cudaSetDevice(ctx->device));
    int rc = flush_all_batches(ctx);
    if (rc) return rc;
    cvo_b200_ctx::Slot& s = ctx->slots[slot];
    // promote: moving becomes fixed (src/cvo.cpp:417) and the new cloud lands in the old fixed buffer; otherwise the
    // moving cloud is replaced in place (two set_pcd() calls without an align() in between, src/cvo.cpp:336-351).
    // The slot's bookkeeping only changes once the upload and the pack launch have been enqueued successfully.
    const int fixed_buf = promote ? 1 - s.fixed_buf : s.fixed_buf;
    const int mb = 1 - fixed_buf;
    rc = upload_cloud(ctx, 0, xyz, feat, n);
    if (rc) return rc;
    PackJob job = {ctx->d_raw_xyz, ctx->d_raw_feat, slot_g(ctx, slot, mb), slot_f(ctx, slot, mb), slot_f4(ctx, slot, mb), n, 0};
    rc = launch_pack(ctx, &job, 1, ctx->d_jobs);
    if (rc) return rc;
    s.fixed_buf = fixed_buf;
    s.n[mb] = n;
    return CVO_B200_OK;
}

int cvo_b200_set_pairs(cvo_b200_ctx* ctx, const int* slots, int n_pairs, const float* fixed_xyz,
                       const float* fixed_feat, const int* n_fixed, const float* moving_xyz, const float* moving_feat,
                       const int* n_moving, int stride_points) {
    return set_pairs_strided(ctx, slots, n_pairs, fixed_xyz, fixed_feat, n_fixed, moving_xyz, moving_feat, n_moving, stride_points, 1);
}

int cvo_b200_align_multi(cvo_b200_ctx* const* ctxs, int n_ctx, int n_pairs, const float* fixed_xyz, const float* fixed_feat,
                         const int* n_fixed, const float* moving_xyz, const float* moving_feat, const int* n_moving,
                         int stride_points, const cvo_b200_params* p, float* transform, int* iters, int* status,
                         float* kernel_ms) {
    if (!ctxs || n_ctx <= 0 || n_pairs <= 0 || !p || !transform) return CVO_B200_ERR_ARG;
    for (int c = 0; c < n_ctx; ++c)
        if (!ctxs[c]) return CVO_B200_ERR_ARG;
    std::vector<int> rc(n_ctx, CVO_B200_OK);
    // pair q -> context q mod n_ctx; every context works through its share in chunks of its slot count on its own
    // host thread (one ctx = one caller thread); results go straight to index q of the caller's arrays: that is the
    // gather (one process holds all devices, so no collective is involved; across processes see sharding.py).
    auto work = [&](int c) {
        cvo_b200_ctx* ctx = ctxs[c];
        const int mine = (n_pairs - c + n_ctx - 1) / n_ctx;
        std::vector<int> slots(ctx->max_slots);
        for (int i = 0; i < ctx->max_slots; ++i) slots[i] = i;
        std::vector<float> tf;
        std::vector<int> it, st;
        float ms = 0.f;
        for (int done = 0; done < mine && rc[c] == CVO_B200_OK; done += ctx->max_slots) {
            const int n = mine - done < ctx->max_slots ? mine - done : ctx->max_slots;
            const size_t first = (size_t)c + (size_t)done * n_ctx;  // global index of the chunk's first pair
            const size_t c3 = (size_t)stride_points * 3, c5 = (size_t)stride_points * 5;
            rc[c] = set_pairs_strided(ctx, slots.data(), n, fixed_xyz + first * c3, fixed_feat + first * c5, n_fixed + first,
                                      moving_xyz + first * c3, moving_feat + first * c5, n_moving + first, stride_points, n_ctx);
            if (rc[c] != CVO_B200_OK) break;
            tf.resize((size_t)n * 16);
            it.resize(n);
            st.resize(n);
            rc[c] = cvo_b200_align(ctx, slots.data(), n, p, nullptr, nullptr, tf.data(), nullptr, it.data(), st.data());
            if (rc[c] != CVO_B200_OK) break;
            ms += ctx->last_ms;
            for (int i = 0; i < n; ++i) {
                const size_t q = first + (size_t)i * n_ctx;
                memcpy(transform + q * 16, tf.data() + (size_t)i * 16, sizeof(float) * 16);
                if (iters) iters[q] = it[i];
                if (status) status[q] = st[i];
            }
        }
        if (kernel_ms) kernel_ms[c] = ms;
    };
    if (n_ctx == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int c = 0; c < n_ctx; ++c) th.emplace_back(work, c);
        for (auto& t : th) t.join();
    }
    for (int c = 0; c < n_ctx; ++c)
        if (rc[c] != CVO_B200_OK) return rc[c];
    return CVO_B200_OK;
}

int cvo_b200_push_frame(cvo_b200_ctx* ctx, int slot, const float* xyz, const float* feat, int n) {
    return push_cloud(ctx, slot, xyz, feat, n, true);
}

int cvo_b200_replace_moving(cvo_b200_ctx* ctx, int slot, const float* xyz, const float* feat, int n) {
    return push_cloud(ctx, slot, xyz, feat, n, false);
}

static int check_image_args(cvo_b200_ctx* ctx, const unsigned char* img3, const unsigned short* depth, int width, int height,
                            int feature_type);
static int enqueue_frontend(cvo_b200_ctx* ctx, const unsigned char* img3, const unsigned short* depth, int width, int height,
                            int dataset_seq, int feature_type, float* raw_xyz, float* raw_feat, PackJob* d_job, float4* out_g,
                            float4* out_f, float* out_f4, cudaStream_t st, SelCtl* h_ctl);

static int push_images(cvo_b200_ctx* ctx, int slot, const unsigned char* img3, const unsigned short* depth, int width,
                       int height, int dataset_seq, int feature_type, int* num_points, bool promote) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (slot < 0 || slot >= ctx->max_slots) return fail_arg(ctx, "slot out of range");
    int rc = check_image_args(ctx, img3, depth, width, height, feature_type);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    rc = flush_all_batches(ctx);
    if (rc) return rc;
    rc = ensure_image_pipe(ctx, width, height);
    if (rc) return rc;
    cvo_b200_ctx::ImagePipe& P = ctx->pipe;
    cvo_b200_ctx::Slot& s = ctx->slots[slot];
    // where the new cloud goes: the first frame of a sequence is the fixed cloud (src/cvo.cpp:326-334), every later
    // one the moving cloud, the previous moving cloud having become the fixed one (:417)
    int fixed_buf = s.fixed_buf, target;
    if (!s.have_fixed) { fixed_buf = 0; target = 0; }
    else if (!s.bound) target = 1 - fixed_buf;
    else if (promote) { fixed_buf = 1 - fixed_buf; target = 1 - fixed_buf; }
    else target = 1 - fixed_buf;  // no align() since the last frame: the moving cloud is replaced (src/cvo.cpp:336-351)
    if (P.next_pending) {  // a prefetched frame is still in flight through the pipe's scratch: this frame replaces it
        CK(cudaEventSynchronize(P.ev_next));
        P.next_pending = false;
    }
    rc = enqueue_frontend(ctx, img3, depth, width, height, dataset_seq, feature_type, ctx->d_raw_xyz, ctx->d_raw_feat, ctx->d_jobs,
                          slot_g(ctx, slot, target), slot_f(ctx, slot, target), slot_f4(ctx, slot, target), ctx->stream, P.h_ctl);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    const SelCtl& c = *P.h_ctl;
    ctx->last_gen_n = c.num_points < ctx->max_points ? c.num_points : ctx->max_points;
    if (num_points) *num_points = c.num_points;
    ctx->last_gen_canny = c.canny_used;
    if (c.status == PCD_STATUS_TOO_MANY_POINTS) return fail_arg(ctx, "the frame yields more points than max_points");
    if (c.num_points <= 0) {
        ctx->err = "empty cloud";
        return CVO_B200_ERR_EMPTY;
    }
    s.fixed_buf = fixed_buf;
    s.n[target] = c.num_points;
    if (!s.have_fixed) s.have_fixed = true;
    else s.bound = true;
    return CVO_B200_OK;
}

// The device image front end (pcd_kernels.cuh) for one frame, enqueued on `st`: image + depth up, 23 launches, the packed
// cloud into (out_g, out_f, out_f4), the selector's control block back into the pinned `h_ctl`.  No synchronisation.
static int enqueue_frontend(cvo_b200_ctx* ctx, const unsigned char* img3, const unsigned short* depth, int width, int height,
                            int dataset_seq, int feature_type, float* raw_xyz, float* raw_feat, PackJob* d_job, float4* out_g,
                            float4* out_f, float* out_f4, cudaStream_t st, SelCtl* h_ctl) {
    cvo_b200_ctx::ImagePipe& P = ctx->pipe;
    const int w = width, h = height, wh = w * h;
    CK(cudaMemcpyAsync(P.d_img3, img3, (size_t)wh * 3, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(P.d_depth, depth, (size_t)wh * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    PcdBuffers B;
    B.w = w; B.h = h; B.img3 = P.d_img3; B.depth = P.d_depth;
    {
        float* p = P.d_pyr;
        int wl = w, hl = h;
        for (int l = 0; l < 3; ++l) {
            const size_t n = (size_t)wl * hl;
            B.I[l] = p; B.dx[l] = p + n; B.dy[l] = p + 2 * n; B.g2[l] = p + 3 * n;
            p += 4 * n;
            wl /= 2; hl /= 2;
        }
    }
    B.ths = P.d_ths;
    B.thsSmoothed = P.d_ths + (size_t)(w / 32) * (h / 32);
    B.map = P.d_map;
    B.randomPattern = P.d_rnd;
    B.ctl = P.d_ctl;
    const int T = 256;
    pcd_gray_kernel<<<(wh + T - 1) / T, T, 0, st>>>(B, 3000);  // num_want (src/pcd_generator.cpp:22)
    {
        int wl = w, hl = h;
        for (int l = 0; l < 3; ++l) {
            if (l > 0) pcd_down_kernel<<<(wl * hl + T - 1) / T, T, 0, st>>>(B.I[l - 1], B.I[l], wl, hl);
            pcd_grad_kernel<<<(wl * hl + T - 1) / T, T, 0, st>>>(B.I[l], B.dx[l], B.dy[l], B.g2[l], wl, hl);
            wl /= 2; hl /= 2;
        }
    }
    const int nb32 = (w / 32) * (h / 32);
    pcd_hist_kernel<<<nb32, 1024, 0, st>>>(B);
    pcd_smooth_kernel<<<(nb32 + T - 1) / T, T, 0, st>>>(B);
    const int max_blocks = ((w + 3) / 4) * ((h + 3) / 4);  // potential 1: 4 x 4 blocks
    for (int stage = 0; stage < 2; ++stage) {
        pcd_clear_map_kernel<<<(wh + T - 1) / T, T, 0, st>>>(B, stage);
        pcd_select_kernel<<<(max_blocks * 16 + T - 1) / T, T, 0, st>>>(B, stage);  // one thread per pot x pot cell
        pcd_decide_kernel<<<1, 1, 0, st>>>(B, stage);
    }
    const int nblk = (wh + 1023) / 1024;
    pcd_count_kernel<FLAG_SELECTED><<<nblk, 1024, 0, st>>>(B, P.d_blockcnt);
    pcd_subsample_kernel<<<nblk, 1024, 0, st>>>(B, P.d_blockcnt);
    pcd_after_subsample_kernel<<<1, 1, 0, st>>>(B);
    {   // low-texture top-up (src/pcd_generator.cpp:135-163): decided on the device, a few empty launches otherwise.
        // Its scratch re-uses level 1 of the pyramid, which nothing reads after the selection.
        CannyScratch cs;
        cs.blurred = reinterpret_cast<uint8_t*>(B.I[1]);
        cs.mag = reinterpret_cast<uint16_t*>(B.dx[1]);  // spans dx[1] + dy[1]
        cs.cls = reinterpret_cast<uint8_t*>(B.g2[1]);
        pcd_blur_kernel<<<(wh + T - 1) / T, T, 0, st>>>(B, cs);
        pcd_sobel_mag_kernel<<<(wh + T - 1) / T, T, 0, st>>>(B, cs);
        pcd_nms_kernel<<<(wh + T - 1) / T, T, 0, st>>>(B, cs);
        pcd_hysteresis_kernel<<<1, 1024, 0, st>>>(B, cs);
        pcd_topup_kernel<<<((w / 8) * (h / 8) + T - 1) / T, T, 0, st>>>(B, cs);
        pcd_topup_done_kernel<<<1, 1, 0, st>>>(B);
    }
    PackJob job = {raw_xyz, raw_feat, out_g, out_f, out_f4, 0, 0};
    CK(cudaMemcpyAsync(d_job, &job, sizeof(PackJob), cudaMemcpyHostToDevice, st));
    pcd_count_kernel<FLAG_POINT><<<nblk, 1024, 0, st>>>(B, P.d_blockcnt);
    pcd_points_kernel<<<nblk, 1024, 0, st>>>(B, P.d_blockcnt, camera_info(dataset_seq), feature_type, raw_xyz, raw_feat,
                                            ctx->max_points, &d_job->n);
    pack_sort_kernel<<<1, kPackThreads, ctx->pack_smem_max, st>>>(d_job, ctx->sort_points);
    CK(cudaGetLastError());
    ctx->launches += 23;
    CK(cudaMemcpyAsync(h_ctl, P.d_ctl, sizeof(SelCtl), cudaMemcpyDeviceToHost, st));
    return CVO_B200_OK;
}

static int check_image_args(cvo_b200_ctx* ctx, const unsigned char* img3, const unsigned short* depth, int width, int height,
                            int feature_type) {
    if (!img3 || !depth) return fail_arg(ctx, "null image pointer");
    if (width < 64 || height < 64 || width % 32 || height % 32 || (long long)width * height > (1 << 24))
        return fail_arg(ctx, "image size must be a multiple of 32 in both directions (the selector's block thresholds, "
                             "thirdparty/PixelSelector2.cpp:367, are only defined then)");
    if (feature_type != 0 && feature_type != 1) return fail_arg(ctx, "feature_type must be 0 (acvo) or 1 (cvo)");
    return CVO_B200_OK;
}

int cvo_b200_prefetch_frame_images(cvo_b200_ctx* ctx, const unsigned char* img3, const unsigned short* depth, int width,
                                   int height, int dataset_seq, int feature_type) {
    if (!ctx) return CVO_B200_ERR_ARG;
    int rc = check_image_args(ctx, img3, depth, width, height, feature_type);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    rc = flush_all_batches(ctx);
    if (rc) return rc;
    rc = ensure_image_pipe(ctx, width, height);
    if (rc) return rc;
    cvo_b200_ctx::ImagePipe& P = ctx->pipe;
    // the previous prefetched cloud must have left the `next` buffers (cvo_b200_push_prefetched_frame copies it on `stream`)
    CK(cudaStreamWaitEvent(ctx->copy_stream, P.ev_taken, 0));
    rc = enqueue_frontend(ctx, img3, depth, width, height, dataset_seq, feature_type, P.d_next_raw_xyz, P.d_next_raw_feat,
                          P.d_next_job, P.d_next_g, P.d_next_f, P.d_next_f4, ctx->copy_stream, P.h_ctl_next);
    if (rc) return rc;
    CK(cudaEventRecord(P.ev_next, ctx->copy_stream));
    P.next_pending = true;
    return CVO_B200_OK;
}

int cvo_b200_push_prefetched_frame(cvo_b200_ctx* ctx, int slot, int promote, int* num_points) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (slot < 0 || slot >= ctx->max_slots) return fail_arg(ctx, "slot out of range");
    cvo_b200_ctx::ImagePipe& P = ctx->pipe;
    if (!P.next_pending) return fail_arg(ctx, "no prefetched frame (cvo_b200_prefetch_frame_images)");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(P.ev_next));
    P.next_pending = false;
    const SelCtl& c = *P.h_ctl_next;
    if (num_points) *num_points = c.num_points;
    ctx->last_gen_canny = c.canny_used;
    if (c.status == PCD_STATUS_TOO_MANY_POINTS) return fail_arg(ctx, "the frame yields more points than max_points");
    if (c.num_points <= 0) {
        ctx->err = "empty cloud";
        return CVO_B200_ERR_EMPTY;
    }
    cvo_b200_ctx::Slot& s = ctx->slots[slot];
    int fixed_buf = s.fixed_buf, target;  // as in push_images
    if (!s.have_fixed) { fixed_buf = 0; target = 0; }
    else if (!s.bound) target = 1 - fixed_buf;
    else if (promote) { fixed_buf = 1 - fixed_buf; target = 1 - fixed_buf; }
    else target = 1 - fixed_buf;
    const size_t n = (size_t)c.num_points;
    CK(cudaMemcpyAsync(slot_g(ctx, slot, target), P.d_next_g, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(slot_f(ctx, slot, target), P.d_next_f, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(slot_f4(ctx, slot, target), P.d_next_f4, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaEventRecord(P.ev_taken, ctx->stream));
    s.fixed_buf = fixed_buf;
    s.n[target] = c.num_points;
    if (!s.have_fixed) s.have_fixed = true;
    else s.bound = true;
    return CVO_B200_OK;
}

int cvo_b200_push_frame_images(cvo_b200_ctx* ctx, int slot, const unsigned char* img3, const unsigned short* depth, int width,
                               int height, int dataset_seq, int feature_type, int* num_points) {
    return push_images(ctx, slot, img3, depth, width, height, dataset_seq, feature_type, num_points, true);
}

int cvo_b200_replace_moving_images(cvo_b200_ctx* ctx, int slot, const unsigned char* img3, const unsigned short* depth,
                                   int width, int height, int dataset_seq, int feature_type, int* num_points) {
    return push_images(ctx, slot, img3, depth, width, height, dataset_seq, feature_type, num_points, false);
}

int cvo_b200_last_generated_cloud(cvo_b200_ctx* ctx, float* xyz, float* feat, int capacity, int* n) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (!n) return fail_arg(ctx, "null pointer");
    CK(cudaSetDevice(ctx->device));
    *n = ctx->last_gen_n;
    const int m = ctx->last_gen_n < capacity ? ctx->last_gen_n : capacity;
    if (m > 0 && xyz) CK(cudaMemcpyAsync(xyz, ctx->d_raw_xyz, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (m > 0 && feat) CK(cudaMemcpyAsync(feat, ctx->d_raw_feat, sizeof(float) * 5 * m, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CVO_B200_OK;
}

int cvo_b200_last_frame_used_canny(const cvo_b200_ctx* ctx) { return ctx ? ctx->last_gen_canny : 0; }

int cvo_b200_reset_slot(cvo_b200_ctx* ctx, int slot) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (slot < 0 || slot >= ctx->max_slots) return fail_arg(ctx, "slot out of range");
    const int rc = flush_all_batches(ctx);
    if (rc) return rc;
    ctx->slots[slot] = cvo_b200_ctx::Slot();
    return CVO_B200_OK;
}

#ifdef CVO_PHASE_CLOCKS
int cvo_b200_phase_clocks(unsigned long long* out16, int reset) {
    cudaMemcpyFromSymbol(out16, g_phase_clocks, sizeof(unsigned long long) * 24);  // (the caller passes 24 slots)
    if (reset) {
        unsigned long long z[24] = {0};
        cudaMemcpyToSymbol(g_phase_clocks, z, sizeof(z));
    }
    return 0;
}
#endif

namespace {
// one warp per problem: the device's step_from_coeffs is warp-cooperative
__global__ void selftest_step_kernel(const double* bcde, int n, float min_step, float max_step, float* out) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const float s = step_from_coeffs(bcde[4 * i], bcde[4 * i + 1], bcde[4 * i + 2], bcde[4 * i + 3], min_step, max_step);
    if (threadIdx.x == 0) out[i] = s;
}
}  // namespace

namespace {
__global__ void selftest_exp_kernel(const float* wvs, int n, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    exp_sek3(wvs + 7 * i, wvs + 7 * i + 3, wvs[7 * i + 6], out + 12 * i, out + 12 * i + 9);
}
}  // namespace

int cvo_b200_selftest_exp_sek3(cvo_b200_ctx* ctx, const float* omega_v_dt, int n, float* dR_dT) {
    if (!ctx || !omega_v_dt || !dR_dT || n < 0) return CVO_B200_ERR_ARG;
    if (n == 0) return CVO_B200_OK;
    float *d_in = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_in, sizeof(float) * 7 * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(float) * 12 * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, omega_v_dt, sizeof(float) * 7 * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        selftest_exp_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(d_in, n, d_out);
        ctx->launches += 1;
        e = cudaMemcpyAsync(dR_dT, d_out, sizeof(float) * 12 * n, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) {
        ctx->err = std::string("selftest_exp_sek3: ") + cudaGetErrorString(e);
        return CVO_B200_ERR_CUDA;
    }
    return CVO_B200_OK;
}

int cvo_b200_selftest_step_size(cvo_b200_ctx* ctx, const double* bcde, int n, float min_step, float max_step, float* out) {
    if (!ctx || !bcde || !out || n < 0) return CVO_B200_ERR_ARG;
    if (n == 0) return CVO_B200_OK;
    double* d_in = nullptr;
    float* d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_in, sizeof(double) * 4 * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(float) * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, bcde, sizeof(double) * 4 * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        selftest_step_kernel<<<n, 32, 0, ctx->stream>>>(d_in, n, min_step, max_step, d_out);
        ctx->launches += 1;
        e = cudaMemcpyAsync(out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) {
        ctx->err = std::string("selftest_step_size: ") + cudaGetErrorString(e);
        return CVO_B200_ERR_CUDA;
    }
    return CVO_B200_OK;
}

int cvo_b200_selftest_rand_bytes(unsigned seed, int n, unsigned char* out) {
    if (n < 0 || !out) return CVO_B200_ERR_ARG;
    glibc_rand_bytes(seed, (size_t)n, out);
    return CVO_B200_OK;
}

int cvo_b200_align(cvo_b200_ctx* ctx, const int* slots, int n_pairs, const cvo_b200_params* p, float* RT_io,
                   float* ell_io, float* transform, float* prev_transform, int* iters, int* status) {
    return run_align(ctx, slots, n_pairs, p, RT_io, ell_io, transform, prev_transform, iters, status, nullptr, 0,
                     nullptr);
}

int cvo_b200_align_begin(cvo_b200_ctx* ctx, const int* slots, int n_pairs, const cvo_b200_params* p, const float* RT_in,
                         const float* ell_in) {
    return run_align_begin(ctx, slots, n_pairs, p, RT_in, ell_in, false, 0);
}

int cvo_b200_align_finish(cvo_b200_ctx* ctx, float* RT_out, float* ell_out, float* transform, float* prev_transform,
                          int* iters, int* status) {
    return run_align_finish(ctx, RT_out, ell_out, transform, prev_transform, iters, status, nullptr, nullptr);
}

int cvo_b200_align_trace(cvo_b200_ctx* ctx, int slot, const cvo_b200_params* p, float* RT_io, float* ell_io,
                         float* transform, float* prev_transform, int* iters, int* status,
                         cvo_b200_iter_rec* trace, int trace_cap, int* trace_len) {
    return run_align(ctx, &slot, 1, p, RT_io, ell_io, transform, prev_transform, iters, status, trace, trace_cap,
                     trace_len);
}

int cvo_b200_eval(cvo_b200_ctx* ctx, int slot, const float* R, const float* T, float ell, const cvo_b200_params* p,
                  cvo_b200_iter_rec* out) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (!R || !T || !p || !out) return fail_arg(ctx, "null pointer");
    cvo_b200_params q = *p;
    q.fixed_iters = 1;
    float RT[12];
    memcpy(RT, R, sizeof(float) * 9);
    memcpy(RT + 9, T, sizeof(float) * 3);
    float ell_io = ell;
    int len = 0;
    return run_align(ctx, &slot, 1, &q, RT, &ell_io, nullptr, nullptr, nullptr, nullptr, out, 1, &len);
}

int cvo_b200_inner_product(cvo_b200_ctx* ctx, int slot, float ell, const cvo_b200_params* p, float* value,
                           double* sum_a, long long* nnz) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (slot < 0 || slot >= ctx->max_slots || !ctx->slots[slot].bound) return fail_arg(ctx, "slot not bound");
    if (!p) return fail_arg(ctx, "null params");
    CK(cudaSetDevice(ctx->device));
    {
        const int rc = flush_all_batches(ctx);
        if (rc) return rc;
    }
    InnerArgs args;
    args.pair = make_pair_dev(ctx, slot);
    args.kp = make_kparams(p, true);
    args.ell = ell;
    args.out = ctx->d_inner;
    int G = ctx->force_G > 0 ? ctx->force_G : 8;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = launch_cluster_kernel(ctx, inner_product_kernel, args, G, 1, nullptr);
    if (rc != CVO_B200_OK) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_inner, ctx->d_inner, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    const double s = ctx->h_inner[0], c = ctx->h_inner[1];
    if (sum_a) *sum_a = s;
    if (nnz) *nnz = (long long)c;
    if (value) *value = (float)(s / c);  // src/adaptive_cvo.cpp:438
    return CVO_B200_OK;
}

int cvo_b200_sync(cvo_b200_ctx* ctx) {
    if (!ctx) return CVO_B200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    {
        const int rc = flush_all_batches(ctx);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CVO_B200_OK;
}

float cvo_b200_last_kernel_ms(const cvo_b200_ctx* ctx) { return ctx ? ctx->last_ms : 0.f; }
long long cvo_b200_kernel_launches(const cvo_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
int cvo_b200_last_cluster_size(const cvo_b200_ctx* ctx) { return ctx ? ctx->last_G : 0; }
int cvo_b200_last_num_clusters(const cvo_b200_ctx* ctx) { return ctx ? ctx->last_nclusters : 0; }
int cvo_b200_set_cluster_size(cvo_b200_ctx* ctx, int g) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (g < 0 || g > kMaxCluster) return fail_arg(ctx, "cluster size must be 0 (automatic) or 1..16");
    ctx->force_G = g;
    return CVO_B200_OK;
}
int cvo_b200_set_group_clusters(cvo_b200_ctx* ctx, int clusters_per_pair) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (clusters_per_pair < 0 || clusters_per_pair > kMaxGroupClusters) return fail_arg(ctx, "clusters per pair must be 0 (automatic) or 1..160");
    ctx->force_group = clusters_per_pair;
    return CVO_B200_OK;
}
int cvo_b200_last_group_clusters(const cvo_b200_ctx* ctx) { return ctx ? ctx->last_group : 0; }
long long cvo_b200_last_total_iterations(const cvo_b200_ctx* ctx) { return ctx ? ctx->last_total_iters : 0; }
long long cvo_b200_last_list_builds(const cvo_b200_ctx* ctx) { return ctx ? ctx->last_list_builds : 0; }
long long cvo_b200_last_list_refines(const cvo_b200_ctx* ctx) { return ctx ? ctx->last_list_refines : 0; }
int cvo_b200_last_list_fill(const cvo_b200_ctx* ctx, long long* entries, long long* slots) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (entries) *entries = ctx->last_xy_entries;
    if (slots) *slots = ctx->last_xy_slots;
    return CVO_B200_OK;
}
int cvo_b200_set_neighbor_lists(cvo_b200_ctx* ctx, int enable, float skin) {
    if (!ctx) return CVO_B200_ERR_ARG;
    if (!(skin >= 0.f && skin <= 1.f)) return fail_arg(ctx, "skin must be in [0, 1]");
    ctx->lists_enabled = enable != 0;
    ctx->list_skin = skin;
    return CVO_B200_OK;
}
int cvo_b200_num_sms(const cvo_b200_ctx* ctx) { return ctx ? ctx->num_sms : 0; }
int cvo_b200_neighbor_lists_active(const cvo_b200_ctx* ctx) {
    return ctx && ctx->lists_enabled && ctx->d_list_entries != nullptr ? 1 : 0;
}
long long cvo_b200_list_scratch_bytes(const cvo_b200_ctx* ctx) {
    if (!ctx || !ctx->d_list_entries) return 0;
    return (long long)ctx->list_ctas * kListAreas * (long long)ctx->list_cap * (long long)sizeof(uint2);
}

}  // extern "C"

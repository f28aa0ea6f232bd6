// cvo_kernels.cuh -- sm_100a kernels for the RKHS SE(3) registration inner loop.
//
// One persistent kernel runs the reference's whole align() loop
// (cpp/rkhs_registration/src/cvo.cpp:361-420, src/adaptive_cvo.cpp:490-555) on the device:
// a thread-block CLUSTER of G CTAs owns one frame pair at a time and iterates
//     transform_pcd -> se_kernel -> compute_flow -> compute_step_size -> Exp_SEK3 update
// without host round trips.  The sparse affinity matrix A (inc/cvo.hpp:92) is never stored:
// both all-pairs passes (flow, step coefficients) re-derive it tile by tile in shared memory
// with an on-the-fly ell-ball cutoff; tile pairs whose bounding boxes are farther apart than
// the ball radius are culled (clouds are Morton-sorted at upload), which is this design's
// replacement for the reference's nanoflann kd-tree (thirdparty/nanoflann.hpp).
// Cross-CTA reductions go through distributed shared memory + cluster barriers.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cvo_b200.h"

namespace cvo_b200 {

namespace cg = cooperative_groups;

// Build-time tuning knobs (scripts/build_variants.py sweeps them; the defaults are the measured best,
// profiles/r01_variant_sweep.txt).
#ifndef CVO_THREADS
#define CVO_THREADS 512
#endif
#ifndef CVO_BODY_ILP
#define CVO_BODY_ILP 2
#endif
#ifndef CVO_UNITS_PER_WARP
#define CVO_UNITS_PER_WARP 4
#endif
#ifndef CVO_GROUP
#define CVO_GROUP 2
#endif
#ifndef CVO_LIST_SETS
#define CVO_LIST_SETS 3
#endif
#ifndef CVO_SELF_SETS
#define CVO_SELF_SETS 3
#endif
#ifndef CVO_BUILD_SEGMENTS
#define CVO_BUILD_SEGMENTS 4
#endif
// L2 policies (measured, profiles/r02_variants.txt): the clouds are re-read every iteration while 148 lists stream through
// L2 between two uses, so cloud lines are loaded evict_last and list lines evict_first: cfg2 17.04k -> 17.50k pairs/s.
#ifndef CVO_L2_POLICIES
#define CVO_L2_POLICIES 1
#endif
#if CVO_L2_POLICIES
#define CVO_CLOUD_EVICT_LAST
#define CVO_LIST_EVICT_FIRST
#endif
constexpr int kThreads = CVO_THREADS;
constexpr int kWarps = kThreads / 32;
// The on-the-fly passes and the list builds keep per-warp queues and row tiles in shared memory: at most 16 warps
// take part in them (with more threads per CTA the others only help with staging and wait at the barriers).
constexpr int kWorkWarps = kWarps < 16 ? kWarps : 16;
constexpr int kTile = 32;
constexpr int kColChunk = 3072;                 // moving-cloud points resident in shared memory per pass
constexpr int kColTiles = kColChunk / kTile;
constexpr int kMaxCluster = 16;
constexpr int kMaxGroupClusters = 160;           // whole-GPU mode: clusters that can work on one pair
constexpr int kNumAcc = 16;
constexpr int kMaxUnits = 256;                  // work units (row tile x column segment) per scheduling round
static_assert(kThreads >= 32 + kMaxUnits, "build_list ranks one unit per thread beside the scanning warp");
constexpr int kUnitAcc = 9;                     // widest per-unit partial record (flow: omega, v, sum, nnz, dl)
constexpr int kQueueCap = 64 + kTile * kTile;   // leftovers (< 32 * CVO_BODY_ILP) + one full tile pair
static_assert(kColChunk <= (1 << 12), "queue entries pack row:5 | col:12 bits");
constexpr int kFlowOff = 4;  // sm.sum[0..3] = B,C,D,E ; sm.sum[kFlowOff + ACC_*] = flow totals
// padding points: far away from everything, but small enough that their squared norm stays finite
constexpr float kRowSentinel = 1.0e15f;
constexpr float kColSentinel = -1.0e15f;
// Prefilter slack: the mask phase tests the EXPANDED form |c|^2 - 2 c.x + |x|^2 < thr (3 FFMA per candidate) and
// only has to be a superset of the exact ball; its rounding error is bounded by ~20 * 2^-24 * (|c|^2 + |x|^2).
constexpr float kPrefilterSlack = 2.0e-6f;

enum PassKind { PASS_FLOW = 0, PASS_XX = 1, PASS_YY = 2, PASS_STEP = 3, PASS_INNER = 4,
                PASS_FLOW_CVO = 5 };  // FLOW without the length-scale gradient term (only acvo uses it)

// Neighbour candidate lists (the GPU counterpart of the reference's kd-tree, thirdparty/nanoflann.hpp): for one
// (rows, cols) cloud pair the (row, col) index pairs inside a ball of radius r_build = r * (1 + skin), kept in an
// L2-resident global scratch area of the CTA and re-used by every all-pairs pass until the pose has moved the
// column cloud by more than the skin (or ell changed the radius).  Only INDICES are stored: the strict ell-ball
// test, the colour gate and the kernel value are still evaluated on the fly in every pass, A never exists.
enum ListKind { LIST_XY = 0, LIST_XX = 1, LIST_YY = 2, LIST_KINDS = 3 };
constexpr int kListAreas = LIST_KINDS + 2;  // per CTA: the three lists, the build staging, the wide (x, y) list
constexpr int kMaxListRounds = 36;  // (row round, column chunk) combinations of one pass: 6 x 6 chunks of 3072 points
constexpr int kListTrip = 128;      // entries one warp handles per trip of a list pass; rounds are padded to it
#ifndef CVO_PREFETCH_TRIPS
#define CVO_PREFETCH_TRIPS 4
#endif
constexpr int kPrefetchTrips = CVO_PREFETCH_TRIPS;  // how many of its own trips ahead a warp prefetches the list into L2
// The quad passes neither clamp their look-ahead loads nor their prefetches to the end of a list (cvo_quads.cuh): the
// scratch allocation ends in this much slack, so they stay inside mapped memory whichever area comes last.
constexpr size_t kListSlackBytes = 64 * 1024;

// accumulator slots of the flow exchange
enum { ACC_W0 = 0, ACC_V0 = 3, ACC_SUMA = 6, ACC_NNZ = 7, ACC_DLXY = 8, ACC_NNZXX = 9, ACC_SXX = 10,
       ACC_NNZYY = 11, ACC_SYY = 12, ACC_FLOW_COUNT = 13 };
static_assert(ACC_FLOW_COUNT <= kNumAcc, "flow accumulators must fit the exchange buffers");

// One packed cloud in HBM: 36 B per point in three planes (see DESIGN.md "Data layout").
struct CloudDev {
    const float4* g;  // {x, y, z, bits of the original (pre-sort) index}
    const float4* f;  // {f0, f1, f2, f3}   -- moved to shared memory by TMA bulk copies, untouched
    const float* f4;  // {f4}               -- idem
    int n;
    int pad;
};

struct PairDev {
    CloudDev x;  // fixed  (cloud_x)
    CloudDev y;  // moving (cloud_y), original positions
};

struct PairState {
    float R[9];
    float T[3];
    float ell;
    float ell_max;
    int iters;
    int status;
    int n_run;
    int n_builds;  // neighbour-list (re)builds of the (x, y) list during this align()
    int n_refines;  // ... and how often one of the pair's lists was narrowed in place instead (refine_list)
    int xy_entries;  // summed over this CTA's (x, y) list builds (rank 0 of the cluster): candidates kept ...
    int xy_slots;    // ... and the slots of the quads that hold them (4 per quad, padding included)
    float tf[16];
    float prev_tf[16];
};

// cvo_b200_params + constants precomputed on the host in the reference's own arithmetic
struct KParams {
    int mode, ell_policy;
    int max_iter, fixed_iters;
    float ell_min;
    float s2;          // sigma*sigma
    float cs2;         // c_sigma*c_sigma
    float sp_thres;
    float log_ratio;   // logf(sp_thres/s2)            (src/cvo.cpp:102; log on a float is f32)
    float d2c_thres;   // colour gate                   (src/cvo.cpp:103 / src/adaptive_cvo.cpp:101)
    float inv2cl2;     // 1/(2 c_ell^2)
    float c2;          // log2(e)/(2 c_ell^2)
    float s2cs2;       // sigma^2 c_sigma^2
    float c_ell;
    float sp_band;     // half-width around sp_thres inside which the kernel value is re-decided exactly
    float t_lim;       // log2(s2 c_sigma^2 / sp_thres), rounded up: a > sp_thres  <=>  d2 c1 + t_c < t_lim
    float inv_c, inv_d;
    float min_step, max_step, eps, eps_2;
    double dl_step;
};

struct IterConsts {
    float tf[12];  // transform: rows of R^T, then -R^T T   (src/cvo.cpp:83-87)
    float d2_thres, d2c_thres, inv2l2, inv_ell3;
    float c1;  // log2(e)/(2 ell^2)
    float ell;
    float omega[3], v[3];
    float temp_coef, m2t, p2t;
};

// Private scratch of one warp: the row tile it currently owns.
struct WarpScratch {
    float4 rowG[kTile];          // {x, y, z, f4}
    float4 rowF[kTile];          // {f0, f1, f2, f3}
    int rowOrig[kTile];          // original row indices (PASS_YY only: quirk Q1 is defined on them)
};

// What sits beside the column geometry depends on the pass.  On-the-fly passes and list builds need the column
// features, the warps' survivor queues and row tiles, and the per-unit partial sums; a pass over a neighbour list
// needs this CTA's rows (geometry only) and, for the STEP pass, the per-column step-size terms.
// Row-sorted compaction of a freshly built (x, y) list (build_list<0>, "quads"): per row of the round how many entries
// it has and where its first quad sits inside its row tile; per row tile the first quad.
struct QuadBuild {
    int rowQ[kColChunk];
    int tileQ[kColTiles + 1];
};
struct FeatStage {
    float4 colF[kColChunk];             // {f0, f1, f2, f3}
    float colF4[kColChunk];             // f4
    union {
        uint32_t queue[kWorkWarps][kQueueCap];  // per warp: in-ball (row, col) pairs waiting for the survivor body
        QuadBuild qb;                           // after the evaluation of a round: the compaction's counters
    };
};
static_assert(sizeof(QuadBuild) <= sizeof(uint32_t) * kWorkWarps * kQueueCap, "the compaction counters live in the queues' memory");
struct StepStage {
    // (list passes: four planes of kColChunk floats each, see plane_ld; nrm and pdt pre-scaled, see step_col)
    float4 colZ1[kColChunk];  // {xi z + v, |xi z + v|^2}                       (src/cvo.cpp:226-228,235)
    float4 colZ2[kColChunk];  // {xi^2 z + xi v, -(xi z + v).(xi^2 z + xi v)}   (src/cvo.cpp:229-230,236)
};
struct BuildUnits {  // neighbour-list build, per unit of the round:
    int off[kMaxUnits];  // where its entries sit in the staging area
    int act[kMaxUnits];  // how many it has
    int pos[kMaxUnits];  // their position in the round's flat list
    int rowCnt[kColChunk];  // (x, y) list: candidates kept per row of the round, counted by the warp that owns the row's tile
};
static_assert(sizeof(BuildUnits) <= sizeof(double) * kMaxUnits * kUnitAcc, "BuildUnits shares the memory of the on-the-fly unit slots");
struct OnTheFlyStage {
    FeatStage fs;
    WarpScratch ws[kWorkWarps];
    union {
        double unitPart[kMaxUnits][kUnitAcc];  // on-the-fly pass: one fixed slot per work unit => scheduling-independent sums
        BuildUnits bu;                         // list build
    };
};
struct ListStage {
    float4 rowG[kColChunk];  // planes x[], y[], z[] (kColChunk floats each) of the round's (transformed) rows
    StepStage ss;
    double warpTot[kWarps][kNumAcc];  // one total per warp, summed in warp order
};

struct ListState {
    float tf[12];     // transform the (x, y) list was built at
    float r0;         // ell-ball radius at build time
    float slack;      // how far the cloud may move / the ball may grow before the list misses a neighbour
    float s_build;    // slack + rounding margin: what the build adds to a pair's own radius
    float thr_build;  // (r0 + s_build)^2: the build prefilter's ball
    float inv_c1;     // 2 l^2 / log2(e) at build time: colour exponent -> squared radius
    int valid;        // 1: usable, 0: must be built, -1: overflowed its scratch for this pair (on-the-fly passes)
    int need;         // (re)build before this iteration's passes
};

// The WIDE (x, y) candidate list: what an all-pairs sweep found within r_e + s + W of the sweep's pose, kept per row tile in
// the staged format (column, row within the tile, t_c).  While it covers the current pose and length-scale, a rebuild of
// the quads is a FILTER of it (one streaming pass, ~45 instructions per 32 entries) instead of another all-pairs sweep.
// Coverage: a pair can only be wanted by a new list (|x_i - T1 y_j| < r_e1 + s1) if it is in the wide one
// (|x_i - Tw y_j| < r_e_w + s_w + W), i.e. as long as  max(0, r1 - r_w) + disp(Tw -> T1) + s1 <= s_w + W  (r_e scales
// with the length-scale and never exceeds r: the pair-specific radii only make the left side smaller).
struct WideState {
    float tf[12];      // transform of the sweep
    float r0;          // ell-ball radius of the sweep
    float slack;       // s_w + W, rounded down: what the coverage test may assume
    float s_build;     // s_w + W + rounding margin, rounded up: what the sweep adds to a pair's own radius
    float thr_build;   // (r0 + s_build)^2: the sweep's prefilter ball
    int valid;         // 1: covers what `slack` says; 0: none (never built, overflowed its area, other pair)
    int make;          // this iteration's sweep also writes the wide list
};

// Identity of the points a shared-memory stage holds: (cloud, first point, count, pose).  `serial` is the iteration
// whose transform was applied, -1 for untransformed points, -2 for "nothing usable".
struct StageTag {
    const float4* g;
    int first, n, serial;
};
__device__ __forceinline__ bool tag_is(const StageTag& t, const float4* g, int first, int n, int serial) {
    return t.g == g && t.first == first && t.n == n && t.serial == serial;
}

// The passes over a neighbour list keep their stages as PLANES (structure of arrays: x[], y[], z[], w[] of kColChunk
// floats each, in the memory of the float4 arrays named below): the 32 entries a warp handles at a time address a few
// consecutive rows and columns of one 32-column tile, so 4-byte gathers from a plane hit 32 different banks (equal
// indices broadcast), while 16-byte gathers of {x, y, z, w} records replay on every pair of indices that agree mod 8
// and move the unused w lane.  An entry holds the BYTE offsets of its row and column inside a plane.
constexpr uint32_t kPlaneBytes = (uint32_t)kColChunk * 4u;
template <int PLANE>
__device__ __forceinline__ float plane_ld(const void* base, uint32_t byte_off) {
    return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + PLANE * kPlaneBytes + byte_off);
}
__device__ __forceinline__ float* plane_of(void* base, int plane) { return reinterpret_cast<float*>(base) + plane * kColChunk; }

struct ListRef {
    uint2* entries;  // the list area: quads ((x, y) list, cvo_quads.cuh) or flat 8-byte entries (self lists)
    uint2* staging;  // build scratch of the CTA (shared by its three lists): per-unit regions before compaction
    uint2* wide;     // the CTA's WIDE (x, y) candidate list (see WideState): [kWideTable (offset, count) records][entries]
    unsigned cap;
};
constexpr int kWideTable = kMaxListRounds * kColTiles;  // one record per (round, row tile) of the CTA

struct Smem {
    float4 colG[kColChunk];   // on-the-fly passes / list builds: {x, y, z, |c|^2} records of the staged (transformed) columns;
                              // list passes: planes x[], y[], z[], w[] of kColChunk floats each (see plane_ld), w = the
                              // (scaled) step-size term of src/cvo.cpp:237 in the STEP pass
    union {
        OnTheFlyStage of;
        ListStage ls;
    } u;
    float colBox[kColTiles][8];  // [0..2] lo, [3..5] hi, [6] max |c|^2
    double blockTot[kNumAcc];
    double flowTot[kNumAcc];  // this CTA's flow-exchange vector (ACC_* layout)
    uint2 lround[LIST_KINDS][kMaxListRounds];  // (offset, entries) of every round of a list; entries % kListTrip == 0
    int lst_base;
    int refineCnt[kWarps], refinePos[kWarps];  // refine_list: entries every warp kept / where they go
    StageTag colTag, rowTag;  // what the column / row stages of the list passes currently hold
    int serial;               // running iteration number of this CTA: identifies "transformed with this iteration's pose"
    float wred[kWarps][6];    // per-warp partial bounding boxes (pair start)
    float ybox[6];            // bounding box of the moving cloud, original coordinates
    ListState lst[LIST_KINDS];
    WideState wide;
    int wide_ovf;  // a warp's share of the wide area overflowed during this sweep
    int lst_used, lst_ovf;
    int next_unit;
    int next_pair;
    int done;
    int k;
    double xchg[2][kMaxCluster][kNumAcc];
    double sum[kFlowOff + kNumAcc];
    IterConsts ic;
    PairState st;
    unsigned long long tma_bar;  // mbarrier the TMA bulk copies of a column chunk complete on
#ifdef CVO_PHASE_CLOCKS
    long long phase_t0;
#endif
};

#ifdef CVO_PRINT_SMEM
template <size_t N> struct SmemSizeIs;
SmemSizeIs<sizeof(Smem)> smem_size_probe;
#endif
static_assert(sizeof(Smem) <= 227 * 1024, "Smem must fit the 227 KB per-CTA shared memory of sm_100");

struct AlignArgs {
    const PairDev* pairs;
    PairState* states;
    int n_pairs;
    int* counter;
    cvo_b200_iter_rec* trace;  // records of pair 0 only (align_trace / eval), or nullptr
    int trace_cap;
    KParams kp;
    // neighbour-list scratch: [gridDim.x][LIST_KINDS + 1] areas of list_cap entries (three lists + build staging)
    uint2* list_entries;     // nullptr: lists disabled, every pass is on the fly
    unsigned list_cap;
    float list_skin;
    float list_wide;      // W / r: extra slack of the wide list, 0 = no wide list (every rebuild is a sweep)
    float list_skin_min;  // absolute floor of the skin [m]: at small length-scales the lists are short and rebuilds dominate
    float list_shrink;  // rebuild a list when ell has shrunk the ball below this fraction of its build radius
    float list_refine_min;  // ... by filtering the old list if it has at least this fraction of a fresh skin to spare
    // Whole-GPU mode for a few large pairs: ALL `group_clusters` clusters of the launch work on the same pair (the pairs
    // are taken one after the other); the cluster totals meet in global memory, see group_allreduce.  <= 1: off.
    int group_clusters;
    double* group_xchg;     // [2][kMaxGroupClusters][kNumAcc]
    unsigned* group_count;  // arrivals of the clusters' rank-0 CTAs, zeroed before the launch
};

struct InnerArgs {
    PairDev pair;
    KParams kp;
    float ell;
    double* out;  // [0] = sum_a, [1] = nnz
};

struct PackJob {
    const float* xyz;   // n x 3
    const float* feat;  // n x 5
    float4* out_g;
    float4* out_f;
    float* out_f4;
    int n;
    int pad;
};

// --------------------------------------------------------------------------------------------
// small helpers
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sums NV per-lane values over the warp in a fixed order.  The first 8 go through a transposing butterfly (each
// xor step halves the number of values a lane still carries: 4 + 2 + 1 + 1 + 1 = 9 shuffles instead of 40); value
// i (i < 8) ends up in lanes with ((lane >> 2) & 7) == i, any further value in every lane.  Lane 0 gets value 0;
// `out_lane(i)` tells which lane holds value i.
template <int NV>
__device__ __forceinline__ void warp_sum_multi(double (&v)[NV], int lane) {
    if (NV >= 8) {
        double h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // xor 16: lanes with bit 4 clear keep 0..3, the others 4..7
            const bool up = (lane & 16) != 0;
            const double keep = up ? v[4 + i] : v[i], send = up ? v[i] : v[4 + i];
            h[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
        double q[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {  // xor 8
            const bool up = (lane & 8) != 0;
            const double keep = up ? h[2 + i] : h[i], send = up ? h[i] : h[2 + i];
            q[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        const bool up = (lane & 4) != 0;  // xor 4
        double r = (up ? q[1] : q[0]) + __shfl_xor_sync(0xffffffffu, up ? q[0] : q[1], 4);
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 1);
        v[0] = r;  // this lane's value index is ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)
#pragma unroll
        for (int i = 8; i < NV; ++i) v[i] = warp_sum(v[i]);
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    }
}
__device__ __forceinline__ int multi_value_index(int lane) {
    return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// all three finite <=> the sum of the magnitudes is finite (NaN and Inf both fail the comparison)
__device__ __forceinline__ bool finite3(float x, float y, float z) {
    return (fabsf(x) + fabsf(y)) + fabsf(z) < __int_as_float(0x7f800000);
}

// sqrt to 2 ulp in one MUFU (the list builds use it inside bounds that carry their own safety factor)
__device__ __forceinline__ float sqrtf_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float exp2f_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// y = R^T p - R^T T with the fma chain documented in DESIGN.md (membership-critical)
__device__ __forceinline__ void apply_tf(const float* tf, float& x, float& y, float& z) {
    const float px = x, py = y, pz = z;
    x = __fadd_rn(__fmaf_rn(tf[2], pz, __fmaf_rn(tf[1], py, __fmul_rn(tf[0], px))), tf[9]);
    y = __fadd_rn(__fmaf_rn(tf[5], pz, __fmaf_rn(tf[4], py, __fmul_rn(tf[3], px))), tf[10]);
    z = __fadd_rn(__fmaf_rn(tf[8], pz, __fmaf_rn(tf[7], py, __fmul_rn(tf[6], px))), tf[11]);
}

// squared distance exactly as nanoflann's L2 tail loop under fp-contract (thirdparty/nanoflann.hpp:402-406)
__device__ __forceinline__ float dist2(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ void mat3_mul(const float* a, const float* b, float* r) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r[i * 3 + j] = __fadd_rn(__fadd_rn(__fmul_rn(a[i * 3], b[j]), __fmul_rn(a[i * 3 + 1], b[3 + j])),
                                     __fmul_rn(a[i * 3 + 2], b[6 + j]));
}
__device__ __forceinline__ void mat3_vec(const float* a, const float* v, float* r) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
        r[i] = __fadd_rn(__fadd_rn(__fmul_rn(a[i * 3], v[0]), __fmul_rn(a[i * 3 + 1], v[1])), __fmul_rn(a[i * 3 + 2], v[2]));
}
__device__ __forceinline__ void skew3(const float* w, float* M) {  // src/LieGroup.cpp:20-27
    M[0] = 0.f;   M[1] = -w[2]; M[2] = w[1];
    M[3] = w[2];  M[4] = 0.f;   M[5] = -w[0];
    M[6] = -w[1]; M[7] = w[0];  M[8] = 0.f;
}
__device__ __forceinline__ float dot3f(const float* a, const float* b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
}

#ifdef CVO_PHASE_CLOCKS  // tuning aid (scripts/build_variants.py clk:CVO_PHASE_CLOCKS, scripts/gpu_phase_clocks.py): cycles
__device__ unsigned long long g_phase_clocks[24];  // thread 0 of every CTA spends per phase, summed over the CTAs
#define CVO_PHASE(i)                                                                        \
    if (threadIdx.x == 0) {                                                                 \
        const long long now = clock64();                                                    \
        atomicAdd(&g_phase_clocks[i], (unsigned long long)(now - sm.phase_t0));             \
        sm.phase_t0 = now;                                                                  \
    }
#else
#define CVO_PHASE(i)
#endif

// --------------------------------------------------------------------------------------------
// scalar epilogue pieces (one thread per CTA; every CTA of a cluster computes the same values)
// --------------------------------------------------------------------------------------------

// update_tf (src/cvo.cpp:83-87) + thresholds of se_kernel (src/cvo.cpp:102-103)
__device__ void prepare_iter(Smem& sm, const KParams& kp, float d2c_thres) {
    IterConsts& ic = sm.ic;
    const float* R = sm.st.R;
    const float* T = sm.st.T;
    float nRt[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            ic.tf[i * 3 + j] = R[j * 3 + i];
            nRt[i * 3 + j] = -R[j * 3 + i];
        }
    mat3_vec(nRt, T, &ic.tf[9]);
    ic.d2c_thres = d2c_thres;
    // Everything below depends on the length-scale only: f64 divisions on one thread, recomputed when ell has changed
    // (never in the fixed-ell benchmark schedule, three times in the stock cvo schedule).  ic.ell < 0: a new pair.
    if (ic.ell != sm.st.ell) {
        const double l = (double)sm.st.ell;
        const double inv = 1.0 / (2.0 * l * l);
        ic.d2_thres = (float)(-2.0 * l * l * (double)kp.log_ratio);
        ic.inv2l2 = (float)inv;
        ic.c1 = (float)(1.4426950408889634 / (2.0 * l * l));
        ic.ell = sm.st.ell;
        const float ell3 = __fmul_rn(__fmul_rn(sm.st.ell, sm.st.ell), sm.st.ell);  // src/adaptive_cvo.cpp:171
        ic.inv_ell3 = 1.0f / ell3;
        ic.temp_coef = (float)inv;  // src/cvo.cpp:241
        ic.m2t = (float)(-2.0 * (double)ic.temp_coef);
        ic.p2t = (float)(2.0 * (double)ic.temp_coef);
    }
}

// tail of compute_flow (src/cvo.cpp:208-209) + the per-iteration constants of compute_step_size (:215-241)
__device__ void finalize_flow(Smem& sm) {
    IterConsts& ic = sm.ic;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        ic.omega[t] = (float)sm.sum[kFlowOff + ACC_W0 + t];
        ic.v[t] = (float)sm.sum[kFlowOff + ACC_V0 + t];
    }
    // (temp_coef = 1 / (2 l^2), src/cvo.cpp:241, and its multiples: prepare_iter, with the other functions of ell)
}

// poly_solver + root selection (src/cvo.cpp:53-69,291-307): smallest positive real root of
// 4E t^3 + 3D t^2 + 2C t + B: closed form on the f32-normalised coefficients, polished to the f64 root by Newton.
// Called by ALL lanes of one warp with the same arguments: lane i % 3 evaluates and polishes root i, then the
// smallest positive root is taken across the lanes.  Every lane returns the same step.
__device__ float step_from_coeffs(double B, double C, double D, double E, float min_step, float max_step) {
    const int which = (threadIdx.x & 31) % 3;
    const float p0 = (float)(4.0 * (double)(float)E);
    const float p1 = (float)(3.0 * (double)(float)D);
    const float p2 = (float)(2.0 * (double)(float)C);
    const float p3 = (float)B;
    const float a2f = p1 / p0, a1f = p2 / p0, a0f = p3 / p0;
    const float kNone = 3.402823466e+38f;
    float best = kNone;
    if (isfinite(a2f) && isfinite(a1f) && isfinite(a0f)) {
        const double a2 = a2f, a1 = a1f, a0 = a0f;
        const double q = (3.0 * a1 - a2 * a2) * (1.0 / 9.0);
        const double r = (9.0 * a2 * a1 - 27.0 * a0 - 2.0 * a2 * a2 * a2) * (1.0 / 54.0);
        const double disc = q * q * q + r * r;
        const double shift = a2 * (1.0 / 3.0);
        // Closed form in f32 as the starting point (the f64 cbrt / acos / cos were the longest dependent chain of the
        // serial section), Newton in f64 to convergence: the root is the f64 root either way.  Which formula applies is
        // decided by the f64 discriminant.
        const float qf = (float)q, rf = (float)r, shiftf = (float)shift;
        float g = 0.f;
        bool have = false;
        if (disc > 0.0) {  // one real root
            if (which == 0) {
                const float sd = sqrtf((float)disc);
                g = cbrtf(rf + sd) + cbrtf(rf - sd) - shiftf;
                have = true;
            }
        } else if (disc == 0.0) {  // a double root
            if (which < 2) {
                const float sr = cbrtf(rf);
                g = (which == 0 ? 2.f * sr : -sr) - shiftf;
                have = true;
            }
        } else {  // three real roots
            float cth = rf * rsqrtf(-qf * qf * qf);
            cth = fminf(1.f, fmaxf(-1.f, cth));
            const float th = acosf(cth);
            g = 2.f * sqrtf(-qf) * cosf((th + (float)which * 6.2831853f) * (1.f / 3.f)) - shiftf;
            have = true;
        }
        double x = (double)g;
        if (have && !isfinite(g)) {  // coefficients outside the f32 range: the same formulas in f64
            if (disc > 0.0) {
                const double sd = sqrt(disc);
                x = cbrt(r + sd) + cbrt(r - sd) - shift;
            } else if (disc == 0.0) {
                const double sr = cbrt(r);
                x = (which == 0 ? 2.0 * sr : -sr) - shift;
            } else {
                double cth = r / sqrt(-q * q * q);
                cth = fmin(1.0, fmax(-1.0, cth));
                x = 2.0 * sqrt(-q) * cos((acos(cth) + (double)which * 6.283185307179586476925286766559) * (1.0 / 3.0)) - shift;
            }
        }
        if (have) {
            for (int it = 0; it < 8; ++it) {  // Newton: two or three steps from an f32-accurate start
                const double f = ((x + a2) * x + a1) * x + a0;
                const double fp = (3.0 * x + 2.0 * a2) * x + a1;
                if (fp == 0.0 || !isfinite(f)) break;
                // f / fp through an f32 reciprocal refined once in f64 (relative error ~1e-14; Newton corrects itself): the
                // IEEE f64 division is the longest dependent chain of this loop
                double inv = (double)__frcp_rn((float)fp);
                inv = inv * (2.0 - fp * inv);
                const double dx = isfinite(inv) ? f * inv : f / fp;
                const double xn = x - dx;
                if (!isfinite(xn)) break;
                x = xn;
                if (fabs(dx) <= 1.0e-13 * fabs(x)) break;
            }
            const float xr = (float)x;
            if (xr > 0.f) best = xr;
        }
    }
    best = warp_min(best);
    float step = (best == kNone) ? min_step : best;
    return step > max_step ? max_step : step;
}

// Exp_SEK3 with K = 1 (src/LieGroup.cpp:159-186), including the small-angle quirk (Jl = I)
__device__ void exp_sek3(const float* w, const float* v, float dt, float* dR, float* dT) {
    const float theta = sqrtf(dot3f(w, w));
    float Jl[9];
    if (theta < 1e-6f) {
#pragma unroll
        for (int i = 0; i < 9; ++i) dR[i] = Jl[i] = (i % 4 == 0) ? 1.f : 0.f;
    } else {
        float A[9], A2[9];
        skew3(w, A);
        mat3_mul(A, A, A2);
        const float theta2 = theta * theta;
        const float st = sinf(dt * theta), ct = cosf(dt * theta);
        const float omc = (1.f - ct) / theta2;
        const float sa = st / theta;
        const float sj = (dt * theta - st) / (theta2 * theta);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float I = (i % 4 == 0) ? 1.f : 0.f;
            dR[i] = __fadd_rn(__fadd_rn(I, __fmul_rn(sa, A[i])), __fmul_rn(omc, A2[i]));
            Jl[i] = __fadd_rn(__fadd_rn(__fmul_rn(dt, I), __fmul_rn(omc, A[i])), __fmul_rn(sj, A2[i]));
        }
    }
    mat3_vec(Jl, v, dT);
}

// Body of align() after compute_step_size (src/cvo.cpp:379-410, src/adaptive_cvo.cpp:508-545)
// Called by all lanes of warp 0; lane 0 applies the update.
__device__ void update_state(Smem& sm, const KParams& kp, int k, cvo_b200_iter_rec* rec) {
    IterConsts& ic = sm.ic;
    PairState& st = sm.st;
    const double B = sm.sum[0], C = sm.sum[1], D = sm.sum[2], E = sm.sum[3];
    const float step = step_from_coeffs(B, C, D, E, kp.min_step, kp.max_step);
    CVO_PHASE(14)
    if ((threadIdx.x & 31) != 0) return;
    const bool stops = !(kp.fixed_iters > 0);
    const float ell_used = st.ell;
    bool stop = false;
    int status = CVO_B200_STATUS_MAX_ITER;
    // the twist and the pose in registers: one round of shared-memory loads instead of one per use
    float om[3], vv[3], Rc[9], Tc[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) { om[t] = ic.omega[t]; vv[t] = ic.v[t]; Tc[t] = st.T[t]; }
#pragma unroll
    for (int t = 0; t < 9; ++t) Rc[t] = st.R[t];
    const float w2 = dot3f(om, om), v2 = dot3f(vv, vv);
    if (!(isfinite(w2) && isfinite(v2))) {
        stop = true;
        status = CVO_B200_STATUS_NAN;
    }
    if (!stop && stops) {
        bool small;
        if (kp.mode == CVO_B200_MODE_ACVO) {  // src/adaptive_cvo.cpp:509, norms in f64
            const double dw = sqrt((double)om[0] * om[0] + (double)om[1] * om[1] + (double)om[2] * om[2]);
            const double dv = sqrt((double)vv[0] * vv[0] + (double)vv[1] * vv[1] + (double)vv[2] * vv[2]);
            small = dw < (double)kp.eps && dv < (double)kp.eps;
        } else {  // src/cvo.cpp:380
            small = sqrtf(w2) < kp.eps && sqrtf(v2) < kp.eps;
        }
        if (small) {
            stop = true;
            status = CVO_B200_STATUS_CONVERGED_TWIST;
        }
    }
    if (!stop) {
        float dR[9], dT[3], RdT[3], Rn[9];
        exp_sek3(om, vv, step, dR, dT);  // src/cvo.cpp:391
        mat3_vec(Rc, dT, RdT);
#pragma unroll
        for (int t = 0; t < 3; ++t) st.T[t] = __fadd_rn(RdT[t], Tc[t]);  // :398
        mat3_mul(Rc, dR, Rn);                                           // :399
#pragma unroll
        for (int t = 0; t < 9; ++t) st.R[t] = Rn[t];
        if (stops) {
            // dist_se3 (src/cvo.cpp:71-81): ||logm(Exp(step*[w^ v;0 0]))||_F in closed form
            const float theta = sqrtf(w2);
            const float dist = (theta < 1e-6f) ? sqrtf(v2) : step * sqrtf(2.f * w2 + v2);
            if (dist < kp.eps_2) {  // :402
                stop = true;
                status = CVO_B200_STATUS_CONVERGED_UPDATE;
            }
        }
    }
    double dl = 0.0;
    if (kp.mode == CVO_B200_MODE_ACVO) {  // src/adaptive_cvo.cpp:271
        const double num = -2.0 * sm.sum[kFlowOff + ACC_DLXY] + sm.sum[kFlowOff + ACC_SXX] + sm.sum[kFlowOff + ACC_SYY];
        const long long den = (long long)sm.sum[kFlowOff + ACC_NNZXX] + (long long)sm.sum[kFlowOff + ACC_NNZYY] -
                              2 * (long long)sm.sum[kFlowOff + ACC_NNZ];
        dl = num / (double)den;
    }
    if (!stop) {
        if (kp.ell_policy == CVO_B200_ELL_SCHEDULE) {  // src/cvo.cpp:408-410
            st.ell = (k > 2) ? 0.10f : st.ell;
            st.ell = (k > 9) ? 0.06f : st.ell;
            st.ell = (k > 19) ? 0.03f : st.ell;
        } else if (kp.ell_policy == CVO_B200_ELL_ADAPTIVE) {  // src/adaptive_cvo.cpp:538-545
            st.ell = (float)((double)st.ell + kp.dl_step * dl);
            if (st.ell >= st.ell_max) {
                st.ell = (float)((double)st.ell_max * 0.7);
                st.ell_max = (float)((double)st.ell_max * 0.7);
            }
            st.ell = (st.ell < kp.ell_min) ? kp.ell_min : st.ell;
        }
    }
    st.n_run = k + 1;
    if (stop) {
        st.iters = k;
        st.status = status;
        sm.done = 1;
    }
    if (rec) {
        rec->ell = ell_used;
        rec->step = step;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            rec->omega[t] = ic.omega[t];
            rec->v[t] = ic.v[t];
            rec->T[t] = st.T[t];
        }
        rec->B = B; rec->C = C; rec->D = D; rec->E = E;
        rec->sum_a = sm.sum[kFlowOff + ACC_SUMA];
        rec->dl = dl;
        rec->nnz = (long long)sm.sum[kFlowOff + ACC_NNZ];
        rec->nnz_xx = (long long)sm.sum[kFlowOff + ACC_NNZXX];
        rec->nnz_yy = (long long)sm.sum[kFlowOff + ACC_NNZYY];
#pragma unroll
        for (int t = 0; t < 9; ++t) rec->R[t] = st.R[t];
    }
}

__device__ void write_tf44(const float* tf12, float* out) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) out[i * 4 + j] = tf12[i * 3 + j];
        out[i * 4 + 3] = tf12[9 + i];
    }
    out[12] = out[13] = out[14] = 0.f;
    out[15] = 1.f;
}

// --------------------------------------------------------------------------------------------
// tile staging
// --------------------------------------------------------------------------------------------

// TMA (cp.async.bulk) + mbarrier plumbing: the feature planes of a column chunk go HBM -> shared memory without
// passing through registers; completion is signalled on an mbarrier by transaction bytes.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// Step-size terms of one (transformed) moving point z (src/cvo.cpp:226-237) through z_{k+1} = omega x z_k
// (= Omega^k y + Omega^(k-1) v): z1 = xi z + v, z2 = xi^2 z + xi v, |z1|^2, -z1.z2, |z2|^2 + 2 z1.z3.
struct StepCol {
    float z1x, z1y, z1z, nrm;  // nrm, pdt, ecn: scaled by -t, 2t, -t (t = 1/(2 l^2)), see step_col
    float z2x, z2y, z2z, pdt;
    float ecn;
};
template <class IC>
__device__ __forceinline__ StepCol step_col(const IC& ic, float yx, float yy, float yz) {
    const float w0 = ic.omega[0], w1 = ic.omega[1], w2 = ic.omega[2];
    StepCol c;
    c.z1x = (w1 * yz - w2 * yy) + ic.v[0];
    c.z1y = (w2 * yx - w0 * yz) + ic.v[1];
    c.z1z = (w0 * yy - w1 * yx) + ic.v[2];
    c.z2x = w1 * c.z1z - w2 * c.z1y; c.z2y = w2 * c.z1x - w0 * c.z1z; c.z2z = w0 * c.z1y - w1 * c.z1x;
    const float z3x = w1 * c.z2z - w2 * c.z2y, z3y = w2 * c.z2x - w0 * c.z2z, z3z = w0 * c.z2y - w1 * c.z2x;
    const float nrm = (c.z1x * c.z1x + c.z1y * c.z1y) + c.z1z * c.z1z;                                   // normxiz2, :235
    const float pdt = -((c.z1x * c.z2x + c.z1y * c.z2y) + c.z1z * c.z2z);                                // xiz_dot_xi2z, :236
    const float ecn = ((c.z2x * c.z2x + c.z2y * c.z2y) + c.z2z * c.z2z) + 2.f * ((c.z1x * z3x + c.z1y * z3y) + c.z1z * z3z);  // :237
    // stored with the coefficients gamma / delta / epsilon multiply them by (src/cvo.cpp:264-270): one FMA per term per entry
    c.nrm = -ic.temp_coef * nrm;
    c.pdt = ic.p2t * pdt;
    c.ecn = -ic.temp_coef * ecn;
    return c;
}

// Stages `ntiles` 32-point tiles starting at point `base` of a packed cloud into shared memory.
//  STAGE_FULL (on-the-fly passes, list builds):
//   * feature planes (20 B / point): two TMA bulk copies issued by one thread, completing on sm.tma_bar;
//   * geometry plane (16 B / point): float4 loads by all threads, the rigid transform applied on the way (this IS
//     transform_pcd, src/cvo.cpp:310-315: the transformed cloud never exists in HBM), |c|^2 appended for the
//     prefilter, and one bounding box per tile reduced with warp shuffles.
//  STAGE_GEOM (FLOW / XX / YY pass over a list): the transformed geometry only.
//  STAGE_STEP (STEP pass over a list): the transformed geometry plus the per-column step-size terms.
// The caller has synchronised the CTA (nobody still reads the previous chunk) and synchronises again afterwards.
enum StageMode { STAGE_FULL = 0, STAGE_GEOM = 1, STAGE_STEP = 2 };
template <int MODE>
__device__ __forceinline__ void stage_tiles(Smem& sm, const CloudDev& c, int base, int ntiles, bool tf,
                                            float sentinel, uint32_t& tma_phase) {
    const int lane = threadIdx.x & 31;
    const float inf = __int_as_float(0x7f800000);
    if (MODE == STAGE_FULL) {
        // the feature stage shares its shared memory with the list passes' row / step stages, which are written
        // with ordinary stores: order those before the bulk copies of the async proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    if (MODE == STAGE_FULL && threadIdx.x == 0) {
        const uint32_t bytes_f = (uint32_t)(ntiles * kTile) * 16u, bytes_f4 = (uint32_t)(ntiles * kTile) * 4u;
        mbar_expect_tx(&sm.tma_bar, bytes_f + bytes_f4);
        tma_bulk_g2s(sm.u.of.fs.colF, c.f + base, bytes_f, &sm.tma_bar);
        tma_bulk_g2s(sm.u.of.fs.colF4, c.f4 + base, bytes_f4, &sm.tma_bar);
    }
    const float* tf12 = sm.ic.tf;
    // all of a thread's points are requested before the first is used: one memory latency per chunk instead of one per point
    constexpr int kPerThread = (kColChunk + kThreads - 1) / kThreads;
    float4 pre[kPerThread];
#ifdef CVO_CLOUD_EVICT_LAST
    unsigned long long l2_keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(l2_keep));
#endif
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
        const int i = threadIdx.x + u * kThreads;
        pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < ntiles * kTile && base + i < c.n) {
#ifdef CVO_CLOUD_EVICT_LAST  // the clouds are re-read every iteration while the lists stream through L2 between two uses
            asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                         : "=f"(pre[u].x), "=f"(pre[u].y), "=f"(pre[u].z), "=f"(pre[u].w) : "l"(c.g + base + i), "l"(l2_keep));
#else
            pre[u] = __ldg(c.g + base + i);
#endif
        }
    }
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
        const int i = threadIdx.x + u * kThreads;
        if (i >= ntiles * kTile) break;
        const int p = base + i;
        bool valid = p < c.n;
        float4 g;
        if (valid) {
            g = pre[u];
            if (tf) apply_tf(tf12, g.x, g.y, g.z);
            // a point with a NaN / Inf coordinate is nobody's neighbour (d2 < thr is false): move it far away so that
            // the branch-free bodies only ever multiply their zero weights with finite numbers
            valid = finite3(g.x, g.y, g.z);
        }
        if (!valid) g = make_float4(sentinel, sentinel, sentinel, 0.f);
        if (MODE == STAGE_FULL) {
            const float c2 = fmaf(g.z, g.z, fmaf(g.y, g.y, g.x * g.x));
            sm.colG[i] = make_float4(g.x, g.y, g.z, c2);
            const float lx = warp_min(valid ? g.x : inf), ly = warp_min(valid ? g.y : inf), lz = warp_min(valid ? g.z : inf);
            const float hx = warp_max(valid ? g.x : -inf), hy = warp_max(valid ? g.y : -inf), hz = warp_max(valid ? g.z : -inf);
            const float c2m = warp_max(valid ? c2 : 0.f);
            if (lane == 0) {
                float* b = sm.colBox[i >> 5];
                b[0] = lx; b[1] = ly; b[2] = lz; b[3] = hx; b[4] = hy; b[5] = hz; b[6] = c2m;
            }
        } else {  // list passes: planes
            plane_of(sm.colG, 0)[i] = g.x; plane_of(sm.colG, 1)[i] = g.y; plane_of(sm.colG, 2)[i] = g.z;
            if (MODE == STAGE_STEP) {
                const StepCol sc = step_col(sm.ic, g.x, g.y, g.z);
                plane_of(sm.colG, 3)[i] = sc.ecn;
                float* z1 = plane_of(sm.u.ls.ss.colZ1, 0);
                float* z2 = plane_of(sm.u.ls.ss.colZ2, 0);
                z1[i] = sc.z1x; z1[i + kColChunk] = sc.z1y; z1[i + 2 * kColChunk] = sc.z1z; z1[i + 3 * kColChunk] = sc.nrm;
                z2[i] = sc.z2x; z2[i + kColChunk] = sc.z2y; z2[i + 2 * kColChunk] = sc.z2z; z2[i + 3 * kColChunk] = sc.pdt;
            }
        }
    }
    if (MODE == STAGE_FULL) {
        mbar_wait(&sm.tma_bar, tma_phase);
        tma_phase ^= 1u;
    }
}

// --------------------------------------------------------------------------------------------
// per-pair kernel value: the three strict gates of se_kernel (src/cvo.cpp:143-153)
// --------------------------------------------------------------------------------------------
// se_kernel's value and gates in the reference's own arithmetic (src/cvo.cpp:146-152): colour distance summed left
// to right, exp() in f64 narrowed to f32, a = ck * k in f32.  Deliberately not inlined: it runs for about one
// candidate in a million (see kernel_a) and must not cost the hot loops registers.
__device__ __noinline__ float kernel_value_exact_d(float ell, float d2c_thres, float s2, float cs2, float c_ell, float sp_thres,
                                                   float d2c, float d2) {
    const double l = (double)ell, cl = (double)c_ell;
    const float k = (float)((double)s2 * exp(-(double)d2 / (2.0 * l * l)));
    const float ck = (float)((double)cs2 * exp(-(double)d2c / (2.0 * cl * cl)));
    const float a = __fmul_rn(ck, k);
    return ((d2c < d2c_thres) && (a > sp_thres)) ? a : 0.f;  // a > sp_thres > 0 when accepted
}
__device__ __forceinline__ float colour_d2(const float4& xf, float xf4, const float4& yf, float yf4);
__device__ __forceinline__ float kernel_value_exact(float ell, float d2c_thres, float s2, float cs2, float c_ell, float sp_thres,
                                                    float4 xf, float xf4, float4 yf, float yf4, float d2) {
    return kernel_value_exact_d(ell, d2c_thres, s2, cs2, c_ell, sp_thres, colour_d2(xf, xf4, yf, yf4), d2);
}

// (feature_x - feature_y).squaredNorm() summed left to right (src/cvo.cpp:145-146); pose-independent.
__device__ __forceinline__ float colour_d2(const float4& xf, float xf4, const float4& yf, float yf4) {
    const float e0 = xf.x - yf.x, e1 = xf.y - yf.y, e2 = xf.z - yf.z, e3 = xf.w - yf.w, e4 = xf4 - yf4;
    float d2c = __fmul_rn(e0, e0);
    d2c = __fadd_rn(d2c, __fmul_rn(e1, e1));
    d2c = __fadd_rn(d2c, __fmul_rn(e2, e2));
    d2c = __fadd_rn(d2c, __fmul_rn(e3, e3));
    d2c = __fadd_rn(d2c, __fmul_rn(e4, e4));
    return d2c;
}

// k = s2 exp(-d2 / 2l^2), ck = c_sigma^2 exp(-d2c / 2c_ell^2), a = ck k (src/cvo.cpp:149-151) as ONE base-2
// exponential of the summed exponents: a = s2 c_sigma^2 2^-(d2 log2e/2l^2 + t_c), t_c = d2c log2e/2c_ell^2 being
// the pose-independent COLOUR EXPONENT of the pair.  Wherever the result can matter (a > sp_thres => exponent
// < 0.33) MUFU.EX2 is good to 2 ulp and the argument to 1 ulp; `near` flags the candidates whose a lies within a
// few ulp of sp_thres (about one in a million): the caller re-decides those with kernel_value_exact so that the
// gate agrees with the CPU path bit for bit.
template <class IC>
__device__ __forceinline__ float kernel_a(const IC& ic, const KParams& kp, float d2, float t_c, bool& near) {
    const float a = __fmul_rn(kp.s2cs2, exp2f_approx(-fmaf(d2, ic.c1, t_c)));
    near = fabsf(a - kp.sp_thres) < kp.sp_band;
    return a;
}

__device__ __forceinline__ uint32_t* sm_queue(const Smem& sm) {
    return const_cast<uint32_t*>(sm.u.of.fs.queue[threadIdx.x >> 5]);
}

// Per-lane f32 partial sums of one work unit (a few dozen terms each, like the reference's per-row f32 sums,
// src/cvo.cpp:197-198); promoted to f64 when the unit is finished (src/cvo.cpp:202-203).
struct FlowPartial {
    float po0, po1, po2, pv0, pv1, pv2, psum, pdl;
    int cnt;
};

// One nonzero of A in compute_step_size (src/cvo.cpp:260-279): beta, gamma, delta, epsilon from the column's
// step-size terms and r = x_i - y_j, and this nonzero's terms of B, C, D, E.
//   The higher powers need no per-entry cross products: with z1 = omega x y + v,
//     xi^3 z = Omega^2 z1 = omega (omega . z1) - |omega|^2 z1,   omega . z1 = omega . v   (omega . (omega x y) = 0)
//     xi^4 z = Omega^3 z1 = -|omega|^2 (omega x z1) = -|omega|^2 z2                       (Omega^3 = -|omega|^2 Omega)
//   so z3 . r = (omega . v)(omega . r) - |omega|^2 (z1 . r) and z4 . r = -|omega|^2 (z2 . r): three dot products per
//   entry instead of four dot products and two cross products.  (The reference forms the powers as matrix products,
//   src/cvo.cpp:229-234: either way the result is the same to f32 rounding.)
struct StepTerms {
    float tB, tC, tD, tE;
};
template <class IC>
__device__ __forceinline__ StepTerms step_terms(const IC& ic, const StepCol& c, float rx, float ry, float rz, float a) {
    const float w0 = ic.omega[0], w1 = ic.omega[1], w2 = ic.omega[2];
    const float ww = (w0 * w0 + w1 * w1) + w2 * w2;                       // loop invariants: the compiler hoists them
    const float wv = (w0 * ic.v[0] + w1 * ic.v[1]) + w2 * ic.v[2];
    const float p1 = (c.z1x * rx + c.z1y * ry) + c.z1z * rz;
    const float p2 = (c.z2x * rx + c.z2y * ry) + c.z2z * rz;
    const float pw = (w0 * rx + w1 * ry) + w2 * rz;
    // gamma = -t (nrm + 2 p2), delta = 2t (pdt - z3 . r), epsil = -t (ecn + 2 z4 . r) with t = temp_coef, the
    // column's nrm / pdt / ecn already scaled (step_col) and the rest folded into per-iteration constants
    const float kG = -2.f * ic.temp_coef, kDw = -ic.p2t * wv, kD1 = ic.p2t * ww, kE = 2.f * ic.temp_coef * ww;
    const float beta = ic.m2t * p1;                                       // :262
    const float gamma = fmaf(kG, p2, c.nrm);                              // :264
    const float delta = fmaf(kDw, pw, fmaf(kD1, p1, c.pdt));              // :267
    const float epsil = fmaf(kE, p2, c.ecn);                              // :270
    StepTerms t;
#ifdef CVO_STEP_F32_PRODUCTS
    // the reference's own mix of f32 products and f64 sums inside a term (src/cvo.cpp:275-279)
    const double bd = (double)beta, gd = (double)gamma, ad = (double)a;
    t.tB = a * beta;
    t.tC = (float)(ad * (gd + (double)(beta * beta) * 0.5));
    t.tD = (float)(ad * ((double)(delta + beta * gamma) + (double)(beta * beta * beta) * (1.0 / 6.0)));
    t.tE = (float)(ad * ((double)(epsil + beta * delta) + 0.5 * bd * bd * gd + 0.5 * gd * gd + (1.0 / 24.0) * (bd * bd) * (bd * bd)));
#else
    // The terms are f32 in the reference too (its `double(A_ij * (...))` promotes a finished f32-by-f64 expression of
    // f32 inputs); here the whole polynomial is f32 FMAs.
    const float b2 = beta * beta;
    t.tB = a * beta;                                                                                      // :275
    t.tC = a * fmaf(0.5f, b2, gamma);                                                                     // :276
    t.tD = a * fmaf(b2 * beta, 1.f / 6.f, fmaf(beta, gamma, delta));                                      // :277
    t.tE = a * fmaf(b2 * b2, 1.f / 24.f, fmaf(0.5f, gamma * (b2 + gamma), fmaf(beta, delta, epsil)));    // :278-279
#endif
    return t;
}
// on-the-fly passes: every nonzero is promoted and summed in f64 (src/cvo.cpp:275-279)
template <class IC>
__device__ __forceinline__ void step_accumulate(const IC& ic, const StepCol& c, float rx, float ry, float rz, float a,
                                                double* acc) {
    const StepTerms t = step_terms(ic, c, rx, ry, rz, a);
    acc[0] += (double)t.tB; acc[1] += (double)t.tC; acc[2] += (double)t.tD; acc[3] += (double)t.tE;
}

// Accumulation tail of a candidate whose kernel value a is known (a = 0 for a rejected one, which then adds +0
// terms: the bodies are BRANCH-FREE so that the compiler can interleave several of them).
template <int KIND, class IC>
__device__ __forceinline__ void accumulate_terms(const IC& ic, const KParams& kp, const float4& xg, const float4& yg,
                                                 float dx, float dy, float dz, float a, bool ok, bool q1_row,
                                                 FlowPartial& fp, double* acc) {
    if (KIND == PASS_FLOW || KIND == PASS_FLOW_CVO) {
        const float cx = xg.y * yg.z - xg.z * yg.y;  // x_i x y_j, src/cvo.cpp:191
        const float cy = xg.z * yg.x - xg.x * yg.z;
        const float cz = xg.x * yg.y - xg.y * yg.x;
        const float ac = kp.inv_c * a, ad = kp.inv_d * a;  // (1/c*Ai), (1/d*Ai), :197-198
        fp.po0 = fmaf(ac, cx, fp.po0); fp.po1 = fmaf(ac, cy, fp.po1); fp.po2 = fmaf(ac, cz, fp.po2);
        fp.pv0 = fmaf(ad, dx, fp.pv0); fp.pv1 = fmaf(ad, dy, fp.pv1); fp.pv2 = fmaf(ad, dz, fp.pv2);
        fp.psum += a;
        if (KIND == PASS_FLOW) fp.pdl = fmaf(ic.inv_ell3 * a, dx * dx + dy * dy + dz * dz, fp.pdl);  // src/adaptive_cvo.cpp:202,228
        fp.cnt += ok ? 1 : 0;
    } else if (KIND == PASS_XX || KIND == PASS_INNER) {
        fp.pdl = fmaf(ic.inv_ell3 * a, dx * dx + dy * dy + dz * dz, fp.pdl);  // src/adaptive_cvo.cpp:210,231
        fp.psum += a;
        fp.cnt += ok ? 1 : 0;
    } else if (KIND == PASS_YY) {
        // quirk Q1: rows i < num_fixed never fill sum_diff_yy_2 (src/adaptive_cvo.cpp:213-223); :256,259 otherwise
        const float aq = q1_row ? a : 0.f;
        fp.pdl = fmaf(ic.inv_ell3 * aq, dx * dx + dy * dy + dz * dz, fp.pdl);
        fp.cnt += ok ? 1 : 0;
    } else {  // PASS_STEP: src/cvo.cpp:249-289
        step_accumulate(ic, step_col(ic, yg.x, yg.y, yg.z), -dx, -dy, -dz, a, acc);  // diff_xy = x - y, :260
    }
}

// Survivor body of the ON-THE-FLY passes: one (row, col) candidate popped from the warp's queue.  All 32 lanes of
// a warp work on 32 different candidates, so the expensive part runs at full lane utilisation.  The three strict
// gates of se_kernel (exact ball test on the nanoflann-ordered d2, colour gate, a > sp_thres) fold into one
// predicate.
template <int KIND>
__device__ __forceinline__ void survivor_body(const Smem& sm, const WarpScratch& ws, const KParams& kp, uint32_t ent,
                                              bool live, int yy_row_min, FlowPartial& fp, double* acc) {
    const IterConsts& ic = sm.ic;
    const int row = (int)(ent >> 12), col = (int)(ent & 0xfffu);
    const float4 xg = ws.rowG[row];
    const float4 xf = ws.rowF[row];
    const float4 yg = sm.colG[col];
    const float4 yf = sm.u.of.fs.colF[col];
    const float yf4 = sm.u.of.fs.colF4[col];
    const float dx = yg.x - xg.x, dy = yg.y - xg.y, dz = yg.z - xg.z;  // diff_yx, src/cvo.cpp:192
    const float d2 = dist2(dx, dy, dz);
    const float d2c = colour_d2(xf, xg.w, yf, yf4);
    bool near;
    float a = kernel_a(ic, kp, d2, __fmul_rn(d2c, kp.c2), near);
    bool ok = (d2c < ic.d2c_thres) && (a > kp.sp_thres);  // src/cvo.cpp:148,152
    if (near) {
        a = kernel_value_exact(ic.ell, ic.d2c_thres, kp.s2, kp.cs2, kp.c_ell, kp.sp_thres, xf, xg.w, yf, yf4, d2);
        ok = a > 0.f;
    }
    ok = ok && live && (d2 < ic.d2_thres);  // the exact strict ball test (thirdparty/nanoflann.hpp:249-253)
    a = ok ? a : 0.f;
    const bool q1 = (KIND == PASS_YY) ? (ws.rowOrig[row] >= yy_row_min) : true;
    accumulate_terms<KIND>(ic, kp, xg, yg, dx, dy, dz, a, ok, q1, fp, acc);
}

// Survivor body of the LIST passes: the candidate comes with its colour exponent t_c (the colour gate was applied
// when the list was built), so only the geometry is touched; the features are fetched from global memory in the
// one-in-a-million case that a sits within a few ulp of sp_thres.
struct ListSrc {
    const CloudDev* rows;
    const CloudDev* cols;
    int row_base, col_base;  // global index of the unit's row 0 / of the staged chunk's column 0
};
// The per-iteration constants a list body reads, held in registers for the whole pass (IterConsts lives in shared memory).
struct HotConsts {
    float c1, d2_thres, inv_ell3, m2t, temp_coef, p2t;
    float omega[3], v[3];
};
__device__ __forceinline__ HotConsts hot_consts(const IterConsts& ic) {
    HotConsts h;
    h.c1 = ic.c1; h.d2_thres = ic.d2_thres; h.inv_ell3 = ic.inv_ell3;
    h.m2t = ic.m2t; h.temp_coef = ic.temp_coef; h.p2t = ic.p2t;
#pragma unroll
    for (int i = 0; i < 3; ++i) { h.omega[i] = ic.omega[i]; h.v[i] = ic.v[i]; }
    return h;
}

// EXACT = false: the branch-free body of the hot loop.  A candidate whose fast kernel value lies within a few ulp of
// sp_thres (`near`, about one in a million) contributes NOTHING here and is reported to the caller, which re-runs it
// with EXACT = true once the trip's bodies are done: no branch sits between the bodies, so the compiler interleaves
// them, and membership in A still agrees with the CPU path bit for bit.
template <int KIND, bool EXACT>
__device__ __forceinline__ bool list_body(const Smem& sm, const HotConsts& hc, const KParams& kp, uint32_t ent, float t_c,
                                          int yy_row_min, const ListSrc& src, FlowPartial& fp, double* acc) {
    const uint32_t rowb = ent >> 16, colb = ent & 0xffffu;  // byte offsets into the row / column planes
    float4 xg, yg;
    xg.x = plane_ld<0>(sm.u.ls.rowG, rowb); xg.y = plane_ld<1>(sm.u.ls.rowG, rowb); xg.z = plane_ld<2>(sm.u.ls.rowG, rowb);
    yg.x = plane_ld<0>(sm.colG, colb); yg.y = plane_ld<1>(sm.colG, colb); yg.z = plane_ld<2>(sm.colG, colb);
    xg.w = yg.w = 0.f;
    const float dx = yg.x - xg.x, dy = yg.y - xg.y, dz = yg.z - xg.z;  // diff_yx, src/cvo.cpp:192
    const float d2 = dist2(dx, dy, dz);
    bool near = false;
    float a;
    if (EXACT) {
        const IterConsts& ic = sm.ic;
        const int ri = src.row_base + (int)(rowb >> 2), ci = src.col_base + (int)(colb >> 2);
        a = kernel_value_exact(ic.ell, ic.d2c_thres, kp.s2, kp.cs2, kp.c_ell, kp.sp_thres, __ldg(src.rows->f + ri),
                               __ldg(src.rows->f4 + ri), __ldg(src.cols->f + ci), __ldg(src.cols->f4 + ci), d2);
    } else {
        a = kernel_a(hc, kp, d2, t_c, near);
    }
    // src/cvo.cpp:152 and the exact strict ball test (thirdparty/nanoflann.hpp:249-253)
    const bool ok = !near && (a > kp.sp_thres) && (d2 < hc.d2_thres);
    a = ok ? a : 0.f;
    if (KIND == PASS_STEP) {  // the column's step-size terms were computed once, when the chunk was staged
        StepCol c;
        c.z1x = plane_ld<0>(sm.u.ls.ss.colZ1, colb); c.z1y = plane_ld<1>(sm.u.ls.ss.colZ1, colb);
        c.z1z = plane_ld<2>(sm.u.ls.ss.colZ1, colb); c.nrm = plane_ld<3>(sm.u.ls.ss.colZ1, colb);
        c.z2x = plane_ld<0>(sm.u.ls.ss.colZ2, colb); c.z2y = plane_ld<1>(sm.u.ls.ss.colZ2, colb);
        c.z2z = plane_ld<2>(sm.u.ls.ss.colZ2, colb); c.pdt = plane_ld<3>(sm.u.ls.ss.colZ2, colb);
        c.ecn = plane_ld<3>(sm.colG, colb);
        // list passes: the four entries a lane handles in one trip are summed in f32, then promoted (flush_partial)
        const StepTerms t = step_terms(hc, c, -dx, -dy, -dz, a);
        fp.po0 += t.tB; fp.po1 += t.tC; fp.pv0 += t.tD; fp.pv1 += t.tE;
    } else {
        // quirk Q1 is defined on the ORIGINAL row index, carried in the w lane of the staged row
        const bool q1 = (KIND == PASS_YY) ? (__float_as_int(xg.w) >= yy_row_min) : true;
        accumulate_terms<KIND>(hc, kp, xg, yg, dx, dy, dz, a, ok, q1, fp, acc);
    }
    return near;
}

template <int KIND>
__device__ __forceinline__ void flush_partial(FlowPartial& fp, double* acc) {
    if (KIND == PASS_STEP) {  // B, C, D, E: one trip's four entries per lane (list passes only)
        acc[0] += (double)fp.po0; acc[1] += (double)fp.po1; acc[2] += (double)fp.pv0; acc[3] += (double)fp.pv1;
        fp.po0 = fp.po1 = fp.pv0 = fp.pv1 = 0.f;
        return;
    }
    if (fp.cnt) {
        if (KIND == PASS_FLOW || KIND == PASS_FLOW_CVO) {
            acc[ACC_W0] += (double)fp.po0; acc[ACC_W0 + 1] += (double)fp.po1; acc[ACC_W0 + 2] += (double)fp.po2;
            acc[ACC_V0] += (double)fp.pv0; acc[ACC_V0 + 1] += (double)fp.pv1; acc[ACC_V0 + 2] += (double)fp.pv2;
            acc[ACC_SUMA] += (double)fp.psum;
            acc[ACC_NNZ] += (double)fp.cnt;
            acc[ACC_DLXY] += (double)fp.pdl;
        } else if (KIND == PASS_XX || KIND == PASS_YY) {  // {nnz, sum} land in ACC_NNZXX.. / ACC_NNZYY.. later
            acc[0] += (double)fp.cnt;
            acc[1] += (double)fp.pdl;
        } else {  // PASS_INNER
            acc[0] += (double)fp.psum;
            acc[1] += (double)fp.cnt;
        }
    }
    fp.po0 = fp.po1 = fp.po2 = fp.pv0 = fp.pv1 = fp.pv2 = fp.psum = fp.pdl = 0.f;
    fp.cnt = 0;
}

// One (row tile, col tile) pair.  Phase 1 (all lanes busy): lane = one row, 32 candidate columns, a conservative
// ball PREFILTER in expanded form (|c|^2 - 2 c.x < thr + slack - |x|^2: 3 FFMA + compare per candidate) -> 32-bit
// candidate mask.  Phase 2: the candidates of all lanes are compacted into the warp's queue (row, col) and popped
// 32 at a time; the survivor body applies the EXACT strict test on the nanoflann-ordered d2 first, then the
// kernel-value / flow / step arithmetic, on full warps.
struct RowRegs {
    float m2x, m2y, m2z;  // -2 x_i
    float x2;             // |x_i|^2
};
// exclusive prefix sum + total of a per-lane count across the warp
__device__ __forceinline__ void warp_scan_count(int cnt, int lane, int& excl, int& total) {
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    excl = incl - cnt;
}

// appends one queue entry per set bit of `mask` (this lane's candidate columns of one col tile) at q[pos...]
__device__ __forceinline__ void push_mask(uint32_t* q, int pos, uint32_t mask, uint32_t base) {
    while (mask) {
        const int jj = __ffs(mask) - 1;
        mask &= mask - 1;
        q[pos++] = base + (uint32_t)jj;
    }
}

__device__ __forceinline__ uint32_t prefilter_tile(const Smem& sm, const RowRegs& rr, int ct, float thr) {
    const float4* cgp = sm.colG + ct * kTile;
    const float t = fmaf(kPrefilterSlack, sm.colBox[ct][6] + rr.x2, thr) - rr.x2;
    uint32_t mask = 0;
#pragma unroll
    for (int jj = 0; jj < kTile; ++jj) {
        const float4 c = cgp[jj];
        const float s = fmaf(c.x, rr.m2x, fmaf(c.y, rr.m2y, fmaf(c.z, rr.m2z, c.w)));
        mask |= (s < t) ? (1u << jj) : 0u;
    }
    return mask;
}

// One row tile against up to two live col tiles (ctB < 0: only ctA).
template <int KIND>
__device__ __forceinline__ void process_tile_group(const Smem& sm, WarpScratch& ws, const KParams& kp, const RowRegs& rr,
                                                   int ctA, int ctB, int lane, int& qn, int yy_row_min, FlowPartial& fp,
                                                   double* acc) {
    const uint32_t maskA = prefilter_tile(sm, rr, ctA, sm.ic.d2_thres);
    const uint32_t maskB = (ctB >= 0) ? prefilter_tile(sm, rr, ctB, sm.ic.d2_thres) : 0u;
    if (__ballot_sync(0xffffffffu, (maskA | maskB) != 0) == 0) return;
    uint32_t* q = sm_queue(sm);
    // the queue holds one full tile pair on top of the leftovers: a (rare) group with more candidates than that
    // is pushed in two rounds (col tile A, then col tile B)
    uint32_t mA = maskA, mB = maskB;
    bool pending = false;
    while (true) {
        const int nA = __popc(mA);
        int excl, total;
        warp_scan_count(nA + __popc(mB), lane, excl, total);
        if (total > kTile * kTile) {  // only possible for the combined round
            mB = 0u;
            pending = true;
            continue;
        }
        const int pos = qn + excl;
        push_mask(q, pos, mA, ((uint32_t)lane << 12) | (uint32_t)(ctA * kTile));
        push_mask(q, pos + nA, mB, ((uint32_t)lane << 12) | (uint32_t)(ctB * kTile));
        qn += total;
        __syncwarp();
#if CVO_BODY_ILP >= 2
        while (qn >= 64) {  // two independent candidates per lane: the compiler interleaves the two bodies
            qn -= 64;
            const uint32_t e0 = q[qn + lane], e1 = q[qn + 32 + lane];
            survivor_body<KIND>(sm, ws, kp, e0, true, yy_row_min, fp, acc);
            survivor_body<KIND>(sm, ws, kp, e1, true, yy_row_min, fp, acc);
        }
#else
        while (qn >= 32) {
            qn -= 32;
            survivor_body<KIND>(sm, ws, kp, q[qn + lane], true, yy_row_min, fp, acc);
        }
#endif
        __syncwarp();
        if (!pending) break;
        pending = false;
        mA = 0u;
        mB = maskB;
    }
}

template <int KIND> struct PassTraits;
template <> struct PassTraits<PASS_FLOW>  { static constexpr int NV = 9; };
template <> struct PassTraits<PASS_XX>    { static constexpr int NV = 2; };
template <> struct PassTraits<PASS_YY>    { static constexpr int NV = 2; };
template <> struct PassTraits<PASS_STEP>  { static constexpr int NV = 4; };
template <> struct PassTraits<PASS_INNER> { static constexpr int NV = 2; };
template <> struct PassTraits<PASS_FLOW_CVO> { static constexpr int NV = 9; };

// One work unit = one 32-row tile against one segment of the staged column tiles, done by ONE warp with no
// block-level synchronisation.  The unit's totals go to its own slot, so the block sum does not depend on
// which warp ran which unit (bit-deterministic under dynamic scheduling).
template <int KIND>
__device__ __forceinline__ void process_unit(Smem& sm, const KParams& kp, const CloudDev& rows, bool row_tf, int tile,
                                             int ct_begin, int ct_end, int slot, bool first_chunk, int yy_row_min) {
    constexpr int NV = PassTraits<KIND>::NV;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = sm.u.of.ws[warp];
    const float inf = __int_as_float(0x7f800000);
    // stage the row tile: registers for the mask phase, warp-private shared memory for the survivor body
    const int p = tile * kTile + lane;
    bool valid = p < rows.n;
    float4 xg = make_float4(0.f, 0.f, 0.f, 0.f), xf = make_float4(0.f, 0.f, 0.f, 0.f);
    int orig = -1;
    if (valid) {
        xg = __ldg(rows.g + p);
        xf = __ldg(rows.f + p);
        orig = __float_as_int(xg.w);
        xg.w = __ldg(rows.f4 + p);
        if (row_tf) apply_tf(sm.ic.tf, xg.x, xg.y, xg.z);
        valid = finite3(xg.x, xg.y, xg.z);  // (see stage_tiles)
    }
    if (!valid) xg = make_float4(kRowSentinel, kRowSentinel, kRowSentinel, xg.w);
    __syncwarp();  // the previous unit's body reads are done
    ws.rowG[lane] = xg;
    ws.rowF[lane] = xf;
    if (KIND == PASS_YY) ws.rowOrig[lane] = orig;
    const float lx = warp_min(valid ? xg.x : inf), ly = warp_min(valid ? xg.y : inf), lz = warp_min(valid ? xg.z : inf);
    const float hx = warp_max(valid ? xg.x : -inf), hy = warp_max(valid ? xg.y : -inf), hz = warp_max(valid ? xg.z : -inf);
    __syncwarp();

    RowRegs rr;
    rr.m2x = -2.f * xg.x; rr.m2y = -2.f * xg.y; rr.m2z = -2.f * xg.z;
    rr.x2 = fmaf(xg.z, xg.z, fmaf(xg.y, xg.y, xg.x * xg.x));
    FlowPartial fp;
    fp.po0 = fp.po1 = fp.po2 = fp.pv0 = fp.pv1 = fp.pv2 = fp.psum = fp.pdl = 0.f;
    fp.cnt = 0;
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    const float thr = sm.ic.d2_thres * 1.0001f;  // boxes are conservative; keep rounding on the safe side
    int qn = 0;                                  // pairs waiting in this warp's queue (warp-uniform)
    for (int c0 = ct_begin; c0 < ct_end; c0 += 32) {
        const int ct = c0 + lane;
        bool live = false;
        if (ct < ct_end) {  // lane tests one column-tile box against the row-tile box
            const float* b = sm.colBox[ct];
            const float gx = fmaxf(0.f, fmaxf(lx - b[3], b[0] - hx));
            const float gy = fmaxf(0.f, fmaxf(ly - b[4], b[1] - hy));
            const float gz = fmaxf(0.f, fmaxf(lz - b[5], b[2] - hz));
            live = (gx * gx + gy * gy + gz * gz) <= thr;
        }
        uint32_t lm = __ballot_sync(0xffffffffu, live);
        while (lm) {
            const int jA = __ffs(lm) - 1;
            lm &= lm - 1;
            int jB = -1 - c0;
#if CVO_GROUP >= 2
            if (lm) {
                jB = __ffs(lm) - 1;
                lm &= lm - 1;
            }
#endif
            process_tile_group<KIND>(sm, ws, kp, rr, c0 + jA, c0 + jB, lane, qn, yy_row_min, fp, acc);
        }
    }
    for (int b = 0; b < qn; b += 32) {  // drain the tail of the queue; idle lanes run entry (0, 0) with a = 0
        const bool live = b + lane < qn;
        survivor_body<KIND>(sm, ws, kp, live ? sm_queue(sm)[b + lane] : 0u, live, yy_row_min, fp, acc);
    }
    __syncwarp();
    flush_partial<KIND>(fp, acc);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double t = warp_sum(acc[i]);
        if (lane == 0) {
            if (first_chunk) sm.u.of.unitPart[slot][i] = t;
            else sm.u.of.unitPart[slot][i] += t;
        }
    }
}

// How one all-pairs pass of rows x cols is cut into work units for the CTA of rank `rank` in a cluster of G:
// a pure function of the sizes, so that the neighbour-list build and every later pass over the list agree.
struct PassGeom {
    int t_begin, my_tiles, total_ct, S, tiles_per_round;
};
__device__ __forceinline__ PassGeom pass_geom(int rows_n, int cols_n, int rank, int G) {
    PassGeom pg;
    const int total_rt = (rows_n + kTile - 1) / kTile;
    pg.t_begin = (total_rt * rank) / G;  // total_rt <= 512, G <= 16
    const int t_end = (total_rt * (rank + 1)) / G;
    pg.my_tiles = t_end - pg.t_begin;
    pg.total_ct = (cols_n + kTile - 1) / kTile;
    // split every row tile's column range into S segments so that there are >= ~4 units per warp
    int S = 1;
    if (pg.my_tiles > 0) {
        S = (CVO_UNITS_PER_WARP * kWorkWarps + pg.my_tiles - 1) / pg.my_tiles;
        const int s_max = max(1, min(pg.total_ct, kColTiles) / 8);
        S = max(1, min(min(S, s_max), kMaxUnits));
    }
    pg.S = S;
    pg.tiles_per_round = max(1, min(kMaxUnits / S, kColTiles));  // a round's rows fit the list passes' row stage
    return pg;
}

// One all-pairs pass of `rows` x `cols` restricted to this CTA's share of the row tiles.  The column cloud is
// staged (and transformed) once per chunk; warps then pull work units from a shared counter.  On return
// sm.blockTot[0 .. NV) holds this CTA's totals (valid for threads after the final barrier).
template <int KIND>
__device__ void run_pass(Smem& sm, const KParams& kp, const CloudDev& rows, bool row_tf, const CloudDev& cols,
                         bool col_tf, int rank, int G, int yy_row_min, uint32_t& tma_phase) {
    constexpr int NV = PassTraits<KIND>::NV;
    const int lane = threadIdx.x & 31;
    const PassGeom pg = pass_geom(rows.n, cols.n, rank, G);
    const int t_begin = pg.t_begin, my_tiles = pg.my_tiles, total_ct = pg.total_ct, S = pg.S;
    const int tiles_per_round = pg.tiles_per_round;
    if (threadIdx.x < kNumAcc) sm.blockTot[threadIdx.x] = 0.0;
    if (threadIdx.x == 0) sm.colTag.serial = sm.rowTag.serial = -2;  // this pass overwrites the list passes' stages
    for (int rb = 0; rb < my_tiles; rb += tiles_per_round) {
        const int ntile = min(tiles_per_round, my_tiles - rb);
        const int nunits = ntile * S;
        for (int cb = 0; cb < total_ct; cb += kColTiles) {
            const int nct = min(kColTiles, total_ct - cb);
            __syncthreads();  // everyone is done with the previous column chunk / unit slots
            stage_tiles<STAGE_FULL>(sm, cols, cb * kTile, nct, col_tf, kColSentinel, tma_phase);
            if (threadIdx.x == 0) sm.next_unit = 0;
            __syncthreads();
            while ((threadIdx.x >> 5) < kWorkWarps) {
                int u = 0;
                if (lane == 0) u = atomicAdd(&sm.next_unit, 1);
                u = __shfl_sync(0xffffffffu, u, 0);
                if (u >= nunits) break;
                const int t = u / S, seg = u - t * S;
                const int c_begin = (int)(((long long)nct * seg) / S), c_end = (int)(((long long)nct * (seg + 1)) / S);
                process_unit<KIND>(sm, kp, rows, row_tf, t_begin + rb + t, c_begin, c_end, u, cb == 0, yy_row_min);
            }
        }
        __syncthreads();
        if (threadIdx.x < NV) {  // fixed-order sum over the unit slots
            double t = 0.0;
            for (int u = 0; u < nunits; ++u) t += sm.u.of.unitPart[u][threadIdx.x];
            sm.blockTot[threadIdx.x] += t;
        }
    }
    __syncthreads();
}


// --------------------------------------------------------------------------------------------
// neighbour candidate lists
// --------------------------------------------------------------------------------------------
// A list entry is (row * 4 << 16 | col * 4, t_c): the index pair (as byte offsets into the planes of the staged rows / columns
// of its round) and its pose-independent colour exponent
// t_c = |f_i - g_j|^2 log2(e) / (2 c_ell^2).  With T = log2(s2 c_sigma^2 / sp_thres) the gate a > sp_thres reads
// d2 log2(e)/(2 l^2) + t_c < T, i.e. every pair has its OWN ball radius r_e = sqrt((T - t_c) 2 l^2 / log2 e) <= r
// (equal colours: r_e = r, the ell-ball; a colour mismatch shrinks it; t_c >= T or a failed colour gate: never a
// neighbour).  The build keeps a pair iff |x_i - y_j| < r_e + s at the build pose, s = skin * r being the slack.
//
// Validity.  At a later iteration (transform T1, length-scale l1) the pair can only pass if |x_i - T1 y_j| <
// r_e (l1/l0); it is in the list if |x_i - T0 y_j| < r_e + s, and |x_i - T0 y_j| <= |x_i - T1 y_j| + disp with
// disp = max_j |(M1 - M0) y_j + (t1 - t0)|.  Since r_e <= r0, the list covers everything that can pass as long as
// max(0, r1 - r0) + disp <= s.  The (x, x) list never moves and rigid motion preserves the (y, y) distances (up to
// the f32 rounding of the transformed coordinates, covered by the margin): those two only follow ell.
// Called by all lanes of warp 0 (after lane 0 ran prepare_iter and a __syncwarp): the 8 box corners go to 8 lanes.
__device__ void list_policy(Smem& sm, bool acvo, float skin, float skin_min, float shrink, float refine_min, float wide_factor) {
    const int lane = threadIdx.x & 31;
    // |(M1 - M0) p + (t1 - t0)| is convex in p: its maximum over the moving cloud's bounding box is attained at one of
    // the 8 corners.  Lanes 0..7: against the pose the (x, y) list was built at; lanes 8..15: against the wide list's.
    double disp_xy = 0.0, disp_wide = 0.0;
    {
        const bool wide_half = (lane & 8) != 0;
        const float* tf0 = wide_half ? sm.wide.tf : sm.lst[LIST_XY].tf;
        const bool have = wide_half ? sm.wide.valid > 0 : sm.lst[LIST_XY].valid > 0;
        const int c = lane & 7;
        // f32 throughout: the differences of the transform entries are exact or nearly so (neighbouring poses), the
        // rounding of the rest (~1e-7 relative of a displacement of centimetres) is five orders below `margin`; the
        // result is rounded UP by 1e-5 relative before it is trusted
        float dm[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) dm[i] = sm.ic.tf[i] - tf0[i];
        const float px = sm.ybox[(c & 1) ? 3 : 0], py = sm.ybox[(c & 2) ? 4 : 1], pz = sm.ybox[(c & 4) ? 5 : 2];
        const float ex = dm[0] * px + dm[1] * py + dm[2] * pz + dm[9];
        const float ey = dm[3] * px + dm[4] * py + dm[5] * pz + dm[10];
        const float ez = dm[6] * px + dm[7] * py + dm[8] * pz + dm[11];
        float d = sqrtf(ex * ex + ey * ey + ez * ez) * 1.00001f;
        if (!(d == d) || !have) d = 1.0e30f;  // NaN state / no such list: never trust it
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
        disp_xy = (double)__shfl_sync(0xffffffffu, d, 0);
        disp_wide = (double)__shfl_sync(0xffffffffu, d, 8);
    }
    __syncwarp();  // every lane has read the build transforms before lane 0 may replace them
    if (lane != 0) return;
    const double r_now = sqrt((double)sm.ic.d2_thres);
    const double margin = 2.0e-5 + 1.0e-5 * r_now;  // f32 rounding of the transformed coordinates, generously
    const int nk = acvo ? LIST_KINDS : 1;
    sm.wide.make = 0;
    for (int kind = 0; kind < nk; ++kind) {
        ListState& L = sm.lst[kind];
        if (L.valid < 0) {  // overflowed earlier for this pair: stay on the fly
            L.need = 0;
            continue;
        }
        bool need = L.valid == 0;
        if (!need) {
            const double disp = kind == LIST_XY ? disp_xy : 0.0;
            // rebuild when something that can pass may be missing, or when ell has shrunk the ball a lot (a list that
            // is much too wide costs more in every pass than one rebuild)
            need = !(fmax(0.0, r_now - (double)L.r0) + disp <= (double)L.slack) || (r_now < (double)shrink * (double)L.r0);
        }
        L.need = need ? 1 : 0;
        if (need) {
            double s = fmax((double)skin * r_now, (double)skin_min);
            if (kind != LIST_XY && L.valid > 0 && r_now <= (double)L.r0) {  // (the (x, y) list is rebuilt: its quads are row-sorted)
                // The ball has shrunk and the old list still covers the pose with room to spare: everything the new
                // list must hold (|x_i - T1 y_j| < r_e1 + s1, r_e1 <= r_e0) is in the old one as long as
                // s1 + disp <= s0, so the new list is a FILTER of the old one (refine_list) -- no all-pairs sweep.
                const double left = (double)L.slack - margin;
                if (left >= (double)refine_min * s) {
                    s = fmin(s, left);
                    L.need = 2;
                }
            }
            if (kind == LIST_XY && wide_factor > 0.f) {
                // the quads as a filter of the wide list (need = 3) while it covers them -- and is not much too wide itself
                WideState& Wd = sm.wide;
                const bool covers = Wd.valid > 0 && fmax(0.0, r_now - (double)Wd.r0) + disp_wide + s + margin <= (double)Wd.slack &&
                                    r_now >= 0.6 * (double)Wd.r0;
                if (covers) {
                    L.need = 3;
                } else {  // this sweep writes a new wide list as well
                    const double sw = s + (double)wide_factor * r_now;
                    const double rbw = r_now + sw + margin;
                    Wd.make = 1;
                    Wd.valid = 0;
                    Wd.r0 = (float)r_now;
                    Wd.slack = (float)(sw * (1.0 - 1.0e-6));
                    Wd.s_build = (float)((sw + margin) * (1.0 + 1.0e-6));
                    Wd.thr_build = (float)(rbw * rbw * (1.0 + 1.0e-6));
#pragma unroll
                    for (int i = 0; i < 12; ++i) Wd.tf[i] = sm.ic.tf[i];
                }
            }
            const double rb = r_now + s + margin;
            if (L.need == 1) L.valid = 0;
            L.r0 = (float)r_now;
            L.slack = (float)(s * (1.0 - 1.0e-6));            // rounded DOWN: what the validity test may assume
            L.s_build = (float)((s + margin) * (1.0 + 1.0e-6));  // rounded UP: what the build adds to r_e
            L.thr_build = (float)(rb * rb * (1.0 + 1.0e-6));  // rounded UP: the build prefilter's ball
            L.inv_c1 = (float)(2.0 * (double)sm.st.ell * (double)sm.st.ell / 1.4426950408889634 * (1.0 + 1.0e-6));
#pragma unroll
            for (int i = 0; i < 12; ++i) L.tf[i] = sm.ic.tf[i];
            if (L.need >= 2) sm.st.n_refines += 1;
            else if (kind == LIST_XY) sm.st.n_builds += 1;
        }
    }
}

// Row tile of a build / list unit: registers for the prefilter, warp-private shared memory for the per-candidate work.
struct RowTile {
    RowRegs rr;
    float lx, ly, lz, hx, hy, hz;  // bounding box of the valid rows
};
template <bool NEED_FEAT, bool NEED_ORIG, bool NEED_BOX>
__device__ __forceinline__ RowTile load_row_tile(const Smem& sm, WarpScratch& ws, const CloudDev& rows, bool row_tf, int tile) {
    const int lane = threadIdx.x & 31;
    const float inf = __int_as_float(0x7f800000);
    const int p = tile * kTile + lane;
    bool valid = p < rows.n;
    float4 xg, xf = make_float4(0.f, 0.f, 0.f, 0.f);
    int orig = -1;
    if (valid) {
        xg = __ldg(rows.g + p);
        orig = __float_as_int(xg.w);
        xg.w = 0.f;
        if (NEED_FEAT) {
            xf = __ldg(rows.f + p);
            xg.w = __ldg(rows.f4 + p);
        }
        if (row_tf) apply_tf(sm.ic.tf, xg.x, xg.y, xg.z);
        valid = finite3(xg.x, xg.y, xg.z);  // (see stage_tiles)
    }
    if (!valid) xg = make_float4(kRowSentinel, kRowSentinel, kRowSentinel, xg.w);
    __syncwarp();  // the previous unit's reads of the warp scratch are done
    ws.rowG[lane] = xg;
    if (NEED_FEAT) ws.rowF[lane] = xf;
    if (NEED_ORIG) ws.rowOrig[lane] = orig;
    RowTile rt;
    if (NEED_BOX) {
        rt.lx = warp_min(valid ? xg.x : inf); rt.ly = warp_min(valid ? xg.y : inf); rt.lz = warp_min(valid ? xg.z : inf);
        rt.hx = warp_max(valid ? xg.x : -inf); rt.hy = warp_max(valid ? xg.y : -inf); rt.hz = warp_max(valid ? xg.z : -inf);
    } else {
        rt.lx = rt.ly = rt.lz = rt.hx = rt.hy = rt.hz = 0.f;
    }
    __syncwarp();
    rt.rr.m2x = -2.f * xg.x; rt.rr.m2y = -2.f * xg.y; rt.rr.m2z = -2.f * xg.z;
    rt.rr.x2 = fmaf(xg.z, xg.z, fmaf(xg.y, xg.y, xg.x * xg.x));
    return rt;
}

// ballot of the column tiles [c0, c0 + 32) of the unit whose box is within sqrt(thr) of the row tile's box
__device__ __forceinline__ uint32_t live_col_tiles(const Smem& sm, const RowTile& rt, int c0, int ct_end, float thr) {
    const int ct = c0 + (threadIdx.x & 31);
    bool live = false;
    if (ct < ct_end) {
        const float* b = sm.colBox[ct];
        const float gx = fmaxf(0.f, fmaxf(rt.lx - b[3], b[0] - rt.hx));
        const float gy = fmaxf(0.f, fmaxf(rt.ly - b[4], b[1] - rt.hy));
        const float gz = fmaxf(0.f, fmaxf(rt.lz - b[5], b[2] - rt.hz));
        live = (gx * gx + gy * gy + gz * gz) <= thr;
    }
    return __ballot_sync(0xffffffffu, live);
}

// Build, per candidate: exact distance at the build pose, colour gate and colour exponent (pose-independent,
// src/cvo.cpp:145-148), and the pair's own radius + slack.  Survivors are appended to the unit's staging region with
// their row index made relative to the round (`row_off` = 32 * the unit's row tile within the round).
// SELF = 0: an (x, y) list entry (byte offsets of the pair, colour exponent).  SELF = 1 / 2: the (x, x) / (y, y) list of
// acvo, whose distances never change (x is never transformed, rigid motion preserves |y_i - y_j|): the entry IS the
// pair of invariants (d2, colour d2), and a pass over it touches no point data at all.  For (y, y) the sign bit of the
// colour distance marks the rows that contribute to the length-scale gradient (always for (x, x); for (y, y) quirk Q1:
// original index >= num_fixed).
// Where a sweep writes the WIDE (x, y) list (WideState): this warp's share of the wide area.
struct WideOut {
    uint2* out;
    int limit, cursor;
    float s_build;  // 0 = this sweep writes no wide list
};
template <int SELF>
__device__ __forceinline__ bool build_test(const Smem& sm, const WarpScratch& ws, const KParams& kp, const ListState& L,
                                           uint32_t ent, bool live, uint32_t row_off, int yy_row_min, uint2& e,
                                           float wide_s_build, bool& keep_wide) {
    const int row = (int)(ent >> 12), col = (int)(ent & 0xfffu);
    const float4 xg = ws.rowG[row];
    const float4 xf = ws.rowF[row];
    const float4 yg = sm.colG[col];
    const float4 yf = sm.u.of.fs.colF[col];
    const float yf4 = sm.u.of.fs.colF4[col];
    const float d2 = dist2(yg.x - xg.x, yg.y - xg.y, yg.z - xg.z);
    const float d2c = colour_d2(xf, xg.w, yf, yf4);
    const float t_c = __fmul_rn(d2c, kp.c2);
    const float re2 = (kp.t_lim - t_c) * L.inv_c1;  // the pair's own squared ball radius (rounded up)
    const float lim = sqrtf_approx(fmaxf(re2, 0.f)) * 1.000002f + L.s_build;
    if (SELF == 0) {  // staged candidate of an (x, y) build: col | row within the tile << 12 (build_append adds rank << 17)
        e = make_uint2(((uint32_t)row << 12) | (uint32_t)col, __float_as_uint(t_c));
    } else {
        const bool q1 = SELF == 1 || ws.rowOrig[row] >= yy_row_min;  // (x, x): every row counts
        e = make_uint2(__float_as_uint(d2), __float_as_uint(d2c) | (q1 ? 0x80000000u : 0u));
    }
    const bool gates = live && (d2c < sm.ic.d2c_thres) && (re2 > 0.f);
    if (SELF == 0) {  // the wide list's ball around the same pair (wide_s_build = 0: not asked for)
        const float limw = lim - L.s_build + wide_s_build;
        keep_wide = gates && (d2 < limw * limw * 1.000001f);
    }
    return gates && (d2 < lim * lim * 1.000001f);
}
// appends the kept candidates of one warp-wide batch in lane order
// (a unit that outgrows the warp's staging segment keeps counting without storing: the build then reports overflow)
// SELF == 0: a row tile of an (x, y) build belongs to ONE warp for the whole column range, and its 32 rows map onto the
// warp's 32 lanes: lane r keeps row r's candidate count in a REGISTER (`row_cnt`).  Every kept candidate gets its RANK
// within its row here -- the row's count so far (one shuffle) plus its position among the batch's lanes with the same
// row.  Those lane sets come from six ballots (keep + the five bits of the row), combined per lane with a few logic
// operations: no shared memory, no warp barrier, nothing the compiler cannot interleave with the next batch.  The rank
// is stored with the candidate, which makes the row-sorted compaction (compact_quads) a scatter of independent entries.
template <int SELF>
__device__ __forceinline__ void build_append(bool keep, uint2 e, uint2* out, int limit, int& cursor, int& row_cnt) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t b = __ballot_sync(0xffffffffu, keep);
    if (SELF == 0) {
        const uint32_t row = (e.x >> 12) & 31u;
        uint32_t same = b, mine = b;  // kept lanes whose row is this lane's candidate's row / whose row is this LANE
#pragma unroll
        for (int bit = 0; bit < 5; ++bit) {
            const uint32_t bb = __ballot_sync(0xffffffffu, (row >> bit) & 1u);
            same &= ((row >> bit) & 1u) ? bb : ~bb;
            mine &= ((lane >> bit) & 1u) ? bb : ~bb;
        }
        const int base = __shfl_sync(0xffffffffu, row_cnt, (int)row);
        e.x |= (uint32_t)(base + __popc(same & ((1u << lane) - 1u))) << 17;
        row_cnt += __popc(mine);
    }
    if (keep && cursor + kTile <= limit) __stcg(out + cursor + __popc(b & ((1u << lane) - 1u)), e);
    cursor += __popc(b);
}
// appends the batch's candidates inside the wide list's ball to the warp's share of the wide area (lane order, no ranks)
__device__ __forceinline__ void wide_append(bool keep, const uint2& e, WideOut& wo) {
    const int lane = threadIdx.x & 31;
    const uint32_t b = __ballot_sync(0xffffffffu, keep);
    if (keep && wo.cursor + kTile <= wo.limit) __stcg(wo.out + wo.cursor + __popc(b & ((1u << lane) - 1u)), e);
    wo.cursor += __popc(b);
}
template <int SELF>
__device__ __forceinline__ void build_eval(Smem& sm, const WarpScratch& ws, const KParams& kp, const ListState& L,
                                           uint32_t ent, bool live, uint32_t row_off, int yy_row_min, uint2* out, int limit,
                                           int& cursor, WideOut& wo, int& row_cnt) {
    uint2 e;
    bool kw = false;
    const bool keep = build_test<SELF>(sm, ws, kp, L, ent, live, row_off, yy_row_min, e, wo.s_build, kw);
    if (SELF == 0 && wo.s_build > 0.f) wide_append(kw, e, wo);
    build_append<SELF>(keep, e, out, limit, cursor, row_cnt);
}
// two batches at once: their loads and arithmetic interleave (the evaluation is latency-bound on one batch)
template <int SELF>
__device__ __forceinline__ void build_eval2(Smem& sm, const WarpScratch& ws, const KParams& kp, const ListState& L,
                                            uint32_t ent0, uint32_t ent1, uint32_t row_off, int yy_row_min, uint2* out,
                                            int limit, int& cursor, WideOut& wo, int& row_cnt) {
    uint2 e0, e1;
    bool kw0 = false, kw1 = false;
    const bool k0 = build_test<SELF>(sm, ws, kp, L, ent0, true, row_off, yy_row_min, e0, wo.s_build, kw0);
    const bool k1 = build_test<SELF>(sm, ws, kp, L, ent1, true, row_off, yy_row_min, e1, wo.s_build, kw1);
    if (SELF == 0 && wo.s_build > 0.f) {
        wide_append(kw0, e0, wo);
        wide_append(kw1, e1, wo);
    }
    build_append<SELF>(k0, e0, out, limit, cursor, row_cnt);
    build_append<SELF>(k1, e1, out, limit, cursor, row_cnt);
}

template <int SELF>
__device__ __forceinline__ int build_unit_write(Smem& sm, WarpScratch& ws, const KParams& kp, const ListState& L,
                                                const CloudDev& rows, bool row_tf, int tile, uint32_t row_off, int yy_row_min,
                                                int ct_begin, int ct_end, uint2* out, int limit, float thr_build, WideOut& wo) {
    const int lane = threadIdx.x & 31;
    const RowTile rt = load_row_tile<true, SELF == 2, true>(sm, ws, rows, row_tf, tile);
    const float thr_box = thr_build * 1.0001f;
    uint32_t* q = sm_queue(sm);
    int qn = 0, cursor = 0, row_cnt = 0;
    for (int c0 = ct_begin; c0 < ct_end; c0 += 32) {
        uint32_t lm = live_col_tiles(sm, rt, c0, ct_end, thr_box);
        while (lm) {
            const int j = __ffs(lm) - 1;
            lm &= lm - 1;
            const uint32_t mask = prefilter_tile(sm, rt.rr, c0 + j, thr_build);
            if (__ballot_sync(0xffffffffu, mask != 0) == 0) continue;
            int excl, total;
            warp_scan_count(__popc(mask), lane, excl, total);
            push_mask(q, qn + excl, mask, ((uint32_t)lane << 12) | (uint32_t)((c0 + j) * kTile));
            qn += total;
            __syncwarp();
            while (qn >= 64) {
                qn -= 64;
                build_eval2<SELF>(sm, ws, kp, L, q[qn + 32 + lane], q[qn + lane], row_off, yy_row_min, out, limit, cursor, wo, row_cnt);
            }
            if (qn >= 32) {
                qn -= 32;
                build_eval<SELF>(sm, ws, kp, L, q[qn + lane], true, row_off, yy_row_min, out, limit, cursor, wo, row_cnt);
            }
            __syncwarp();
        }
    }
    if (qn > 0) build_eval<SELF>(sm, ws, kp, L, lane < qn ? q[lane] : 0u, lane < qn, row_off, yy_row_min, out, limit, cursor, wo, row_cnt);
    if (SELF == 0) sm.u.of.bu.rowCnt[row_off + lane] = row_cnt;  // lane r: candidates kept for row r of the tile
    __syncwarp();
    return cursor;
}

// A row tile of the (x, y) list as a FILTER of the wide list (WideState): the tile's wide candidates stream past, those
// inside r_e + s of the current pose and length-scale (build_test's criterion; the colour gate was applied by the sweep)
// are ranked and staged exactly like a sweep's.
__device__ __forceinline__ int filter_unit_write(Smem& sm, WarpScratch& ws, const KParams& kp, const ListState& L,
                                                 const CloudDev& rows, int tile, uint32_t row_off, const uint2* src, int n,
                                                 uint2* out, int limit) {
    const int lane = threadIdx.x & 31;
    load_row_tile<false, false, false>(sm, ws, rows, false, tile);
    int cursor = 0, row_cnt = 0;
    for (int i1 = 0; i1 < n; i1 += 128) {  // four batches per step: loads, then tests, then appends
        uint2 ev[4];
        bool keep[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ev[j] = (i1 + 32 * j + lane < n) ? __ldcg(src + i1 + 32 * j + lane) : make_uint2(0u, 0x7f800000u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // (idle lanes: row 0, column 0, t_c = +inf: never kept)
            const float4 xg = ws.rowG[(ev[j].x >> 12) & 31u];
            const float4 yg = sm.colG[ev[j].x & 0xfffu];
            const float d2 = dist2(yg.x - xg.x, yg.y - xg.y, yg.z - xg.z);
            const float re2 = (kp.t_lim - __uint_as_float(ev[j].y)) * L.inv_c1;
            const float lim = sqrtf_approx(fmaxf(re2, 0.f)) * 1.000002f + L.s_build;
            keep[j] = (re2 > 0.f) && (d2 < lim * lim * 1.000001f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i1 + 32 * j >= n) break;
            build_append<0>(keep[j], ev[j], out, limit, cursor, row_cnt);
        }
    }
    sm.u.of.bu.rowCnt[row_off + lane] = row_cnt;
    __syncwarp();
    return cursor;
}

// pulls the next work unit of a sweep from the CTA's shared counter (warp-uniform result)
__device__ __forceinline__ int next_unit(Smem& sm) {
    int u = 0;
    if ((threadIdx.x & 31) == 0) u = atomicAdd(&sm.next_unit, 1);
    return __shfl_sync(0xffffffffu, u, 0);
}

// (Re)builds one neighbour list for this CTA's share of the row tiles.  Per round (row chunk x column chunk):
//   evaluate  every warp pulls work units (row tile x column segment), runs prefilter -> queue -> per-candidate
//             evaluation and appends the unit's entries to ITS OWN segment of the staging area;
//   compact   an exclusive scan of the unit counts in unit order gives every unit its place in the round's FLAT
//             list, the entries are copied there and the round is padded to a whole trip with entries that can never
//             pass ((row 0, col 0) are real points, t_c = +inf gives a = 0).
// Which warp evaluated which unit does not matter: the flat list is a pure function of the inputs.  On return
// sm.lst[kind].valid is 1, or -1 if a scratch area was too small.
template <int SELF>
__device__ __forceinline__ bool compact_quads(Smem& sm, const ListRef& lr, int kind, int round, int ntile, int Sb);

// (x, y) list only -- `from_wide`: the round's candidates come from the wide list (filter_unit_write) instead of the
// all-pairs sweep; otherwise, if sm.wide.make is set, the sweep also writes a new wide list (per warp a share of the wide
// area that runs on from round to round; per (round, row tile) an (offset, count) record at the head of the area).
template <int SELF>
__device__ void build_list(Smem& sm, const KParams& kp, const CloudDev& rows, bool row_tf, const CloudDev& cols, bool col_tf,
                           int rank, int G, int yy_row_min, uint32_t& tma_phase, int kind, const ListRef& lr,
                           bool from_wide = false) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PassGeom pg = pass_geom(rows.n, cols.n, rank, G);
    const ListState& L = sm.lst[kind];
    WarpScratch& ws = sm.u.of.ws[warp < kWorkWarps ? warp : 0];
    const int seg = (int)(lr.cap / kWorkWarps) & ~3;  // this warp's staging segment
    uint2* const stage = lr.staging + (size_t)warp * seg;
    const bool make_wide = SELF == 0 && !from_wide && sm.wide.make != 0;
    const int wseg = (int)((lr.cap - kWideTable) / kWorkWarps) & ~3;  // this warp's share of the wide area
    WideOut wo;
    wo.out = lr.wide + kWideTable + (size_t)warp * wseg;
    wo.limit = wseg;
    wo.cursor = 0;
    wo.s_build = make_wide ? sm.wide.s_build : 0.f;
    const float thr_build = make_wide ? sm.wide.thr_build : L.thr_build;
    if (threadIdx.x == 0) {
        sm.lst_used = 0;
        sm.lst_ovf = 0;
        sm.wide_ovf = 0;
        sm.colTag.serial = sm.rowTag.serial = -2;  // the build's stage overwrites the list passes' stages
    }
    int round = 0;
    bool stop = false;
    for (int rb = 0; rb < pg.my_tiles && !stop; rb += pg.tiles_per_round) {
        const int ntile = min(pg.tiles_per_round, pg.my_tiles - rb);
        for (int cb = 0; cb < pg.total_ct && !stop; cb += kColTiles, ++round) {
            const int nct = min(kColTiles, pg.total_ct - cb);
            // the build's own unit decomposition (the list passes do not use units): enough column segments per row
            // tile that the 16 warps end together -- the evaluation cost per unit varies a lot
            // ((x, y) list: one unit per row tile -- the warp that owns a tile ranks its candidates row by row, build_append)
            const int Sb = SELF == 0 ? 1 : max(1, min(min(kMaxUnits / ntile, CVO_BUILD_SEGMENTS), nct / 8));
            const int nunits = ntile * Sb;
            __syncthreads();
            if (round >= kMaxListRounds) {  // uniform: every thread counts the rounds itself.  (sm.lst_ovf is only ever READ
                if (threadIdx.x == 0) sm.lst_ovf = 1;  // behind the barrier that follows the evaluation, where warps set it.)
                stop = true;
                break;
            }
            stage_tiles<STAGE_FULL>(sm, cols, cb * kTile, nct, col_tf, kColSentinel, tma_phase);
            if (threadIdx.x == 0) sm.next_unit = 0;
            if (SELF == 0)
                for (int i = threadIdx.x; i < ntile * kTile; i += kThreads) sm.u.of.bu.rowCnt[i] = 0;
            __syncthreads();
            CVO_PHASE(6)
            int wcur = 0;  // entries this warp has staged in this round
            while (warp < kWorkWarps) {  // evaluate
                const int u = next_unit(sm);
                if (u >= nunits) break;
                const int t = u / Sb, sg = u - t * Sb;
                const int c_begin = (nct * sg) / Sb, c_end = (nct * (sg + 1)) / Sb;
                int c;
                if (SELF == 0 && from_wide) {
                    uint2 rec = make_uint2(0u, 0u);
                    if (lane == 0) rec = __ldcg(lr.wide + round * kColTiles + t);
                    rec.x = __shfl_sync(0xffffffffu, rec.x, 0);
                    rec.y = __shfl_sync(0xffffffffu, rec.y, 0);
                    c = filter_unit_write(sm, ws, kp, L, rows, pg.t_begin + rb + t, (uint32_t)(t * kTile),
                                          lr.wide + kWideTable + rec.x, (int)rec.y, stage + wcur, seg - wcur);
                } else {
                    const int w0 = wo.cursor;
                    wo.out = lr.wide + kWideTable + (size_t)warp * wseg + w0;
                    wo.limit = wseg - w0;
                    wo.cursor = 0;
                    c = build_unit_write<SELF>(sm, ws, kp, L, rows, row_tf, pg.t_begin + rb + t, (uint32_t)(t * kTile),
                                               yy_row_min, c_begin, c_end, stage + wcur, seg - wcur, thr_build, wo);
                    if (make_wide && lane == 0) {
                        __stcg(lr.wide + round * kColTiles + t, make_uint2((unsigned)(warp * wseg + w0), (unsigned)wo.cursor));
                        if (w0 + wo.cursor > wseg) sm.wide_ovf = 1;
                    }
                    wo.cursor = min(w0 + wo.cursor, wseg);
                }
                if (lane == 0) {
                    sm.u.of.bu.off[u] = warp * seg + wcur;
                    sm.u.of.bu.act[u] = c;
                    if (wcur + c > seg) sm.lst_ovf = 1;
                }
                wcur = min(wcur + c, seg);
            }
            CVO_PHASE(20)  // instrumented variant: warp 0's units; what follows is the wait for the slowest warp
            __syncthreads();
            CVO_PHASE(7)
            if (SELF == 0) {  // the (x, y) list: row-sorted quads (see compact_quads)
                if (!compact_quads<SELF>(sm, lr, kind, round, ntile, Sb)) {
                    stop = true;
                    break;
                }
                continue;
            }
            if (warp == 0) {  // places in the flat list: exclusive scan of the unit counts in unit order
                int base = 0;
                for (int i0 = 0; i0 < nunits; i0 += 32) {
                    const int c = (i0 + lane < nunits) ? sm.u.of.bu.act[i0 + lane] : 0;
                    int excl, total;
                    warp_scan_count(c, lane, excl, total);
                    if (i0 + lane < nunits) sm.u.of.bu.pos[i0 + lane] = base + excl;
                    base += total;
                }
                const int padded = (base + kListTrip - 1) / kListTrip * kListTrip;
                const int at = sm.lst_used;
                const bool fits = !sm.lst_ovf && (unsigned)(at + padded) <= lr.cap;
                if (fits)  // padding: t_c = +inf (pair list) / d2 = 1e30, colour d2 = +inf (self lists) => a = 0, finite terms
                    for (int i = base + lane; i < padded; i += 32)
                        __stcg(lr.entries + at + i, make_uint2(SELF ? __float_as_uint(1.0e30f) : 0u, 0x7f800000u));
                __syncwarp();
                if (lane == 0) {
                    if (fits) {
                        sm.lround[kind][round] = make_uint2((unsigned)at, (unsigned)padded);
                        sm.lst_base = at;
                        sm.lst_used = at + padded;
                    } else {
                        sm.lst_ovf = 1;
                    }
                    sm.next_unit = 0;
                }
            }
            __syncthreads();
            if (sm.lst_ovf) {
                stop = true;
                break;
            }
            CVO_PHASE(8)
            while (true) {  // compaction: staging segments -> flat list
                const int u = next_unit(sm);
                if (u >= nunits) break;
                const uint2* src = lr.staging + sm.u.of.bu.off[u];
                uint2* dst = lr.entries + sm.lst_base + sm.u.of.bu.pos[u];
                const int c = sm.u.of.bu.act[u];
                int i = lane;
                for (; i + 224 < c; i += 256) {  // eight loads in flight per lane
                    uint2 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __ldcg(src + i + 32 * j);
#pragma unroll
                    for (int j = 0; j < 8; ++j) __stcg(dst + i + 32 * j, v[j]);
                }
                for (; i + 96 < c; i += 128) {
                    const uint2 v0 = __ldcg(src + i), v1 = __ldcg(src + i + 32), v2 = __ldcg(src + i + 64), v3 = __ldcg(src + i + 96);
                    __stcg(dst + i, v0); __stcg(dst + i + 32, v1); __stcg(dst + i + 64, v2); __stcg(dst + i + 96, v3);
                }
                for (; i < c; i += 32) __stcg(dst + i, __ldcg(src + i));
            }
        }
    }
    __syncthreads();  // the list (global memory) is complete and visible to the whole CTA
    CVO_PHASE(9)
    if (threadIdx.x == 0) {
        sm.lst[kind].valid = sm.lst_ovf ? -1 : 1;
        if (make_wide) sm.wide.valid = sm.wide_ovf ? 0 : 1;  // (an overflowed wide list is simply not used: sweeps go on)
    }
    __syncthreads();
}

// STEP pass over a list right after a pass that staged the same transformed columns: only the step-size terms are new.
__device__ __forceinline__ void stage_step_terms(Smem& sm, int n) {
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const StepCol sc = step_col(sm.ic, plane_of(sm.colG, 0)[i], plane_of(sm.colG, 1)[i], plane_of(sm.colG, 2)[i]);
        plane_of(sm.colG, 3)[i] = sc.ecn;
        float* z1 = plane_of(sm.u.ls.ss.colZ1, 0);
        float* z2 = plane_of(sm.u.ls.ss.colZ2, 0);
        z1[i] = sc.z1x; z1[i + kColChunk] = sc.z1y; z1[i + 2 * kColChunk] = sc.z1z; z1[i + 3 * kColChunk] = sc.nrm;
        z2[i] = sc.z2x; z2[i + kColChunk] = sc.z2y; z2[i + 2 * kColChunk] = sc.z2z; z2[i + 3 * kColChunk] = sc.pdt;
    }
}

// Stages this CTA's rows [first, first + n) of a packed cloud for a pass over a list: geometry only (transformed
// if the rows are the moving cloud), the original index in the w lane; rows past the cloud's end are far away.
__device__ __forceinline__ void stage_rows(Smem& sm, const CloudDev& c, int first, int n, bool tf) {
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const int p = first + i;
        float4 g = make_float4(kRowSentinel, kRowSentinel, kRowSentinel, __int_as_float(-1));
        if (p < c.n) {
            float4 q = __ldg(c.g + p);
            if (tf) apply_tf(sm.ic.tf, q.x, q.y, q.z);
            if (finite3(q.x, q.y, q.z)) g = q;  // (see stage_tiles)
        }
        plane_of(sm.u.ls.rowG, 0)[i] = g.x; plane_of(sm.u.ls.rowG, 1)[i] = g.y; plane_of(sm.u.ls.rowG, 2)[i] = g.z;
    }
}

// Narrows a valid list in place after ell has shrunk (list_policy, need == 2): the new list -- every pair within
// r_e + s of the CURRENT pose and length-scale -- is a filter of the old one, so one streaming pass over the old
// entries replaces the all-pairs sweep of a rebuild.  Per round: every warp filters a contiguous range of trips into the
// same range of the staging area (order kept, so the list stays a pure function of the inputs), the 16 counts are
// scanned, the ranges are copied back behind one another and the round is padded to a whole trip.  The narrowed
// rounds only ever move towards the front of the list area, behind the read position.
//   SELF == 0: entries (row, col, t_c); the distance is measured on the staged rows / transformed columns.
//   SELF != 0: entries (d2, colour d2 | Q1 flag): d2 is pose-independent, nothing is staged.
template <int SELF>
__device__ __noinline__ void refine_list(Smem& sm, const KParams& kp, const CloudDev& rows, const CloudDev& cols, int rank, int G,
                            uint32_t& tma_phase, int kind, const ListRef& lr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PassGeom pg = pass_geom(rows.n, cols.n, rank, G);
    const ListState& L = sm.lst[kind];
    const float t_lim = kp.t_lim, inv_c1 = L.inv_c1, s_build = L.s_build, c2 = kp.c2;
    if (threadIdx.x == 0) sm.lst_used = 0;
    int round = 0;
    for (int rb = 0; rb < pg.my_tiles; rb += pg.tiles_per_round) {
        const int ntile = min(pg.tiles_per_round, pg.my_tiles - rb);
        for (int cb = 0; cb < pg.total_ct; cb += kColTiles, ++round) {
            const int nct = min(kColTiles, pg.total_ct - cb);
            __syncthreads();
            if (SELF == 0) {  // the stages of a list pass (run_pass_list finds them afterwards)
                const int row_first = (pg.t_begin + rb) * kTile, col_first = cb * kTile;
                const bool have_cols = tag_is(sm.colTag, cols.g, col_first, nct * kTile, sm.serial);
                const bool have_rows = tag_is(sm.rowTag, rows.g, row_first, ntile * kTile, -1);
                if (!have_cols) stage_tiles<STAGE_GEOM>(sm, cols, col_first, nct, true, kColSentinel, tma_phase);
                if (!have_rows) stage_rows(sm, rows, row_first, ntile * kTile, false);
                __syncthreads();
                if (threadIdx.x == 0) {
                    sm.colTag.g = cols.g; sm.colTag.first = col_first; sm.colTag.n = nct * kTile; sm.colTag.serial = sm.serial;
                    sm.rowTag.g = rows.g; sm.rowTag.first = row_first; sm.rowTag.n = ntile * kTile; sm.rowTag.serial = -1;
                }
            }
            const uint2 rd = sm.lround[kind][round];
            const int ntrip = (int)rd.y / kListTrip;
            const int t_begin = (ntrip * warp) / kWarps, t_end = (ntrip * (warp + 1)) / kWarps;
            const uint2* src = lr.entries + rd.x + lane;
            uint2* const dst = lr.staging + (size_t)t_begin * kListTrip;
            int cursor = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const uint2* q = src + (size_t)t * kListTrip;
                uint2 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = __ldcg(q + j * kTile);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float d2, t_c;
                    if (SELF == 0) {
                        const uint32_t rowb = v[j].x >> 16, colb = v[j].x & 0xffffu;
                        d2 = dist2(plane_ld<0>(sm.colG, colb) - plane_ld<0>(sm.u.ls.rowG, rowb),
                                   plane_ld<1>(sm.colG, colb) - plane_ld<1>(sm.u.ls.rowG, rowb),
                                   plane_ld<2>(sm.colG, colb) - plane_ld<2>(sm.u.ls.rowG, rowb));
                        t_c = __uint_as_float(v[j].y);
                    } else {
                        d2 = __uint_as_float(v[j].x);
                        t_c = __fmul_rn(__uint_as_float(v[j].y & 0x7fffffffu), c2);
                    }
                    const float re2 = (t_lim - t_c) * inv_c1;  // (padding: t_c = +inf, never kept)
                    const float lim = sqrtf_approx(fmaxf(re2, 0.f)) * 1.000002f + s_build;
                    const bool keep = (re2 > 0.f) && (d2 < lim * lim * 1.000001f);
                    const uint32_t b = __ballot_sync(0xffffffffu, keep);
                    if (keep) __stcg(dst + cursor + __popc(b & ((1u << lane) - 1u)), v[j]);
                    cursor += __popc(b);
                }
            }
            if (lane == 0) sm.refineCnt[warp] = cursor;
            __syncthreads();
            if (warp == 0) {
                const int c = lane < kWarps ? sm.refineCnt[lane] : 0;
                int excl, total;
                warp_scan_count(c, lane, excl, total);
                if (lane < kWarps) sm.refinePos[lane] = excl;
                const int padded = (total + kListTrip - 1) / kListTrip * kListTrip;
                const int at = sm.lst_used;
                for (int i = total + lane; i < padded; i += 32)
                    __stcg(lr.entries + at + i, make_uint2(SELF ? __float_as_uint(1.0e30f) : 0u, 0x7f800000u));
                __syncwarp();
                if (lane == 0) {
                    sm.lround[kind][round] = make_uint2((unsigned)at, (unsigned)padded);
                    sm.lst_base = at;
                    sm.lst_used = at + padded;
                }
            }
            __syncthreads();
            {
                const uint2* from = dst;
                uint2* to = lr.entries + sm.lst_base + sm.refinePos[warp];
                int i = lane;
                for (; i + 96 < cursor; i += 128) {
                    const uint2 v0 = __ldcg(from + i), v1 = __ldcg(from + i + 32), v2 = __ldcg(from + i + 64), v3 = __ldcg(from + i + 96);
                    __stcg(to + i, v0); __stcg(to + i + 32, v1); __stcg(to + i + 64, v2); __stcg(to + i + 96, v3);
                }
                for (; i < cursor; i += 32) __stcg(to + i, __ldcg(from + i));
            }
        }
    }
    __syncthreads();
}

// One all-pairs pass over a valid neighbour list.  Rows and columns of the round are staged once; the round's flat
// list is then dealt to the warps a trip (kListTrip entries: four coalesced 8-byte loads per lane, the next trip's
// loads in flight behind this trip's arithmetic) at a time, round-robin, so every warp does the same amount of
// branch-free work and there is no per-unit overhead.  Per-lane f32 partials are promoted to f64 every few trips
// (src/cvo.cpp:197-203); the warp totals are summed in warp order: bit-deterministic.
template <int KIND>
__device__ void run_pass_list(Smem& sm, const KParams& kp, const CloudDev& rows, bool row_tf, const CloudDev& cols,
                              bool col_tf, int rank, int G, int yy_row_min, uint32_t& tma_phase, int kind, const ListRef& lr) {
    constexpr int NV = PassTraits<KIND>::NV;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PassGeom pg = pass_geom(rows.n, cols.n, rank, G);
    ListSrc src;
    src.rows = &rows;
    src.cols = &cols;
    const HotConsts hc = hot_consts(sm.ic);
    FlowPartial fp;
    fp.po0 = fp.po1 = fp.po2 = fp.pv0 = fp.pv1 = fp.pv2 = fp.psum = fp.pdl = 0.f;
    fp.cnt = 0;
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    int round = 0;
    for (int rb = 0; rb < pg.my_tiles; rb += pg.tiles_per_round) {
        const int ntile = min(pg.tiles_per_round, pg.my_tiles - rb);
        for (int cb = 0; cb < pg.total_ct; cb += kColTiles, ++round) {
            const int nct = min(kColTiles, pg.total_ct - cb);
            __syncthreads();  // everyone is done with the previous round's stage (and its tags are written)
            // Stage only what is not there already: the fixed cloud's rows survive from pass to pass and from
            // iteration to iteration, and the STEP pass finds the columns the FLOW pass transformed.
            const int row_first = (pg.t_begin + rb) * kTile, col_first = cb * kTile;
            const int row_serial = row_tf ? sm.serial : -1, col_serial = col_tf ? sm.serial : -1;
            const bool have_cols = tag_is(sm.colTag, cols.g, col_first, nct * kTile, col_serial);
            const bool have_rows = tag_is(sm.rowTag, rows.g, row_first, ntile * kTile, row_serial);
            if (KIND == PASS_STEP) {
                if (have_cols) stage_step_terms(sm, nct * kTile);
                else stage_tiles<STAGE_STEP>(sm, cols, col_first, nct, col_tf, kColSentinel, tma_phase);
            } else if (!have_cols) {
                stage_tiles<STAGE_GEOM>(sm, cols, col_first, nct, col_tf, kColSentinel, tma_phase);
            }
            if (!have_rows) stage_rows(sm, rows, row_first, ntile * kTile, row_tf);
            __syncthreads();
            if (threadIdx.x == 0) {  // read again only after the next barrier
                sm.colTag.g = cols.g; sm.colTag.first = col_first; sm.colTag.n = nct * kTile; sm.colTag.serial = col_serial;
                sm.rowTag.g = rows.g; sm.rowTag.first = row_first; sm.rowTag.n = ntile * kTile; sm.rowTag.serial = row_serial;
            }
            src.row_base = row_first;
            src.col_base = col_first;
            const uint2 rd = sm.lround[kind][round];
            const int ntrip = (int)rd.y / kListTrip;
            const uint2* e0 = lr.entries + rd.x;
            const uint2* e = e0 + lane;
            // Register sets (A, B, C) rotate between "being processed" and "being loaded" without ever being copied: a
            // copy would have to wait for the load it copies.  Loads past the warp's last trip are clamped to it
            // (always readable, never a branch).
            uint2 a0, a1, a2, a3, b0, b1, b2, b3;
            // The lists stream from HBM (they are larger than this SM's share of L2): an L2 prefetch a few trips
            // ahead (8 lines of 128 B per trip, one per lane 0..7) leaves the register loads only L2 latency to cover.
#define CVO_LOAD_TRIP(x0, x1, x2, x3, tt)                                                        \
    {                                                                                            \
        const uint2* q = e + (size_t)min((tt), ntrip - 1) * kListTrip;                           \
        x0 = __ldcg(q); x1 = __ldcg(q + kTile); x2 = __ldcg(q + 2 * kTile); x3 = __ldcg(q + 3 * kTile); \
        if (lane < 8) {                                                                          \
            const uint2* f = e0 + (size_t)min((tt) + kPrefetchTrips * kWarps, ntrip - 1) * kListTrip + lane * 16; \
            asm volatile("prefetch.global.L2 [%0];" ::"l"(f));                                   \
        }                                                                                        \
    }
#define CVO_RUN_TRIP(x0, x1, x2, x3)                                                             \
    {                                                                                            \
        const bool n0 = list_body<KIND, false>(sm, hc, kp, x0.x, __uint_as_float(x0.y), yy_row_min, src, fp, acc); \
        const bool n1 = list_body<KIND, false>(sm, hc, kp, x1.x, __uint_as_float(x1.y), yy_row_min, src, fp, acc); \
        const bool n2 = list_body<KIND, false>(sm, hc, kp, x2.x, __uint_as_float(x2.y), yy_row_min, src, fp, acc); \
        const bool n3 = list_body<KIND, false>(sm, hc, kp, x3.x, __uint_as_float(x3.y), yy_row_min, src, fp, acc); \
        if (__any_sync(0xffffffffu, n0 | n1 | n2 | n3)) { /* about one trip in ten thousand */    \
            if (n0) list_body<KIND, true>(sm, hc, kp, x0.x, 0.f, yy_row_min, src, fp, acc);       \
            if (n1) list_body<KIND, true>(sm, hc, kp, x1.x, 0.f, yy_row_min, src, fp, acc);       \
            if (n2) list_body<KIND, true>(sm, hc, kp, x2.x, 0.f, yy_row_min, src, fp, acc);       \
            if (n3) list_body<KIND, true>(sm, hc, kp, x3.x, 0.f, yy_row_min, src, fp, acc);       \
        }                                                                                        \
        if (KIND == PASS_STEP) flush_partial<KIND>(fp, acc);                                     \
    }
            int t = warp;
            if (t < ntrip) {
#if CVO_LIST_SETS == 3
                // three register sets: the loads run TWO trips ahead of the arithmetic (a trip of interleaved bodies is
                // shorter than the latency of a list line that misses L2)
                uint2 c0, c1, c2, c3;
                CVO_LOAD_TRIP(a0, a1, a2, a3, t)
                CVO_LOAD_TRIP(b0, b1, b2, b3, t + kWarps)
#pragma unroll 1
                while (true) {
                    CVO_LOAD_TRIP(c0, c1, c2, c3, t + 2 * kWarps)
                    CVO_RUN_TRIP(a0, a1, a2, a3)
                    t += kWarps;
                    if (t >= ntrip) break;
                    CVO_LOAD_TRIP(a0, a1, a2, a3, t + 2 * kWarps)
                    CVO_RUN_TRIP(b0, b1, b2, b3)
                    if (KIND != PASS_STEP) flush_partial<KIND>(fp, acc);  // <= 12 terms per f32 partial, like a short row of A
                    t += kWarps;
                    if (t >= ntrip) break;
                    CVO_LOAD_TRIP(b0, b1, b2, b3, t + 2 * kWarps)
                    CVO_RUN_TRIP(c0, c1, c2, c3)
                    t += kWarps;
                    if (t >= ntrip) break;
                }
#else
                CVO_LOAD_TRIP(a0, a1, a2, a3, t)
#pragma unroll 1
                while (true) {
                    CVO_LOAD_TRIP(b0, b1, b2, b3, t + kWarps)
                    CVO_RUN_TRIP(a0, a1, a2, a3)
                    t += kWarps;
                    if (t >= ntrip) break;
                    CVO_LOAD_TRIP(a0, a1, a2, a3, t + kWarps)
                    CVO_RUN_TRIP(b0, b1, b2, b3)
                    if (KIND != PASS_STEP) flush_partial<KIND>(fp, acc);  // <= 8 terms per f32 partial, like a short row of A
                    t += kWarps;
                    if (t >= ntrip) break;
                }
#endif
            }
#undef CVO_LOAD_TRIP
#undef CVO_RUN_TRIP
            flush_partial<KIND>(fp, acc);
        }
    }
    warp_sum_multi<NV>(acc, lane);
    if (NV >= 8) {
        if ((lane & 3) == 0) sm.u.ls.warpTot[warp][multi_value_index(lane)] = acc[0];
        if (lane == 0) {
#pragma unroll
            for (int i = 8; i < NV; ++i) sm.u.ls.warpTot[warp][i] = acc[i];
        }
    } else if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sm.u.ls.warpTot[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < kNumAcc) {  // fixed-order sum over the warps
        double t = 0.0;
        if (threadIdx.x < NV)
            for (int w = 0; w < kWarps; ++w) t += sm.u.ls.warpTot[w][threadIdx.x];
        sm.blockTot[threadIdx.x] = t;
    }
    __syncthreads();
}

// A pass over the (x, x) or (y, y) list of acvo (src/adaptive_cvo.cpp:159-160,205-231,243-265): the entries are the
// pose-independent pairs (d2, colour d2), so nothing is staged and no point is touched -- the list streams through the
// warps (same trips, same two register sets as run_pass_list) and every entry costs a dozen instructions.
//   acc[0] = nnz, acc[1] = sum a * d2 / ell^3 (for (y, y): only the rows quirk Q1 lets through)
template <bool EXACT>
__device__ __forceinline__ bool self_body(const IterConsts& ic, const KParams& kp, float c1, float d2_thres, float inv_ell3,
                                          uint32_t d2_bits, uint32_t d2c_bits, float& pdl, int& cnt) {
    const float d2 = __uint_as_float(d2_bits);
    const float d2c = __uint_as_float(d2c_bits & 0x7fffffffu);
    const bool q1 = (d2c_bits >> 31) != 0;
    bool near = false;
    float a;
    if (EXACT) {
        a = kernel_value_exact_d(ic.ell, ic.d2c_thres, kp.s2, kp.cs2, kp.c_ell, kp.sp_thres, d2c, d2);
    } else {
        HotConsts h;  // only c1 is read by kernel_a
        h.c1 = c1;
        a = kernel_a(h, kp, d2, __fmul_rn(d2c, kp.c2), near);
    }
    const bool ok = !near && (a > kp.sp_thres) && (d2 < d2_thres);
    a = ok ? a : 0.f;
    pdl = fmaf(inv_ell3 * (q1 ? a : 0.f), d2, pdl);  // src/adaptive_cvo.cpp:210,231 / :256,259
    cnt += ok ? 1 : 0;
    return near;
}

template <int KIND>  // PASS_XX or PASS_YY
__device__ void run_pass_self(Smem& sm, const KParams& kp, const CloudDev& rows, const CloudDev& cols, int rank, int G, int kind,
                              const ListRef& lr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PassGeom pg = pass_geom(rows.n, cols.n, rank, G);
    const float c1 = sm.ic.c1, d2_thres = sm.ic.d2_thres, inv_ell3 = sm.ic.inv_ell3;
    float pdl = 0.f;
    int cnt = 0;
    double acc[2] = {0.0, 0.0};
    int round = 0;
    for (int rb = 0; rb < pg.my_tiles; rb += pg.tiles_per_round)
        for (int cb = 0; cb < pg.total_ct; cb += kColTiles, ++round) {
            const uint2 rd = sm.lround[kind][round];
            const int ntrip = (int)rd.y / kListTrip;
            const uint2* e0 = lr.entries + rd.x;
            const uint2* e = e0 + lane;
            uint2 a0, a1, a2, a3, b0, b1, b2, b3;
#define CVO_LOAD_TRIP(x0, x1, x2, x3, tt)                                                        \
    {                                                                                            \
        const uint2* q = e + (size_t)min((tt), ntrip - 1) * kListTrip;                           \
        x0 = __ldcg(q); x1 = __ldcg(q + kTile); x2 = __ldcg(q + 2 * kTile); x3 = __ldcg(q + 3 * kTile); \
        if (lane < 8) {                                                                          \
            const uint2* f = e0 + (size_t)min((tt) + kPrefetchTrips * kWarps, ntrip - 1) * kListTrip + lane * 16; \
            asm volatile("prefetch.global.L2 [%0];" ::"l"(f));                                   \
        }                                                                                        \
    }
#define CVO_RUN_TRIP(x0, x1, x2, x3)                                                             \
    {                                                                                            \
        const bool n0 = self_body<false>(sm.ic, kp, c1, d2_thres, inv_ell3, x0.x, x0.y, pdl, cnt); \
        const bool n1 = self_body<false>(sm.ic, kp, c1, d2_thres, inv_ell3, x1.x, x1.y, pdl, cnt); \
        const bool n2 = self_body<false>(sm.ic, kp, c1, d2_thres, inv_ell3, x2.x, x2.y, pdl, cnt); \
        const bool n3 = self_body<false>(sm.ic, kp, c1, d2_thres, inv_ell3, x3.x, x3.y, pdl, cnt); \
        if (__any_sync(0xffffffffu, n0 | n1 | n2 | n3)) {                                        \
            if (n0) self_body<true>(sm.ic, kp, c1, d2_thres, inv_ell3, x0.x, x0.y, pdl, cnt);    \
            if (n1) self_body<true>(sm.ic, kp, c1, d2_thres, inv_ell3, x1.x, x1.y, pdl, cnt);    \
            if (n2) self_body<true>(sm.ic, kp, c1, d2_thres, inv_ell3, x2.x, x2.y, pdl, cnt);    \
            if (n3) self_body<true>(sm.ic, kp, c1, d2_thres, inv_ell3, x3.x, x3.y, pdl, cnt);    \
        }                                                                                        \
        acc[1] += (double)pdl;  /* <= 4 terms per f32 partial */                                 \
        pdl = 0.f;                                                                               \
    }
            int t = warp;
            if (t < ntrip) {
#if CVO_SELF_SETS == 3  // three register sets, loads two trips ahead (see run_pass_list)
                uint2 g0, g1, g2, g3;
                CVO_LOAD_TRIP(a0, a1, a2, a3, t)
                CVO_LOAD_TRIP(b0, b1, b2, b3, t + kWarps)
#pragma unroll 1
                while (true) {
                    CVO_LOAD_TRIP(g0, g1, g2, g3, t + 2 * kWarps)
                    CVO_RUN_TRIP(a0, a1, a2, a3)
                    t += kWarps;
                    if (t >= ntrip) break;
                    CVO_LOAD_TRIP(a0, a1, a2, a3, t + 2 * kWarps)
                    CVO_RUN_TRIP(b0, b1, b2, b3)
                    t += kWarps;
                    if (t >= ntrip) break;
                    CVO_LOAD_TRIP(b0, b1, b2, b3, t + 2 * kWarps)
                    CVO_RUN_TRIP(g0, g1, g2, g3)
                    t += kWarps;
                    if (t >= ntrip) break;
                }
#else
                CVO_LOAD_TRIP(a0, a1, a2, a3, t)
#pragma unroll 1
                while (true) {
                    CVO_LOAD_TRIP(b0, b1, b2, b3, t + kWarps)
                    CVO_RUN_TRIP(a0, a1, a2, a3)
                    t += kWarps;
                    if (t >= ntrip) break;
                    CVO_LOAD_TRIP(a0, a1, a2, a3, t + kWarps)
                    CVO_RUN_TRIP(b0, b1, b2, b3)
                    t += kWarps;
                    if (t >= ntrip) break;
                }
#endif
            }
#undef CVO_LOAD_TRIP
#undef CVO_RUN_TRIP
        }
    acc[0] = (double)cnt;
    __syncthreads();  // the previous pass is done with the warp totals
    acc[0] = warp_sum(acc[0]);
    acc[1] = warp_sum(acc[1]);
    if (lane == 0) {
        sm.u.ls.warpTot[warp][0] = acc[0];
        sm.u.ls.warpTot[warp][1] = acc[1];
    }
    __syncthreads();
    if (threadIdx.x < kNumAcc) {  // fixed-order sum over the warps
        double t = 0.0;
        if (threadIdx.x < 2)
            for (int w = 0; w < kWarps; ++w) t += sm.u.ls.warpTot[w][threadIdx.x];
        sm.blockTot[threadIdx.x] = t;
    }
    __syncthreads();
}

#include "cvo_quads.cuh"

// All-gather of the per-CTA totals through distributed shared memory; every CTA of the cluster ends with
// identical cluster totals in sm.sum[dst_off ...] (summed in rank order).
template <int NV>
__device__ __forceinline__ void cluster_allreduce(Smem& sm, cg::cluster_group& cluster, const double* src, int buf,
                                                  int dst_off) {
    const int rank = (int)cluster.block_rank(), G = (int)cluster.num_blocks();
    if (G == 1) {  // one CTA per pair (the batch benchmark): the CTA's totals ARE the pair's totals (0.0 + x == x)
        if (threadIdx.x < NV) sm.sum[dst_off + threadIdx.x] = src[threadIdx.x];
        __syncthreads();
        return;
    }
    if (threadIdx.x < NV) {
        const double v = src[threadIdx.x];
        for (int r = 0; r < G; ++r) {
            double* dst = cluster.map_shared_rank(&sm.xchg[buf][rank][threadIdx.x], r);
            *dst = v;
        }
    }
    cluster.sync();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (int r = 0; r < G; ++r) t += sm.xchg[buf][r][threadIdx.x];
        sm.sum[dst_off + threadIdx.x] = t;
    }
    __syncthreads();
}

// Whole-GPU mode: the H clusters of the launch work on one pair.  After the cluster all-reduce every CTA holds its
// cluster's totals in sm.sum[dst_off ...]; rank 0 of every cluster publishes them in global memory and arrives on a
// counter, every CTA waits for the H arrivals of this round and sums the H records in cluster order -- identical
// totals everywhere, bit-deterministic.  Two record buffers alternate: a cluster can only reach round r + 2 after
// every CTA has contributed to round r + 1, i.e. has finished reading round r.  All CTAs of the launch are resident
// (the host launches no more clusters than the device holds), so the wait cannot deadlock.
template <int NV>
__device__ __forceinline__ void group_allreduce(Smem& sm, const AlignArgs& args, int H, int cid, int crank, int dst_off,
                                                unsigned& round) {
    if (H <= 1) return;
    round += 1;
    double* rec = args.group_xchg + (size_t)(round & 1u) * kMaxGroupClusters * kNumAcc;
    if (crank == 0 && threadIdx.x < NV) __stcg(rec + cid * kNumAcc + threadIdx.x, sm.sum[dst_off + threadIdx.x]);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (crank == 0) {
            __threadfence();  // the CTA's records (ordered before this thread by the barrier) before the arrival
            atomicAdd(args.group_count, 1u);
        }
        const unsigned target = round * (unsigned)H;
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(args.group_count) : "memory");
        } while (seen < target);
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (int c = 0; c < H; ++c) t += __ldcg(rec + c * kNumAcc + threadIdx.x);
        sm.sum[dst_off + threadIdx.x] = t;
    }
    __syncthreads();
}

// --------------------------------------------------------------------------------------------
// the persistent align kernel
// --------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads, 1) align_kernel(const AlignArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    // rank / G: this CTA's place among the CTAs that share a pair -- the cluster, or in whole-GPU mode all H clusters
    const int crank = (int)cluster.block_rank(), cG = (int)cluster.num_blocks();
    const int H = args.group_clusters > 1 ? args.group_clusters : 1, cid = (int)blockIdx.x / cG;
    const int rank = H > 1 ? cid * cG + crank : crank, G = H * cG;
    unsigned group_round = 0;
    const KParams& kp = args.kp;
    const bool acvo = kp.mode == CVO_B200_MODE_ACVO;
    const int max_iter = kp.fixed_iters > 0 ? kp.fixed_iters : kp.max_iter;
    uint32_t tma_phase = 0;  // parity of sm.tma_bar; every thread tracks it (all threads stage every chunk)
    if (threadIdx.x == 0) {
        mbar_init(&sm.tma_bar, 1);
        sm.serial = 0;
    }
    __syncthreads();
    const bool use_lists = args.list_entries != nullptr;
    ListRef lref[LIST_KINDS];
#pragma unroll
    for (int i = 0; i < LIST_KINDS; ++i) {
        const size_t area = (size_t)blockIdx.x * kListAreas;
        lref[i].entries = args.list_entries + (area + i) * args.list_cap;
        lref[i].staging = args.list_entries + (area + LIST_KINDS) * args.list_cap;
        lref[i].wide = args.list_entries + (area + LIST_KINDS + 1) * args.list_cap;
        lref[i].cap = args.list_cap;
    }

    cluster.sync();  // every CTA of the cluster runs before anyone writes into its shared memory

    for (int group_pi = 0;; ++group_pi) {
        if (H == 1) {  // a free cluster pulls the next pair
            if (crank == 0 && threadIdx.x == 0) {
                const int idx = atomicAdd(args.counter, 1);
                for (int r = 0; r < cG; ++r) *cluster.map_shared_rank(&sm.next_pair, r) = idx;
            }
            cluster.sync();
        }
        const int pi = H == 1 ? sm.next_pair : group_pi;  // whole-GPU mode: every cluster takes every pair, in order
        if (pi >= args.n_pairs) break;
        const PairDev pair = args.pairs[pi];
        if (threadIdx.x == 0) {
            sm.st = args.states[pi];
            sm.st.iters = max_iter;
            sm.st.status = CVO_B200_STATUS_MAX_ITER;
            sm.st.n_run = 0;
            sm.st.n_builds = 0;
            sm.st.n_refines = 0;
            sm.st.xy_entries = sm.st.xy_slots = 0;
            sm.done = 0;
            sm.ic.ell = -1.f;  // the constants cached by length-scale (prepare_iter) belong to the previous pair
            sm.colTag.serial = sm.rowTag.serial = -2;
#pragma unroll
            for (int i = 0; i < LIST_KINDS; ++i) sm.lst[i].valid = sm.lst[i].need = 0;
            sm.wide.valid = sm.wide.make = 0;
        }
        __syncthreads();
        if (use_lists) {  // bounding box of the moving cloud (original coordinates), for list_policy
            const float inf = __int_as_float(0x7f800000);
            float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
            for (int i = threadIdx.x; i < pair.y.n; i += kThreads) {
                const float4 g = __ldg(pair.y.g + i);
                if (!finite3(g.x, g.y, g.z)) continue;  // such a point is never staged as is
                lo[0] = fminf(lo[0], g.x); lo[1] = fminf(lo[1], g.y); lo[2] = fminf(lo[2], g.z);
                hi[0] = fmaxf(hi[0], g.x); hi[1] = fmaxf(hi[1], g.y); hi[2] = fmaxf(hi[2], g.z);
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float l = warp_min(lo[a]), h = warp_max(hi[a]);
                if ((threadIdx.x & 31) == 0) {
                    sm.wred[threadIdx.x >> 5][a] = l;
                    sm.wred[threadIdx.x >> 5][3 + a] = h;
                }
            }
            __syncthreads();
            if (threadIdx.x < 6) {
                float v = sm.wred[0][threadIdx.x];
                for (int w = 1; w < kWarps; ++w)
                    v = threadIdx.x < 3 ? fminf(v, sm.wred[w][threadIdx.x]) : fmaxf(v, sm.wred[w][threadIdx.x]);
                sm.ybox[threadIdx.x] = v;
            }
            __syncthreads();
        }

        if (threadIdx.x < 32) {  // iteration 0: update_tf (src/cvo.cpp:368) + which lists to build
            if (threadIdx.x == 0) {
                sm.serial += 1;
                prepare_iter(sm, kp, kp.d2c_thres);
            }
            __syncwarp();
            if (use_lists) list_policy(sm, acvo, args.list_skin, args.list_skin_min, args.list_shrink, args.list_refine_min, args.list_wide);
        }
        __syncthreads();
#ifdef CVO_PHASE_CLOCKS
        if (threadIdx.x == 0) sm.phase_t0 = clock64();
#endif
        for (int k = 0; k < max_iter; ++k) {
            // transform_pcd + se_kernel + compute_flow (src/cvo.cpp:371-374)
            CVO_PHASE(0)
            if (use_lists && sm.lst[LIST_XY].need == 1) build_list<0>(sm, kp, pair.x, false, pair.y, true, rank, G, 0, tma_phase, LIST_XY, lref[LIST_XY]);
            else if (use_lists && sm.lst[LIST_XY].need == 3) build_list<0>(sm, kp, pair.x, false, pair.y, true, rank, G, 0, tma_phase, LIST_XY, lref[LIST_XY], true);
            if (use_lists && acvo) {  // all builds come before the first pass: the stages the FLOW pass fills survive to the STEP pass
                if (sm.lst[LIST_XX].need == 1) build_list<1>(sm, kp, pair.x, false, pair.x, false, rank, G, 0, tma_phase, LIST_XX, lref[LIST_XX]);
                else if (sm.lst[LIST_XX].need == 2) refine_list<1>(sm, kp, pair.x, pair.x, rank, G, tma_phase, LIST_XX, lref[LIST_XX]);
                if (sm.lst[LIST_YY].need == 1) build_list<2>(sm, kp, pair.y, true, pair.y, true, rank, G, pair.x.n, tma_phase, LIST_YY, lref[LIST_YY]);
                else if (sm.lst[LIST_YY].need == 2) refine_list<2>(sm, kp, pair.y, pair.y, rank, G, tma_phase, LIST_YY, lref[LIST_YY]);
            }
            CVO_PHASE(1)
            const bool list_xy = use_lists && sm.lst[LIST_XY].valid > 0;
            // nnz(A) / sum(A) are only observable through the trace of pair 0 (and, for acvo, through dl)
            const bool stats = args.trace != nullptr && pi == 0;
            if (list_xy && acvo) run_pass_quads<PASS_FLOW, true>(sm, kp, pair.x, pair.y, rank, G, tma_phase, lref[LIST_XY]);
            else if (list_xy && stats) run_pass_quads<PASS_FLOW_CVO, true>(sm, kp, pair.x, pair.y, rank, G, tma_phase, lref[LIST_XY]);
            else if (list_xy) run_pass_quads<PASS_FLOW_CVO, false>(sm, kp, pair.x, pair.y, rank, G, tma_phase, lref[LIST_XY]);
            else run_pass<PASS_FLOW>(sm, kp, pair.x, false, pair.y, true, rank, G, 0, tma_phase);
            CVO_PHASE(2)
            if (threadIdx.x < ACC_FLOW_COUNT) sm.flowTot[threadIdx.x] = threadIdx.x < 9 ? sm.blockTot[threadIdx.x] : 0.0;
            if (acvo) {  // Axx, Ayy (src/adaptive_cvo.cpp:159-160)
                if (use_lists && sm.lst[LIST_XX].valid > 0)
                    run_pass_self<PASS_XX>(sm, kp, pair.x, pair.x, rank, G, LIST_XX, lref[LIST_XX]);
                else run_pass<PASS_XX>(sm, kp, pair.x, false, pair.x, false, rank, G, 0, tma_phase);
                if (threadIdx.x < 2) sm.flowTot[ACC_NNZXX + threadIdx.x] = sm.blockTot[threadIdx.x];
                if (use_lists && sm.lst[LIST_YY].valid > 0)
                    run_pass_self<PASS_YY>(sm, kp, pair.y, pair.y, rank, G, LIST_YY, lref[LIST_YY]);
                else run_pass<PASS_YY>(sm, kp, pair.y, true, pair.y, true, rank, G, pair.x.n, tma_phase);
                if (threadIdx.x < 2) sm.flowTot[ACC_NNZYY + threadIdx.x] = sm.blockTot[threadIdx.x];
            }
            __syncthreads();
            cluster_allreduce<ACC_FLOW_COUNT>(sm, cluster, sm.flowTot, 0, kFlowOff);
            group_allreduce<ACC_FLOW_COUNT>(sm, args, H, cid, crank, kFlowOff, group_round);
            if (threadIdx.x == 0) finalize_flow(sm);
            __syncthreads();
            CVO_PHASE(3)
            // compute_step_size (src/cvo.cpp:377)
            if (list_xy) run_pass_quads<PASS_STEP, false>(sm, kp, pair.x, pair.y, rank, G, tma_phase, lref[LIST_XY]);
            else run_pass<PASS_STEP>(sm, kp, pair.x, false, pair.y, true, rank, G, 0, tma_phase);
            CVO_PHASE(4)
            cluster_allreduce<4>(sm, cluster, sm.blockTot, 1, 0);
            group_allreduce<4>(sm, args, H, cid, crank, 0, group_round);
            CVO_PHASE(10)
            if (threadIdx.x < 32) {  // the serial section of the iteration, on warp 0 (its parallel parts use the lanes)
                // remember the transform used by this iteration: it is what the reference multiplies
                // into accum_transform when the loop exits here (quirk Q3, src/cvo.cpp:413-414)
                if (threadIdx.x == 0) write_tf44(sm.ic.tf, sm.st.prev_tf);
                cvo_b200_iter_rec* rec = nullptr;
                if (args.trace && pi == 0 && rank == 0 && k < args.trace_cap) rec = args.trace + k;
                update_state(sm, kp, k, rec);
                __syncwarp();
                CVO_PHASE(11)
                if (!sm.done && k + 1 < max_iter) {  // the next iteration's update_tf + list decisions, same serial section
                    if (threadIdx.x == 0) {
                        sm.serial += 1;
                        prepare_iter(sm, kp, kp.d2c_thres);
                    }
                    __syncwarp();
                    CVO_PHASE(12)
                    if (use_lists) list_policy(sm, acvo, args.list_skin, args.list_skin_min, args.list_shrink, args.list_refine_min, args.list_wide);
                    CVO_PHASE(13)
                }
            }
            __syncthreads();
            CVO_PHASE(5)
            if (sm.done) break;
        }
        if (threadIdx.x == 0 && rank == 0) {
            prepare_iter(sm, kp, kp.d2c_thres);  // final update_tf(), src/cvo.cpp:415
            write_tf44(sm.ic.tf, sm.st.tf);
            args.states[pi] = sm.st;
        }
        __syncthreads();
    }
}

// acvo::function_inner_product (src/adaptive_cvo.cpp:385-439): untransformed clouds, colour gate from sp_thres.
__global__ void __launch_bounds__(kThreads, 1) inner_product_kernel(const InnerArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), G = (int)cluster.num_blocks();
    uint32_t tma_phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&sm.tma_bar, 1);
#pragma unroll
        for (int i = 0; i < 9; ++i) sm.st.R[i] = (i % 4 == 0) ? 1.f : 0.f;
        sm.st.T[0] = sm.st.T[1] = sm.st.T[2] = 0.f;
        sm.st.ell = args.ell;
        sm.ic.ell = -1.f;
        prepare_iter(sm, args.kp, args.kp.d2c_thres);
    }
    cluster.sync();  // (also: every CTA of the cluster runs before the all-reduce writes into its shared memory)
    run_pass<PASS_INNER>(sm, args.kp, args.pair.x, false, args.pair.y, false, rank, G, 0, tma_phase);
    cluster_allreduce<2>(sm, cluster, sm.blockTot, 0, 0);
    if (rank == 0 && threadIdx.x == 0) {
        args.out[0] = sm.sum[0];
        args.out[1] = sm.sum[1];
    }
}

// --------------------------------------------------------------------------------------------
// upload-time packing: Morton sort + 32-byte rows  (replaces the tail of set_pcd, src/cvo.cpp:343-356)
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

constexpr int kPackThreads = 1024;

// One CTA per cloud: bounding box -> 30-bit Morton key -> bitonic sort of (key, index) in shared
// memory -> gather into {x,y,z,f0} / {f1..f4} rows.  Ties break on the original index, so the
// packed order is a pure function of the input.
__global__ void __launch_bounds__(kPackThreads, 1) pack_sort_kernel(const PackJob* jobs, int sort_points) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    __shared__ float sbox[6][32];
    __shared__ float bb[6];
    const PackJob job = jobs[blockIdx.x];
    const int n = job.n;
    if (n <= 0) return;
    int npad = 1;
    while (npad < n) npad <<= 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int i = threadIdx.x; i < n; i += kPackThreads) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = job.xyz[3 * i + a];
            if (!isfinite(v)) continue;  // keeps the Morton grid of the finite points intact
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float l = warp_min(lo[a]), h = warp_max(hi[a]);
        if (lane == 0) {
            sbox[a][warp] = l;
            sbox[3 + a][warp] = h;
        }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float l = warp_min(sbox[a][lane]), h = warp_max(sbox[3 + a][lane]);
            if (lane == 0) {
                bb[a] = l;
                bb[3 + a] = h;
            }
        }
    }
    __syncthreads();
    float scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float ext = bb[3 + a] - bb[a];
        scale[a] = ext > 0.f ? 1023.999f / ext : 0.f;
    }
    for (int i = threadIdx.x; i < npad; i += kPackThreads) {
        unsigned long long key = ~0ull;
        if (i < n) {
            uint32_t code = 0;
            if (sort_points) {
                const uint32_t qx = (uint32_t)((job.xyz[3 * i + 0] - bb[0]) * scale[0]);
                const uint32_t qy = (uint32_t)((job.xyz[3 * i + 1] - bb[1]) * scale[1]);
                const uint32_t qz = (uint32_t)((job.xyz[3 * i + 2] - bb[2]) * scale[2]);
                code = spread10(qx) | (spread10(qy) << 1) | (spread10(qz) << 2);
            }
            key = ((unsigned long long)code << 32) | (unsigned long long)(uint32_t)i;
        }
        keys[i] = key;
    }
    __syncthreads();
    if (sort_points) {
        for (int k = 2; k <= npad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (npad >> 1); t += kPackThreads) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int p = i | j;
                    const unsigned long long a = keys[i], b = keys[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        keys[i] = b;
                        keys[p] = a;
                    }
                }
                __syncthreads();
            }
        }
    }
    for (int i = threadIdx.x; i < n; i += kPackThreads) {
        const int src = (int)(uint32_t)(keys[i] & 0xffffffffull);
        const float* p = job.xyz + 3 * src;
        const float* f = job.feat + 5 * src;
        job.out_g[i] = make_float4(p[0], p[1], p[2], __int_as_float(src));
        job.out_f[i] = make_float4(f[0], f[1], f[2], f[3]);
        job.out_f4[i] = f[4];
    }
}

}  // namespace cvo_b200

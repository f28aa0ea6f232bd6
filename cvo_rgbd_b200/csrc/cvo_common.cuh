// cvo_common.cuh -- tuning switches, constants, device-side structures (clouds, pair state, shared-memory layout, launch arguments) and small helpers
// (included by cvo_kernels.cuh inside namespace cvo_b200; see that file for the overall design)
#pragma once


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
constexpr int kMaxColChunks = 6;    // column chunks of a pass: 6 x 3072 = 18 432 points (> the 16 384 a context can hold)
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
    // per column chunk of the (x, y) list: the column tiles its sweep found live in ANY of the row rounds that share the chunk
    // (the quad passes stage a chunk once for all of them, and only these tiles)
    uint32_t colMask[kMaxColChunks][kColTiles / 32];
    uint2 lround[LIST_KINDS][kMaxListRounds];  // (offset, entries) of every round of a list; entries % kListTrip == 0
    int lst_base;
    int refineCnt[kWarps], refinePos[kWarps];  // refine_list: entries every warp kept / where they go
    StageTag colTag, rowTag;  // what the column / row stages of the list passes currently hold
    int serial;               // running iteration number of this CTA: identifies "transformed with this iteration's pose"
    float wred[kWarps][6];    // per-warp partial bounding boxes (pair start)
    float ybox[6];            // bounding box of the moving cloud, original coordinates
    ListState lst[LIST_KINDS];
    WideState wide;
    float tf_prev[12];  // the transform of the previous iteration (IterConsts::tf layout): the direction the pose is moving in
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
    float list_ahead;     // build the (x, y) list this many skins ahead of the motion (0 = at the current pose)
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

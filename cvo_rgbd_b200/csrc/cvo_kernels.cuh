// cvo_kernels.cuh -- sm_100a kernels for the RKHS SE(3) registration inner loop.
//
// One persistent kernel (align_kernel, below) runs the reference's whole align() loop
// (cpp/rkhs_registration/src/cvo.cpp:361-420, src/adaptive_cvo.cpp:490-555) on the device: a thread-block CLUSTER of
// G CTAs -- or, in whole-GPU mode, every cluster of the launch -- owns one frame pair at a time and iterates
//     transform_pcd -> se_kernel -> compute_flow -> compute_step_size -> Exp_SEK3 update -> stop tests -> ell policy
// without host round trips.  The sparse affinity matrix A (inc/cvo.hpp:92) is never stored.  What replaces the
// reference's kd-tree (thirdparty/nanoflann.hpp, rebuilt twice per se_kernel call) is a per-pair NEIGHBOUR CANDIDATE
// LIST kept across iterations in HBM scratch: index pairs plus the pose-independent colour exponent, built by an
// all-pairs sweep over Morton-sorted tiles (bounding-box culling, expanded-form prefilter, exact evaluation) with a
// skin, valid while the pose and the length-scale stay inside it; both passes of an iteration (flow, step
// coefficients) stream the list and evaluate the strict ell-ball test and the kernel value from the freshly transformed
// geometry.  Cross-CTA reductions go through distributed shared memory + cluster barriers (and global memory between
// clusters in whole-GPU mode); every sum has a fixed order: results are bit-deterministic.
//
// Files (all included here, inside namespace cvo_b200, in this order):
//   cvo_common.cuh    tuning switches, constants, structures (clouds, pair state, shared-memory layout, arguments), helpers
//   cvo_epilogue.cuh  the scalar part of an iteration: update_tf, thresholds, line search, Exp_SEK3, stop tests, ell policies
//   cvo_stage.cuh     tile staging: TMA bulk copies (features), transformed column geometry
//   cvo_onthefly.cuh  exact kernel values / gates; the on-the-fly all-pairs passes (fallback, inner product)
//   cvo_lists.cuh     list validity policy, sweep, wide list + filter, narrowing in place, self-list passes of acvo
//   cvo_quads.cuh     the (x, y) list as row-sorted quads: compaction and the two hot passes (packed f32x2)
//   (this file)       cluster / whole-GPU reductions, align_kernel, inner_product_kernel
//   cvo_pack.cuh      upload-time Morton sort + packing (pack_sort_kernel)
// DESIGN.md section 3 describes the design and its measurements.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cvo_b200.h"

namespace cvo_b200 {

#include "cvo_common.cuh"
#include "cvo_epilogue.cuh"
#include "cvo_stage.cuh"
#include "cvo_onthefly.cuh"
#include "cvo_lists.cuh"
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
            if (threadIdx.x < 12) sm.tf_prev[threadIdx.x] = sm.ic.tf[threadIdx.x];  // iteration 0: no motion known yet
            __syncwarp();
            if (use_lists) list_policy(sm, acvo, args.list_skin, args.list_skin_min, args.list_shrink, args.list_refine_min, args.list_wide, args.list_ahead);
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
                if (threadIdx.x < 12) sm.tf_prev[threadIdx.x] = sm.ic.tf[threadIdx.x];
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
                    if (use_lists) list_policy(sm, acvo, args.list_skin, args.list_skin_min, args.list_shrink, args.list_refine_min, args.list_wide, args.list_ahead);
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

#include "cvo_pack.cuh"

}  // namespace cvo_b200

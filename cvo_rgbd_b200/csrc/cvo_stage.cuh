// cvo_stage.cuh -- tile staging: TMA bulk copies of the feature planes, transformed column geometry, per-column step terms
// (included by cvo_kernels.cuh inside namespace cvo_b200; see that file for the overall design)
#pragma once

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
                                            float sentinel, uint32_t& tma_phase, const float* tf_override = nullptr,
                                            const uint32_t* tile_mask = nullptr) {
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
    const float* tf12 = tf_override ? tf_override : sm.ic.tf;  // (a list sweep stages the columns at the list's own pose)
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
        // (tile_mask: a quad pass only stages the column tiles its list references -- a CTA that holds a few row tiles of
        // a large pair touches a small part of the moving cloud; warp-uniform: a warp stages whole tiles)
        if (tile_mask && !((tile_mask[i >> 10] >> ((i >> 5) & 31)) & 1u)) continue;
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
        if (tile_mask && !((tile_mask[i >> 10] >> ((i >> 5) & 31)) & 1u)) continue;
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

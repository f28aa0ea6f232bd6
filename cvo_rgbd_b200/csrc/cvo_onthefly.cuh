// cvo_onthefly.cuh -- kernel values and gates in the reference's arithmetic; the on-the-fly all-pairs passes (tile culling, prefilter, queues, survivor body)
// (included by cvo_kernels.cuh inside namespace cvo_b200; see that file for the overall design)
#pragma once

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
    int row_tile0, row_stride;  // the round's l-th staged row tile is tile row_tile0 + l * row_stride of the cloud
    int col_base;               // global index of the staged chunk's column 0
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
// How the row tiles are dealt to the CTAs that share a pair.  ROUND-ROBIN (tile t -> CTA t mod G) when the moving cloud
// fits one column chunk: Morton neighbours have similar candidate densities, so every CTA gets its share of the dense and
// of the sparse regions (3000 points on 16 CTAs: -5 % against contiguous ranges, whose slowest CTA keeps the others
// waiting at the cluster barriers).  CONTIGUOUS ranges when there are several column chunks: a CTA's rows then
// reference few column tiles, and the quad passes stage only those (Smem::colMask) -- 10 000 points on 112 CTAs: 43.6 us
// per iteration against 52.0 round-robin (profiles/r02_group_mode.txt).
struct PassGeom {
    int t_begin, t_stride;  // this CTA's l-th row tile is tile t_begin + l * t_stride of the cloud
    int my_tiles, total_ct, S, tiles_per_round;
};
__device__ __forceinline__ PassGeom pass_geom(int rows_n, int cols_n, int rank, int G) {
    PassGeom pg;
    const int total_rt = (rows_n + kTile - 1) / kTile;
    pg.total_ct = (cols_n + kTile - 1) / kTile;
    if (pg.total_ct <= kColTiles) {  // the moving cloud is one column chunk: balance wins
        pg.t_begin = rank;
        pg.t_stride = G;
        pg.my_tiles = rank < total_rt ? (total_rt - rank + G - 1) / G : 0;
    } else {  // several column chunks: locality wins (see above)
        pg.t_begin = (total_rt * rank) / G;
        pg.t_stride = 1;
        pg.my_tiles = (total_rt * (rank + 1)) / G - pg.t_begin;
    }
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
    const int t_begin = pg.t_begin, t_stride = pg.t_stride, my_tiles = pg.my_tiles, total_ct = pg.total_ct, S = pg.S;
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
                process_unit<KIND>(sm, kp, rows, row_tf, t_begin + (rb + t) * t_stride, c_begin, c_end, u, cb == 0, yy_row_min);
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

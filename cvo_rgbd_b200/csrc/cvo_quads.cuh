// cvo_quads.cuh -- the (x, y) neighbour list as ROW-SORTED QUADS and the two passes over it (included by
// cvo_kernels.cuh inside namespace cvo_b200).
//
// Layout.  The candidates of one round (row chunk x column chunk) are sorted by row; every row's run is padded to a
// multiple of four, so the list is a sequence of QUADS: four candidates of ONE row.  A quad is stored as three
// parallel arrays (structure of arrays, 26 B per quad = 6.5 B per candidate; 8 B before):
//     ntc  float4   the four NEGATED colour exponents -t_c (padding: -inf  =>  a = 0)
//     cols 4 x u16  shared-memory byte addresses of the four columns' x inside the staged column planes
//     row  u16      shared-memory byte address of the row's x inside the staged row planes
// (addresses relative to the CTA's shared-memory window: the operands of the hot loop are LDS [field + window + plane
// offset], see lds_at)
// and a round is padded to a whole TRIP of 32 quads: lane l of a warp takes quad l of the trip with one LDG.128, one
// LDG.64 and one LDG.U16, all three coalesced.
//
// Why rows.  With d = y_j - x_i (diff_yx, src/cvo.cpp:192) the flow of a row factors:
//     sum_j a_ij (x_i x y_j) = x_i x sum_j a_ij d_ij        (x_i x x_i = 0)        src/cvo.cpp:191,197
//     sum_j a_ij (y_j - x_i) =       sum_j a_ij d_ij                                 src/cvo.cpp:192,198
// so a candidate costs three FMAs (a d) instead of a cross product and six, and the cross product is taken once per
// quad -- the reference's own order of operations (per-row f32 sums `Ai*cross_xy`, `Ai*diff_yx`, then f64 across rows).
// The step-size terms (src/cvo.cpp:226-279) factor the same way: with the ROW vectors u_i = omega x x_i + v and
// w_i = omega x u_i,
//     xi z_j       = z1 = u_i + omega x d                 z1 . r = -u_i . d                        (r = x_i - y_j = -d)
//     xi^2 z_j     = z2 = w_i + omega (omega . d) - |omega|^2 d
//     S           := |omega x d|^2 = |omega|^2 |d|^2 - (omega . d)^2
//     z2 . r       = S - w_i . d                          |z1|^2 = |u_i|^2 + S - 2 w_i . d
//     |z2|^2       = |w_i|^2 + |omega|^2 (S - 2 w_i . d)
//     -z1 . z2     = 0                                    (z2 = omega x z1: the reference's xiz_dot_xi2z, :236, is pure
//                                                          f32 rounding noise around 0; its effect on D is < 1e-9 relative)
//     |z2|^2 + 2 z1 . z3 = -|z2|^2                        (z3 = omega x z2: the reference's epsil_const, :237)
//   with Q := 3 S - 4 w_i . d:
//     beta = 2t u_i . d     gamma = -t (|u_i|^2 + Q)     delta = 2t (omega . v)(omega . d) - |omega|^2 beta
//     epsil = t (|w_i|^2 + |omega|^2 Q)                                                            (t = 1 / (2 l^2))
//   (the rows are staged as 2t u_i and -4 w_i, so beta and Q's last term are plain dot products with d)
// Three dot products per candidate (u.d, w.d, omega.d), nothing per COLUMN: the eight per-column planes the STEP pass
// used to stage and gather (nine shared loads per candidate) are gone; eight per-ROW terms are staged instead and read
// once per quad.
//
// Packed arithmetic.  A lane holds its quad as two PAIRS of candidates and evaluates them with the packed f32x2
// instructions of sm_100 (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.rn.f32x2): one issue slot per two candidates for all of
// the geometry, the kernel exponent and the polynomial.  The kernel is issue-bound, so this is
// where the time goes down; every operation is the same correctly rounded f32 operation as before.
//
// Gates.  a > sp_thres implies the strict ell-ball test d2 < d2_thres whenever c_sigma^2 <= 1 (k = a / ck >= a) and a
// is not within the re-decision band of sp_thres, so the fast path tests only a; candidates inside the band (about one
// in a million) are re-decided per candidate by kernel_value_exact + the exact ball test, as before.  (With
// c_sigma > 1 -- neither reference class -- the host keeps the passes on the fly.)
#pragma once

namespace quads {

constexpr int kQuadTrip = 32;            // quads per trip (one per lane)
constexpr int kQuadBytes = 26;           // 16 (ntc) + 8 (cols) + 2 (row)
#ifdef CVO_TRIPS_CONTIGUOUS              // tuning: every warp takes a contiguous sixteenth of a round's trips
constexpr int kTripStride = 1;           // (measured: cfg2 +0.4 %, stock cvo -4 %)
#else
constexpr int kTripStride = kWarps;      // the warps take the trips w, w + 16, ...
#endif

__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ float hsum(float2 a) { return a.x + a.y; }

// Shared-memory operands of the hot loop.  A quad's row / column fields are the shared-memory byte addresses of the
// point's x inside the staged row / column planes, relative to the CTA's shared-memory WINDOW (written by compact_quads
// of the same kernel, so the layout can never disagree): one LDS [field + window + plane offset] per operand, the
// window base in a uniform register, no address arithmetic per operand.  (In a cluster launch the shared::cta window
// of CTA rank r starts at r << 24: a bare offset would address rank 0's shared memory.)
__device__ __forceinline__ uint32_t smem_window(const void* p) {  // (volatile: kept in a register, not re-derived per trip)
    uint32_t w;
    asm volatile("and.b32 %0, %1, 0xff000000;" : "=r"(w) : "r"(smem_u32(p)));
    return w;
}
__device__ __forceinline__ uint32_t smem_offset(const void* p) { return smem_u32(p) & 0x00ffffffu; }
template <uint32_t OFF>
__device__ __forceinline__ float lds_at(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
// plane offsets relative to a row's x: the row geometry planes, then the STEP pass's row terms (ListStage layout)
constexpr uint32_t kRowZ1 = (uint32_t)sizeof(float4) * kColChunk;      // ls.ss.colZ1 - ls.rowG
constexpr uint32_t kRowZ2 = 2u * (uint32_t)sizeof(float4) * kColChunk;  // ls.ss.colZ2 - ls.rowG
static_assert(offsetof(ListStage, ss) == kRowZ1 && offsetof(StepStage, colZ2) == kRowZ1, "row-term planes follow the row planes");

struct Round {  // one round's arrays inside the CTA's list area
    const float4* ntc;
    const uint2* cols;
    const unsigned short* row;
    int ntrip;
};
__device__ __forceinline__ Round round_ref(const ListRef& lr, uint2 rd) {
    const char* base = reinterpret_cast<const char*>(lr.entries) + (size_t)rd.x * 16u;
    Round r;
    r.ntc = reinterpret_cast<const float4*>(base);
    r.cols = reinterpret_cast<const uint2*>(base + (size_t)rd.y * 16u);
    r.row = reinterpret_cast<const unsigned short*>(base + (size_t)rd.y * 24u);
    r.ntrip = (int)rd.y / kQuadTrip;
    return r;
}

struct Quad {
    float4 ntc;
    uint2 cols;
    uint32_t row;
};
// Quad `qi` of the round (lane l of a warp: its trip's first quad + l) and, with the same address registers, the L2
// prefetch of the lines kPrefetchTrips of the warp's trips further down the round (one instruction per array with an
// immediate offset: no address arithmetic, no predicate; the eight lanes of a line coalesce into one request).  Neither the loads
// two trips ahead nor the prefetches are clamped to the round: the list scratch ends in a slack region that keeps them
// inside mapped memory (kListSlackBytes), and what is loaded past the warp's share is never used.
__device__ __forceinline__ Quad load_quad(const Round& r, uint32_t qi, bool pf_lane) {
#ifdef CVO_EXP_SAMETRIP  // timing experiment only (wrong results): every trip re-reads the round's first trips (cache hits)
    qi &= 511u;
#endif
#ifdef CVO_CLAMP_LOADS
    qi = min(qi, (uint32_t)(r.ntrip * kQuadTrip - 1));
#endif
    Quad v;
    const float4* pn = r.ntc + qi;
    const uint2* pc = r.cols + qi;
    const unsigned short* pr = r.row + qi;
#ifdef CVO_LIST_EVICT_FIRST
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.ntc.x), "=f"(v.ntc.y), "=f"(v.ntc.z), "=f"(v.ntc.w) : "l"(pn), "l"(pol));
    asm volatile("ld.global.cg.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.cols.x), "=r"(v.cols.y) : "l"(pc), "l"(pol));
#else
    v.ntc = __ldcg(pn);
    v.cols = __ldcg(pc);
#endif
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=r"(v.row) : "l"(pr));  // (zero-extended into the 32-bit register)
#ifndef CVO_NO_LIST_PREFETCH
    if (pf_lane) {  // (all lanes: the eight lanes of a line coalesce)
        asm volatile("prefetch.global.L2 [%0+%1];" ::"l"(pn), "n"(kPrefetchTrips * kTripStride * kQuadTrip * 16));
        asm volatile("prefetch.global.L2 [%0+%1];" ::"l"(pc), "n"(kPrefetchTrips * kTripStride * kQuadTrip * 8));
        asm volatile("prefetch.global.L2 [%0+%1];" ::"l"(pr), "n"(kPrefetchTrips * kTripStride * kQuadTrip * 2));
    }
#endif
    return v;
}

// Geometry and gated kernel values of one quad.
struct QuadGeom {
    float xr, yr, zr;              // the row
    float2 dxa, dya, dza, d2a;     // candidates 0, 1: d = y - x, |d|^2 (nanoflann's accumulation order, see dist2)
    float2 dxb, dyb, dzb, d2b;     // candidates 2, 3
    float2 aa, ab;                 // a (0 where a gate failed)
    uint32_t rw;                   // the row's shared-memory address, window included
};

// `near`: some candidate of the quad has its fast kernel value inside the re-decision band around sp_thres.
__device__ __forceinline__ bool quad_geom(const HotConsts& hc, const KParams& kp, uint32_t win, const Quad& q, QuadGeom& g) {
    constexpr uint32_t P = kPlaneBytes;
    // field | window in one byte permute each: bytes {f0, f1, w2, w3}
    const uint32_t rw = g.rw = __byte_perm(q.row, win, 0x7610);
    g.xr = lds_at<0>(rw); g.yr = lds_at<P>(rw); g.zr = lds_at<2 * P>(rw);
    const uint32_t c0 = __byte_perm(q.cols.x, win, 0x7610), c1 = __byte_perm(q.cols.x, win, 0x7632);
    const uint32_t c2 = __byte_perm(q.cols.y, win, 0x7610), c3 = __byte_perm(q.cols.y, win, 0x7632);
    g.dxa = __fadd2_rn(make_float2(lds_at<0>(c0), lds_at<0>(c1)), bc(-g.xr));
    g.dya = __fadd2_rn(make_float2(lds_at<P>(c0), lds_at<P>(c1)), bc(-g.yr));
    g.dza = __fadd2_rn(make_float2(lds_at<2 * P>(c0), lds_at<2 * P>(c1)), bc(-g.zr));
    g.dxb = __fadd2_rn(make_float2(lds_at<0>(c2), lds_at<0>(c3)), bc(-g.xr));
    g.dyb = __fadd2_rn(make_float2(lds_at<P>(c2), lds_at<P>(c3)), bc(-g.yr));
    g.dzb = __fadd2_rn(make_float2(lds_at<2 * P>(c2), lds_at<2 * P>(c3)), bc(-g.zr));
    g.d2a = __ffma2_rn(g.dza, g.dza, __ffma2_rn(g.dya, g.dya, __fmul2_rn(g.dxa, g.dxa)));
    g.d2b = __ffma2_rn(g.dzb, g.dzb, __ffma2_rn(g.dyb, g.dyb, __fmul2_rn(g.dxb, g.dxb)));
    // a = s2 c_sigma^2 2^-(d2 c1 + t_c)  (kernel_a; -(d2 c1 + t_c) = fma(d2, -c1, -t_c) bit for bit)
    const float2 ea = __ffma2_rn(g.d2a, bc(-hc.c1), make_float2(q.ntc.x, q.ntc.y));
    const float2 eb = __ffma2_rn(g.d2b, bc(-hc.c1), make_float2(q.ntc.z, q.ntc.w));
    const float2 fa = __fmul2_rn(bc(kp.s2cs2), make_float2(exp2f_approx(ea.x), exp2f_approx(ea.y)));
    const float2 fb = __fmul2_rn(bc(kp.s2cs2), make_float2(exp2f_approx(eb.x), exp2f_approx(eb.y)));
    const float2 ma = __fadd2_rn(fa, bc(-kp.sp_thres)), mb = __fadd2_rn(fb, bc(-kp.sp_thres));  // a - sp_thres: same sign as (a > sp_thres)
    const float m = fminf(fminf(fabsf(ma.x), fabsf(ma.y)), fminf(fabsf(mb.x), fabsf(mb.y)));
    g.aa = make_float2(ma.x > 0.f ? fa.x : 0.f, ma.y > 0.f ? fa.y : 0.f);  // src/cvo.cpp:152
    g.ab = make_float2(mb.x > 0.f ? fb.x : 0.f, mb.y > 0.f ? fb.y : 0.f);
    return m < kp.sp_band;
}

// The re-decision of a quad that has a candidate inside the band: every candidate's value and gates in the reference's
// own arithmetic (kernel_value_exact: features from global memory) and the exact strict ball test.  About one trip in
// ten thousand gets here; not inlined so that it costs the hot loop no registers.
__device__ __noinline__ float quad_exact1(const Smem& sm, const KParams& kp, const ListSrc& src, uint32_t rowb, uint32_t colb, float d2) {
    const IterConsts& ic = sm.ic;
    const int rl = (int)((rowb - quads::smem_offset(sm.u.ls.rowG)) >> 2);  // row within the round's staged rows
    const int ri = (src.row_tile0 + (rl >> 5) * src.row_stride) * kTile + (rl & 31), ci = src.col_base + (int)((colb - quads::smem_offset(sm.colG)) >> 2);
    float a = kernel_value_exact(ic.ell, ic.d2c_thres, kp.s2, kp.cs2, kp.c_ell, kp.sp_thres, __ldg(src.rows->f + ri), __ldg(src.rows->f4 + ri),
                                 __ldg(src.cols->f + ci), __ldg(src.cols->f4 + ci), d2);
    if (!(d2 < ic.d2_thres)) a = 0.f;  // thirdparty/nanoflann.hpp:249-253
    return a;
}
// Called (warp-uniformly) when some lane of the warp saw a candidate inside the band: every lane re-decides the candidates
// of ITS quad that sit inside the band.  Padding candidates (-t_c = -inf, a = 0) are far outside it.
__device__ __forceinline__ void redecide(const Smem& sm, const HotConsts& hc, const KParams& kp, const ListSrc& src, const Quad& q, QuadGeom& g) {
    const float nt[4] = {q.ntc.x, q.ntc.y, q.ntc.z, q.ntc.w};
    const float d2[4] = {g.d2a.x, g.d2a.y, g.d2b.x, g.d2b.y};
    const uint32_t cb[4] = {q.cols.x & 0xffffu, q.cols.x >> 16, q.cols.y & 0xffffu, q.cols.y >> 16};
    float a[4] = {g.aa.x, g.aa.y, g.ab.x, g.ab.y};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float d2e = d2[e];
        asm volatile("" : "+f"(d2e));  // pins this arithmetic to the cold path (the compiler hoisted it above the vote otherwise)
        const float fast = __fmul_rn(kp.s2cs2, exp2f_approx(fmaf(d2e, -hc.c1, nt[e])));
        if (fabsf(fast - kp.sp_thres) < kp.sp_band) a[e] = quad_exact1(sm, kp, src, q.row, cb[e], d2[e]);
    }
    g.aa = make_float2(a[0], a[1]);
    g.ab = make_float2(a[2], a[3]);
}

// ---- FLOW (src/cvo.cpp:164-210; acvo: + the (x, y) term of the length-scale gradient, src/adaptive_cvo.cpp:202,228) ----
// STATS: nnz(A) and sum(A) are wanted (acvo always: nnz enters dl; cvo only for the trace / eval hook).
template <int KIND, bool STATS>
__device__ __forceinline__ void flow_quad(const HotConsts& hc, const KParams& kp, const QuadGeom& g, FlowPartial& fp) {
    const float sx = hsum(__ffma2_rn(g.ab, g.dxb, __fmul2_rn(g.aa, g.dxa)));  // sum_j a_ij d_ij over the quad, f32 like
    const float sy = hsum(__ffma2_rn(g.ab, g.dyb, __fmul2_rn(g.aa, g.dya)));  // the reference's per-row Ai*diff_yx
    const float sz = hsum(__ffma2_rn(g.ab, g.dzb, __fmul2_rn(g.aa, g.dza)));
    fp.po0 = fmaf(kp.inv_c, g.yr * sz - g.zr * sy, fp.po0);  // (1/c) x_i x sum a d
    fp.po1 = fmaf(kp.inv_c, g.zr * sx - g.xr * sz, fp.po1);
    fp.po2 = fmaf(kp.inv_c, g.xr * sy - g.yr * sx, fp.po2);
    fp.pv0 = fmaf(kp.inv_d, sx, fp.pv0);
    fp.pv1 = fmaf(kp.inv_d, sy, fp.pv1);
    fp.pv2 = fmaf(kp.inv_d, sz, fp.pv2);
    if (KIND == PASS_FLOW) fp.pdl = fmaf(hc.inv_ell3, hsum(__ffma2_rn(g.ab, g.d2b, __fmul2_rn(g.aa, g.d2a))), fp.pdl);
    if (STATS) {
        fp.psum += hsum(__fadd2_rn(g.aa, g.ab));
        fp.cnt += (g.aa.x > 0.f) + (g.aa.y > 0.f) + (g.ab.x > 0.f) + (g.ab.y > 0.f);
    }
}
// per-lane f32 partials (a few quads) -> f64 (src/cvo.cpp:202-203)
template <int KIND, bool STATS>
__device__ __forceinline__ void flush_flow(FlowPartial& fp, double* acc) {
    acc[ACC_W0] += (double)fp.po0; acc[ACC_W0 + 1] += (double)fp.po1; acc[ACC_W0 + 2] += (double)fp.po2;
    acc[ACC_V0] += (double)fp.pv0; acc[ACC_V0 + 1] += (double)fp.pv1; acc[ACC_V0 + 2] += (double)fp.pv2;
    fp.po0 = fp.po1 = fp.po2 = fp.pv0 = fp.pv1 = fp.pv2 = 0.f;
    if (KIND == PASS_FLOW) {
        acc[ACC_DLXY] += (double)fp.pdl;
        fp.pdl = 0.f;
    }
    if (STATS) {
        acc[ACC_SUMA] += (double)fp.psum;
        acc[ACC_NNZ] += (double)fp.cnt;
        fp.psum = 0.f;
        fp.cnt = 0;
    }
}

// ---- STEP (src/cvo.cpp:213-289): this quad's terms of B, C, D, E ----
struct StepRow {  // the row's staged terms, see stage_row_step_terms
    float bx, by, bz;   // 2t u_i           (beta  = b_i . d)
    float qx, qy, qz;   // -4 w_i           (Q     = 3 S + q_i . d)
    float guu, ew2;     // -t |u_i|^2, t |w_i|^2
};
struct StepConsts {  // per-iteration scalars of the row-factored form
    float w0, w1, w2;   // omega
    float ww3;          // 3 |omega|^2      Q = 3 S - 4 w.d = ww3 |d|^2 - 3 (omega.d)^2 + q_i . d
    float mt;           // -t               gamma = guu + mt Q
    float dk1, nww;     // 2t (omega.v), -|omega|^2      delta = dk1 omega.d + nww beta
    float etw;          // t |omega|^2      epsil = ew2 + etw Q
};
__device__ __forceinline__ StepConsts step_consts(const HotConsts& hc) {
    StepConsts s;
    s.w0 = hc.omega[0]; s.w1 = hc.omega[1]; s.w2 = hc.omega[2];
    const float ww = (s.w0 * s.w0 + s.w1 * s.w1) + s.w2 * s.w2;
    const float wv = (s.w0 * hc.v[0] + s.w1 * hc.v[1]) + s.w2 * hc.v[2];
    s.ww3 = 3.f * ww;
    s.mt = -hc.temp_coef;
    s.dk1 = hc.p2t * wv;
    s.nww = -ww;
    s.etw = hc.temp_coef * ww;
    return s;
}
// One pair of candidates: beta .. epsil (src/cvo.cpp:260-271) from three dot products, then this pair's UNWEIGHTED
// polynomial terms (the brackets of :275-279); the caller weights them with a and sums.
//   C:  gamma + beta^2/2 =: c
//   D:  delta + beta gamma + beta^3/6        = delta + beta (gamma + beta^2/6)
//   E:  epsil + beta delta + beta^2 gamma/2 + gamma^2/2 + beta^4/24  = epsil + beta delta + c^2/2 - beta^4/12
__device__ __forceinline__ void step_pair(const StepConsts& sc, const StepRow& r, float2 dx, float2 dy, float2 dz, float2 d2,
                                          float2& beta, float2& tC, float2& tD, float2& tE) {
    beta = __ffma2_rn(bc(r.bz), dz, __ffma2_rn(bc(r.by), dy, __fmul2_rn(bc(r.bx), dx)));                 // :262
    const float2 qd = __ffma2_rn(bc(r.qz), dz, __ffma2_rn(bc(r.qy), dy, __fmul2_rn(bc(r.qx), dx)));      // -4 w_i . d
    const float2 po = __ffma2_rn(bc(sc.w2), dz, __ffma2_rn(bc(sc.w1), dy, __fmul2_rn(bc(sc.w0), dx)));  // omega . d
    const float2 Q = __ffma2_rn(__fmul2_rn(bc(-3.f), po), po, __ffma2_rn(bc(sc.ww3), d2, qd));           // 3 |omega x d|^2 - 4 w.d
    const float2 gamma = __ffma2_rn(bc(sc.mt), Q, bc(r.guu));                                             // :264
    const float2 delta = __ffma2_rn(bc(sc.dk1), po, __fmul2_rn(bc(sc.nww), beta));                        // :267
    const float2 epsil = __ffma2_rn(bc(sc.etw), Q, bc(r.ew2));                                            // :270
    const float2 b2 = __fmul2_rn(beta, beta);
    tC = __ffma2_rn(bc(0.5f), b2, gamma);
    tD = __ffma2_rn(beta, __ffma2_rn(b2, bc(1.f / 6.f), gamma), delta);
    tE = __ffma2_rn(__fmul2_rn(b2, b2), bc(-1.f / 12.f), __ffma2_rn(__fmul2_rn(bc(0.5f), tC), tC, __ffma2_rn(beta, delta, epsil)));
}
__device__ __forceinline__ void step_quad(const StepConsts& sc, uint32_t rowb, const QuadGeom& g, FlowPartial& fp) {  // rowb: window included
    constexpr uint32_t P = kPlaneBytes;  // the step stage holds the ROW terms: planes bx, by, bz, guu | qx, qy, qz, ew2
    StepRow r;
    r.bx = lds_at<kRowZ1>(rowb); r.by = lds_at<kRowZ1 + P>(rowb); r.bz = lds_at<kRowZ1 + 2 * P>(rowb); r.guu = lds_at<kRowZ1 + 3 * P>(rowb);
    r.qx = lds_at<kRowZ2>(rowb); r.qy = lds_at<kRowZ2 + P>(rowb); r.qz = lds_at<kRowZ2 + 2 * P>(rowb); r.ew2 = lds_at<kRowZ2 + 3 * P>(rowb);
    float2 bA, cA, dA, eA, bB, cB, dB, eB;
    step_pair(sc, r, g.dxa, g.dya, g.dza, g.d2a, bA, cA, dA, eA);
    step_pair(sc, r, g.dxb, g.dyb, g.dzb, g.d2b, bB, cB, dB, eB);
    // A_ij times the brackets (:275-279); the terms of up to three quads are summed in f32 (each is an f32 value already),
    // then promoted (flush_step; the reference promotes every term)
    fp.po0 += hsum(__ffma2_rn(g.ab, bB, __fmul2_rn(g.aa, bA)));
    fp.po1 += hsum(__ffma2_rn(g.ab, cB, __fmul2_rn(g.aa, cA)));
    fp.po2 += hsum(__ffma2_rn(g.ab, dB, __fmul2_rn(g.aa, dA)));
    fp.pv0 += hsum(__ffma2_rn(g.ab, eB, __fmul2_rn(g.aa, eA)));
}
__device__ __forceinline__ void flush_step(FlowPartial& fp, double* acc) {
    acc[0] += (double)fp.po0; acc[1] += (double)fp.po1; acc[2] += (double)fp.po2; acc[3] += (double)fp.pv0;
    fp.po0 = fp.po1 = fp.po2 = fp.pv0 = 0.f;
}

}  // namespace quads

// Per-ROW step-size terms of the round's staged rows (untransformed fixed cloud): with u = omega x x + v, w = omega x u
// and t = 1 / (2 l^2) the planes hold 2t u, -t |u|^2 | -4 w, t |w|^2 (the factors the quads' polynomial needs, see
// step_pair).  Runs once per iteration, n <= kColChunk rows.
__device__ __forceinline__ void stage_row_step_terms(Smem& sm, int n) {
    const IterConsts& ic = sm.ic;
    const float w0 = ic.omega[0], w1 = ic.omega[1], w2 = ic.omega[2], v0 = ic.v[0], v1 = ic.v[1], v2 = ic.v[2];
    const float t = ic.temp_coef, t2 = ic.p2t;
    float* z1 = plane_of(sm.u.ls.ss.colZ1, 0);
    float* z2 = plane_of(sm.u.ls.ss.colZ2, 0);
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const float x = plane_of(sm.u.ls.rowG, 0)[i], y = plane_of(sm.u.ls.rowG, 1)[i], z = plane_of(sm.u.ls.rowG, 2)[i];
        const float ux = (w1 * z - w2 * y) + v0, uy = (w2 * x - w0 * z) + v1, uz = (w0 * y - w1 * x) + v2;
        const float wx = w1 * uz - w2 * uy, wy = w2 * ux - w0 * uz, wz = w0 * uy - w1 * ux;
        z1[i] = t2 * ux; z1[i + kColChunk] = t2 * uy; z1[i + 2 * kColChunk] = t2 * uz;
        z1[i + 3 * kColChunk] = -t * ((ux * ux + uy * uy) + uz * uz);
        z2[i] = -4.f * wx; z2[i + kColChunk] = -4.f * wy; z2[i + 2 * kColChunk] = -4.f * wz;
        z2[i + 3 * kColChunk] = t * ((wx * wx + wy * wy) + wz * wz);
    }
}

// Shared memory of the scatter (compact_quads): per warp kScatterQuads quads -- the ntc part (16 B per quad) in the memory
// of the column geometry / feature stages at the start of Smem, the cols + row part (10 B per quad) behind the
// compaction's counters in the memory of the queues and the warps' row tiles.
constexpr int kScatterQuads = 416;
constexpr size_t kScatterRegB = (sizeof(QuadBuild) + 15) & ~size_t(15);
static_assert(offsetof(Smem, colG) == 0 && offsetof(Smem, u) == sizeof(float4) * kColChunk && offsetof(OnTheFlyStage, fs) == 0 &&
              offsetof(FeatStage, colF) == 0, "the scatter's first region starts at the start of Smem");
static_assert((size_t)kWorkWarps * kScatterQuads * 16 <= sizeof(float4) * kColChunk + offsetof(FeatStage, queue), "scatter region A");
static_assert(offsetof(OnTheFlyStage, ws) == sizeof(FeatStage) &&
              kScatterRegB + (size_t)kWorkWarps * kScatterQuads * 10 <= sizeof(uint32_t) * kWorkWarps * kQueueCap + sizeof(WarpScratch) * kWorkWarps,
              "scatter region B");

// Row-sorted compaction of the round that build_list<0> has just evaluated: the staged candidates (column, row within
// the tile, rank within the row, t_c; one unit per row tile) become the round's quads.
//   count    done while the tiles were evaluated (build_append: the warp that owns a tile counts its rows and ranks
//            every kept candidate within its row, in evaluation order -- a pure function of the inputs);
//   place    quads per row -> exclusive scan inside the tile -> exclusive scan over the tiles -> the round's region in
//            the list area (ntc | cols | row arrays, padded to a whole trip with quads that can never pass);
//   scatter  (row, rank) fixes a candidate's quad and slot, so the staged entries are scattered independently of one
//            another by all warps, with many loads in flight.
// Returns false (and sets sm.lst_ovf) if the region does not fit the list area.
template <int SELF>
__device__ __forceinline__ bool compact_quads(Smem& sm, const ListRef& lr, int kind, int round, int ntile, int Sb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QuadBuild& qb = sm.u.of.fs.qb;
    const BuildUnits& bu = sm.u.of.bu;
    const float ninf = -__int_as_float(0x7f800000);
    // a quad addresses its operands by their absolute shared-memory byte address inside the list passes' planes (lds_at)
    const uint32_t col_addr0 = quads::smem_offset(sm.colG), row_addr0 = quads::smem_offset(sm.u.ls.rowG);
    if (row_addr0 + kPlaneBytes > 0x10000u || col_addr0 + kPlaneBytes > 0x10000u) {  // (cannot happen with this Smem layout)
        if (threadIdx.x == 0) sm.lst_ovf = 1;
        return false;
    }
    if (sm.lst_ovf) return false;  // a unit outgrew its staging segment: its tail was never stored (stable since the barrier)
    // ---- quads per row (the rows were counted while the units were evaluated, build_append) -> offsets inside the tile
    for (int t = warp; t < ntile; t += kWarps) {
        const int q = (bu.rowCnt[t * kTile + lane] + 3) >> 2;
        int excl, total;
        warp_scan_count(q, lane, excl, total);
        qb.rowQ[t * kTile + lane] = excl;
        if (lane == 0) qb.tileQ[t] = total;
    }
    __syncthreads();
    // ---- place
    if (warp == 0) {
        int base = 0;
        for (int i0 = 0; i0 < ntile; i0 += 32) {
            const int c = (i0 + lane < ntile) ? qb.tileQ[i0 + lane] : 0;
            int excl, total;
            warp_scan_count(c, lane, excl, total);
            __syncwarp();
            if (i0 + lane < ntile) qb.tileQ[i0 + lane] = base + excl;
            base += total;
        }
        if (lane == 0) qb.tileQ[ntile] = base;  // (the scatter reads a tile's quad count as a difference)
        const int nq = (base + quads::kQuadTrip - 1) / quads::kQuadTrip * quads::kQuadTrip;
        const unsigned units16 = ((unsigned)nq * quads::kQuadBytes + 15u) / 16u;  // the region, in 16-byte units
        const unsigned at = (unsigned)sm.lst_used;
        const bool fits = !sm.lst_ovf && (unsigned long long)(at + units16) * 16ull <= (unsigned long long)lr.cap * 8ull;
        if (fits) {  // the quads behind the last row: never pass (-t_c = -inf), address row 0 / column 0
            char* rb = reinterpret_cast<char*>(lr.entries) + (size_t)at * 16u;
            for (int q = base + lane; q < nq; q += 32) {
                __stcg(reinterpret_cast<float4*>(rb) + q, make_float4(ninf, ninf, ninf, ninf));
                __stcg(reinterpret_cast<uint2*>(rb + (size_t)nq * 16u) + q, make_uint2(col_addr0 * 0x10001u, col_addr0 * 0x10001u));
                reinterpret_cast<unsigned short*>(rb + (size_t)nq * 24u)[q] = (unsigned short)row_addr0;
            }
        }
        __syncwarp();
        if (lane == 0) {
            if (fits) {
                sm.lround[kind][round] = make_uint2(at, (unsigned)nq);
                sm.lst_base = (int)at;
                sm.lst_used = (int)(at + units16);
                int kept = 0;  // work accounting (cvo_b200_last_list_fill)
                for (int u = 0; u < ntile * Sb; ++u) kept += bu.act[u];
                sm.st.xy_entries += kept;
                sm.st.xy_slots += 4 * nq;
            } else {
                sm.lst_ovf = 1;
            }
        }
    }
    __syncthreads();
    CVO_PHASE(21)  // instrumented variant: count + place
    if (sm.lst_ovf) return false;
    // ---- scatter, through shared memory: (row, rank) fixes a candidate's quad and slot -- SLOT-MAJOR inside the row:
    // candidate `rank` of a row with n quads is slot rank / n of quad rank % n, so that the lanes holding consecutive
    // quads of a row read consecutive candidates (consecutive columns, different shared-memory banks) in each of their
    // four slots.  Scattering 4- and 2-byte fields straight into the list costs one 32-byte store transaction per field
    // (measured: 120 k cycles per build).  Instead every warp assembles the quads of a row tile, kScatterQuads at a time,
    // in its own piece of the shared memory the evaluation has finished with (column geometry and features, the tail
    // of the queues, the row tiles) and copies them out with coalesced 16 / 8 / 2-byte stores.
    {
        const unsigned nq = sm.lround[kind][round].y;
        char* rb = reinterpret_cast<char*>(lr.entries) + (size_t)sm.lst_base * 16u;
        float4* g_ntc = reinterpret_cast<float4*>(rb);
        uint2* g_cols = reinterpret_cast<uint2*>(rb + (size_t)nq * 16u);
        unsigned short* g_row = reinterpret_cast<unsigned short*>(rb + (size_t)nq * 24u);
        float* s_ntc = reinterpret_cast<float*>(reinterpret_cast<char*>(&sm) + (size_t)warp * kScatterQuads * 16);
        char* regB = reinterpret_cast<char*>(sm.u.of.fs.queue) + kScatterRegB + (size_t)warp * kScatterQuads * 10;
        unsigned short* s_cols = reinterpret_cast<unsigned short*>(regB);
        unsigned short* s_row = reinterpret_cast<unsigned short*>(regB + kScatterQuads * 8);
        const float4 ninf4 = make_float4(ninf, ninf, ninf, ninf);
        const uint2 col00 = make_uint2(col_addr0 * 0x10001u, col_addr0 * 0x10001u);
        for (int t = warp < kWorkWarps ? warp : ntile; t < ntile; t += kWorkWarps) {  // (the warps that own a staging piece)
            const int c = bu.act[t], tile_q = qb.tileQ[t], tile_nq = qb.tileQ[t + 1] - tile_q;
            const uint2* src = lr.staging + bu.off[t];
            for (int q0 = 0; q0 < tile_nq; q0 += kScatterQuads) {  // (one chunk for all but very dense tiles)
                const int nqc = min(kScatterQuads, tile_nq - q0);
                __syncwarp();  // the previous chunk has been copied out
                for (int i = lane; i < nqc; i += 32) {  // unused slots: never pass, address column 0
                    reinterpret_cast<float4*>(s_ntc)[i] = ninf4;
                    reinterpret_cast<uint2*>(s_cols)[i] = col00;
                }
                __syncwarp();
                for (int i1 = 0; i1 < c; i1 += 256) {
                    uint2 ev[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) ev[j] = (i1 + 32 * j + lane < c) ? __ldcg(src + i1 + 32 * j + lane) : make_uint2(0xffffffffu, 0u);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint2 e = ev[j];
                        if (e.x == 0xffffffffu) continue;  // (no candidate has rank 32767)
                        const int r = t * kTile + (int)((e.x >> 12) & 31u), rank = (int)(e.x >> 17);
                        const int nqr = (bu.rowCnt[r] + 3) >> 2;
                        const int slot = (rank >= nqr) + (rank >= 2 * nqr) + (rank >= 3 * nqr);  // rank / nqr, rank < 4 nqr
                        const int q = qb.rowQ[r] + (rank - slot * nqr) - q0;
                        if ((unsigned)q >= (unsigned)nqc) continue;  // another chunk's quad
                        s_ntc[q * 4 + slot] = -__uint_as_float(e.y);
                        s_cols[q * 4 + slot] = (unsigned short)(col_addr0 + ((e.x & 0xfffu) << 2));
                        if (slot == 0) s_row[q] = (unsigned short)(row_addr0 + ((uint32_t)r << 2));
                    }
                }
                __syncwarp();
                for (int i = lane; i < nqc; i += 32) {
                    __stcg(g_ntc + tile_q + q0 + i, reinterpret_cast<const float4*>(s_ntc)[i]);
                    __stcg(g_cols + tile_q + q0 + i, reinterpret_cast<const uint2*>(s_cols)[i]);
                    g_row[tile_q + q0 + i] = s_row[i];
                }
            }
        }
    }
    CVO_PHASE(22)  // instrumented variant: warp 0's tiles of the scatter; the wait for the slowest warp follows
    return true;
}

// One pass over the (x, y) list in quad form; same staging rules, trips, rotating register sets and fixed-order
// reduction as run_pass_list.  KIND: PASS_FLOW (acvo), PASS_FLOW_CVO, PASS_STEP.
template <int KIND, bool STATS>
__device__ void run_pass_quads(Smem& sm, const KParams& kp, const CloudDev& rows, const CloudDev& cols, int rank, int G,
                               uint32_t& tma_phase, const ListRef& lr) {
    constexpr int NV = PassTraits<KIND>::NV;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PassGeom pg = pass_geom(rows.n, cols.n, rank, G);
    ListSrc src;
    src.rows = &rows;
    src.cols = &cols;
    const HotConsts hc = hot_consts(sm.ic);
    const quads::StepConsts sc = quads::step_consts(hc);
    const uint32_t win = quads::smem_window(&sm);  // this CTA's shared-memory window (cluster rank << 24)
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
            CVO_PHASE(KIND == PASS_STEP ? 18 : 16)
            // Stage only what is not there already: the fixed cloud's rows survive from pass to pass and from iteration to
            // iteration, the STEP pass finds the columns the FLOW pass transformed and adds the per-row step-size terms.
            const int row_tile0 = pg.t_begin + rb * pg.t_stride, row_first = row_tile0 * kTile, col_first = cb * kTile;
            const bool have_cols = tag_is(sm.colTag, cols.g, col_first, nct * kTile, sm.serial);
            const bool have_rows = tag_is(sm.rowTag, rows.g, row_first, ntile * kTile, -1);
            if (!have_cols) stage_tiles<STAGE_GEOM>(sm, cols, col_first, nct, true, kColSentinel, tma_phase, nullptr, sm.colMask[cb / kColTiles]);
            if (!have_rows) stage_rows(sm, rows, row_tile0, pg.t_stride, ntile * kTile, false);
            if (KIND == PASS_STEP) {
                if (!have_rows) __syncthreads();
                stage_row_step_terms(sm, ntile * kTile);
            }
            CVO_PHASE(KIND == PASS_STEP ? 19 : 17)
            __syncthreads();
            if (threadIdx.x == 0) {  // read again only after the next barrier
                sm.colTag.g = cols.g; sm.colTag.first = col_first; sm.colTag.n = nct * kTile; sm.colTag.serial = sm.serial;
                sm.rowTag.g = rows.g; sm.rowTag.first = row_first; sm.rowTag.n = ntile * kTile; sm.rowTag.serial = -1;
            }
            CVO_PHASE(15)  // instrumented variant: staging of both passes
            src.row_tile0 = row_tile0;
            src.row_stride = pg.t_stride;
            src.col_base = col_first;
            // The warps take the round's trips round-robin (one quad per lane per trip).  Three register sets rotate between "being processed" and "being loaded" (never copied), so
            // the loads run TWO trips ahead of the arithmetic and the L2 prefetch kPrefetchTrips trips ahead of them.
            const quads::Round rd = quads::round_ref(lr, sm.lround[LIST_XY][round]);
            constexpr uint32_t kStep = quads::kTripStride * quads::kQuadTrip;  // quads between two trips of a warp
            const int t0 = quads::kTripStride > 1 ? warp : (rd.ntrip * warp) / kWarps;
            int left = quads::kTripStride > 1 ? (rd.ntrip - warp + kWarps - 1) / kWarps : (rd.ntrip * (warp + 1)) / kWarps - t0;
            if (left > 0) {
#ifdef CVO_PF_LANE_PREDICATE  // tuning: one prefetching lane per line (measured: 2 % SLOWER than all lanes)
                const bool pf_lane = (lane & 7) == 0;
#else
                const bool pf_lane = true;
#endif
                uint32_t qi = (uint32_t)(t0 * quads::kQuadTrip + lane);
                quads::Quad qa = quads::load_quad(rd, qi, pf_lane), qb2 = quads::load_quad(rd, qi + kStep, pf_lane), qc;
#define CVO_QUAD_TRIP(q)                                                                       \
    {                                                                                          \
        quads::QuadGeom g;                                                                     \
        const bool near = quads::quad_geom(hc, kp, win, q, g);                                     \
        if (__any_sync(0xffffffffu, near)) quads::redecide(sm, hc, kp, src, q, g);             \
        if (KIND == PASS_STEP) quads::step_quad(sc, g.rw, g, fp);                                \
        else quads::flow_quad<KIND, STATS>(hc, kp, g, fp);                                     \
    }
#pragma unroll 1
                while (true) {
                    qc = quads::load_quad(rd, qi + 2 * kStep, pf_lane);
                    CVO_QUAD_TRIP(qa)
                    qi += kStep;
                    if (--left == 0) break;
                    qa = quads::load_quad(rd, qi + 2 * kStep, pf_lane);
                    CVO_QUAD_TRIP(qb2)
                    if (KIND != PASS_STEP) quads::flush_flow<KIND, STATS>(fp, acc);  // <= 3 quads (a few rows) per f32 partial
                    else quads::flush_step(fp, acc);
                    qi += kStep;
                    if (--left == 0) break;
                    qb2 = quads::load_quad(rd, qi + 2 * kStep, pf_lane);
                    CVO_QUAD_TRIP(qc)
                    qi += kStep;
                    if (--left == 0) break;
                }
#undef CVO_QUAD_TRIP
            }
            if (KIND != PASS_STEP) quads::flush_flow<KIND, STATS>(fp, acc);
            else quads::flush_step(fp, acc);
        }
    }
    CVO_PHASE(KIND == PASS_STEP ? 4 : 2)  // instrumented variant: warp 0's trips; what follows is the wait for the slowest warp
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
        double tt = 0.0;
        if (threadIdx.x < NV)
            for (int w = 0; w < kWarps; ++w) tt += sm.u.ls.warpTot[w][threadIdx.x];
        sm.blockTot[threadIdx.x] = tt;
    }
    __syncthreads();
    CVO_PHASE(8)  // instrumented variant: reduction tail of both passes (incl. waiting for the slowest warp)
}

// cvo_lists.cuh -- neighbour candidate lists: validity policy, all-pairs sweep, wide list and filter, narrowing in place, the self-list passes of acvo
// (included by cvo_kernels.cuh inside namespace cvo_b200; see that file for the overall design)
#pragma once

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
__device__ void list_policy(Smem& sm, bool acvo, float skin, float skin_min, float shrink, float refine_min, float wide_factor, float ahead) {
    const int lane = threadIdx.x & 31;
    // |(M1 - M0) p + (t1 - t0)| is convex in p: its maximum over the moving cloud's bounding box is attained at one of
    // the 8 corners.  Lanes 0..7: against the pose the (x, y) list was built at; lanes 8..15: against the wide list's.
    // Lanes 16..23: against the previous iteration's pose -- how far, and which way, the cloud has just moved.
    double disp_xy = 0.0, disp_wide = 0.0, disp_last = 0.0;
    {
        const int grp = (lane >> 3) & 3;
        const bool wide_half = grp == 1;
        const float* tf0 = grp == 0 ? sm.lst[LIST_XY].tf : grp == 1 ? sm.wide.tf : sm.tf_prev;
        const bool have = grp == 0 ? sm.lst[LIST_XY].valid > 0 : grp == 1 ? sm.wide.valid > 0 : true;
        (void)wide_half;
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
        disp_last = (double)__shfl_sync(0xffffffffu, d, 16);
    }
    __syncwarp();  // every lane has read the build transforms before lane 0 may replace them
    if (lane != 0) return;
    {   // Fast path, f32: nothing to do while every list covers the pose (the common iteration).  The same inequalities as
        // below; f32 rounding (1e-8 m) is three orders below the margin the builds add to their radii.
        const float r_f = sqrtf(sm.ic.d2_thres);
        bool all_fine = true;
        for (int kind = 0; kind < (acvo ? LIST_KINDS : 1); ++kind) {
            const ListState& L = sm.lst[kind];
            if (L.valid < 0) continue;
            const float disp = kind == LIST_XY ? (float)disp_xy : 0.f;
            all_fine = all_fine && L.valid > 0 && (fmaxf(0.f, r_f - L.r0) + disp) * 1.000001f <= L.slack && r_f >= shrink * L.r0 * 1.000001f;
        }
        if (all_fine) {
            for (int kind = 0; kind < LIST_KINDS; ++kind) sm.lst[kind].need = 0;
            sm.wide.make = 0;
            return;
        }
    }
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
            double alpha = 0.0;  // (see below: the (x, y) list is built `ahead` skins ahead of the motion)
            if (kind == LIST_XY && ahead > 0.f && disp_last > 1.0e-7 && disp_last < 1.0e29)
                alpha = fmin((double)ahead * s * (1.0 - 1.0e-6) / disp_last, 8.0);
            const double lead = alpha * disp_last;  // distance between the current pose and the build pose
            if (kind == LIST_XY && wide_factor > 0.f) {
                // the quads as a filter of the wide list (need = 3) while it covers them -- and is not much too wide itself
                WideState& Wd = sm.wide;
                const bool covers = Wd.valid > 0 && fmax(0.0, r_now - (double)Wd.r0) + disp_wide + lead + s + margin <= (double)Wd.slack &&
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
                    for (int i = 0; i < 12; ++i) Wd.tf[i] = sm.ic.tf[i] + (float)alpha * (sm.ic.tf[i] - sm.tf_prev[i]);  // = L.tf
                }
            }
            const double rb = r_now + s + margin;
            if (L.need == 1) L.valid = 0;
            L.r0 = (float)r_now;
            L.slack = (float)(s * (1.0 - 1.0e-6));            // rounded DOWN: what the validity test may assume
            L.s_build = (float)((s + margin) * (1.0 + 1.0e-6));  // rounded UP: what the build adds to r_e
            L.thr_build = (float)(rb * rb * (1.0 + 1.0e-6));  // rounded UP: the build prefilter's ball
            L.inv_c1 = (float)(2.0 * (double)sm.st.ell * (double)sm.st.ell / 1.4426950408889634 * (1.0 + 1.0e-6));
            // The (x, y) list is built AHEAD of the motion: at the pose extrapolated along the last iteration's change by
            // alpha steps, alpha chosen so that the current pose sits `ahead` (0.5) skins behind the build pose.  The displacement
            // is linear along that line (the corner bound holds for any affine map), so while the pose keeps moving the
            // same way it first approaches the build pose and then leaves it: the list lives for up to 1.5 skins of
            // travel instead of 1.  A wrong guess costs nothing but the lifetime: coverage is tested against L.tf as ever.
#pragma unroll
            for (int i = 0; i < 12; ++i) L.tf[i] = sm.ic.tf[i] + (float)alpha * (sm.ic.tf[i] - sm.tf_prev[i]);
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
                                                int ct_begin, int ct_end, uint2* out, int limit, float thr_build, WideOut& wo,
                                                uint32_t* col_mask) {
    const int lane = threadIdx.x & 31;
    const RowTile rt = load_row_tile<true, SELF == 2, true>(sm, ws, rows, row_tf, tile);
    const float thr_box = thr_build * 1.0001f;
    uint32_t* q = sm_queue(sm);
    int qn = 0, cursor = 0, row_cnt = 0;
    for (int c0 = ct_begin; c0 < ct_end; c0 += 32) {
        uint32_t lm = live_col_tiles(sm, rt, c0, ct_end, thr_box);
        if (SELF == 0 && lm != 0u && lane == 0) atomicOr(col_mask + (c0 >> 5), lm);  // (the quad passes stage only these)
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
    // a sweep finds the live column tiles anew (a filter keeps its wide sweep's, a superset); tile 0 of every chunk always:
    // the padding slots of the quads address column 0 and must read finite coordinates
    if (SELF == 0 && !from_wide && threadIdx.x < kMaxColChunks * (kColTiles / 32))
        sm.colMask[threadIdx.x / (kColTiles / 32)][threadIdx.x % (kColTiles / 32)] = (threadIdx.x % (kColTiles / 32)) == 0 ? 1u : 0u;
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
            stage_tiles<STAGE_FULL>(sm, cols, cb * kTile, nct, col_tf, kColSentinel, tma_phase, SELF == 0 ? L.tf : nullptr);
            if (threadIdx.x == 0) sm.next_unit = 0;
            if (SELF == 0) {
                for (int i = threadIdx.x; i < ntile * kTile; i += kThreads) sm.u.of.bu.rowCnt[i] = 0;
            }
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
                    c = filter_unit_write(sm, ws, kp, L, rows, pg.t_begin + (rb + t) * pg.t_stride, (uint32_t)(t * kTile),
                                          lr.wide + kWideTable + rec.x, (int)rec.y, stage + wcur, seg - wcur);
                } else {
                    const int w0 = wo.cursor;
                    wo.out = lr.wide + kWideTable + (size_t)warp * wseg + w0;
                    wo.limit = wseg - w0;
                    wo.cursor = 0;
                    c = build_unit_write<SELF>(sm, ws, kp, L, rows, row_tf, pg.t_begin + (rb + t) * pg.t_stride, (uint32_t)(t * kTile),
                                               yy_row_min, c_begin, c_end, stage + wcur, seg - wcur, thr_build, wo,
                                               sm.colMask[cb / kColTiles]);
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
                        __stcg(lr.entries + at + i, make_uint2(__float_as_uint(1.0e30f), 0x7f800000u));
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

// Stages n rows of this CTA -- its row tiles tile0, tile0 + stride, ... of a packed cloud -- for a pass over a list:
// geometry only (transformed if the rows are the moving cloud); rows past the cloud's end are far away.
__device__ __forceinline__ void stage_rows(Smem& sm, const CloudDev& c, int tile0, int stride, int n, bool tf) {
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const int p = (tile0 + (i >> 5) * stride) * kTile + (i & 31);
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
// Self lists only (SELF = 1, 2; the row-sorted (x, y) list is rebuilt or filtered from the wide list instead): the
// entries are (d2, colour d2 | Q1 flag), d2 is pose-independent, nothing is staged.
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
                    const float d2 = __uint_as_float(v[j].x);
                    const float t_c = __fmul_rn(__uint_as_float(v[j].y & 0x7fffffffu), c2);
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

// cvo_epilogue.cuh -- the scalar part of an iteration: update_tf, thresholds, flow finalisation, line search (cubic), Exp_SEK3, stop tests, ell policies
// (included by cvo_kernels.cuh inside namespace cvo_b200; see that file for the overall design)
#pragma once

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

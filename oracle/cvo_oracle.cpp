/*
 * oracle/cvo_oracle.cpp -- CPU parity oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Plain C++17 + OpenMP restatement of the reference's registration hot path.
 * It follows, function by function (paths under
 * /root/reference/cpp/rkhs_registration/):
 *
 *   se_kernel              src/cvo.cpp:99-161, src/adaptive_cvo.cpp:92-151
 *   compute_flow           src/cvo.cpp:164-210, src/adaptive_cvo.cpp:154-272
 *   compute_step_size      src/cvo.cpp:213-308 (= src/adaptive_cvo.cpp:275-370)
 *   transform_pcd/update_tf src/cvo.cpp:83-87,310-315
 *   align                  src/cvo.cpp:361-420, src/adaptive_cvo.cpp:490-555
 *   function_inner_product src/adaptive_cvo.cpp:385-439
 *   skew / Exp_SEK3        src/LieGroup.cpp:20-27,159-186
 *   ball query semantics   thirdparty/nanoflann.hpp:249-253 (strict <),
 *                          :375-411 (d2 accumulation order)
 *
 * Eigen / TBB are not available in this image, so:
 *   - Eigen::SparseMatrix<float,RowMajor> -> hand CSR, columns ascending
 *     (what setFromTriplets produces, src/cvo.cpp:159-160);
 *   - Eigen dense f32 expressions -> explicit f32 arithmetic in the operand
 *     order of the expression, no FMA contraction (build with
 *     -ffp-contract=off) except where noted;
 *   - MatrixXf::eigenvalues() on the 3x3 companion matrix -> closed-form cubic
 *     in f64 on the f32-normalised coefficients, Newton-polished;
 *   - Matrix4f::log().norm() -> closed form  s*sqrt(2|w|^2+|v|^2)  (valid for
 *     s*|w| < pi; Q2 small-angle branch handled separately);
 *   - tbb::parallel_for -> OpenMP parallel for over the same index ranges;
 *     cross-row reductions are done in f64 in ROW ORDER (the reference's order
 *     is lock-acquisition order, i.e. non-deterministic).
 *
 * Floating-point conventions that decide set membership (and that the CUDA
 * path follows independently):
 *   y_j  = fma(m2,p2, fma(m1,p1, m0*p0)) + t        (transform_pcd)
 *   d2   = fma(dz,dz, fma(dy,dy, dx*dx))            (nanoflann tail loop with
 *          GCC's default -ffp-contract=fast under -march=native, cm/CMakeLists.txt:13)
 *   d2c  = sequential f32 sum of the 5 squared feature differences
 *   log() of the threshold ratios is evaluated in f32: with `using namespace
 *   std;` (inc/cvo.hpp:52) the call log(float) resolves to std::log(float).
 *
 * Parity status: UNPINNED by reference tests (the reference has none); pinned
 * by oracle/numpy_ref.py (independent f64 restatement) and by oracle/_ref
 * (reference nanoflann). See oracle/cvo_oracle.h.
 */
#include "cvo_oracle.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_USE_REF_NANOFLANN
/* Compiled only by oracle/Makefile target `ref`, from the reference tree where
 * it lies (-I/root/reference/cpp/rkhs_registration/thirdparty). */
#include "nanoflann.hpp"
#include "KDTreeVectorOfVectorsAdaptor.h"
#endif

namespace {

typedef std::array<float, 3> vec3;
typedef std::vector<vec3> cloud_t;  // inc/data_type.h:30 (Eigen::Vector3f -> array)

struct Csr {
    int rows = 0;
    std::vector<int64_t> rowptr;
    std::vector<int> col;
    std::vector<float> val;
    int64_t nnz() const { return (int64_t)col.size(); }
};

inline float dist2_nanoflann(const float* a, const float* b) {
    // thirdparty/nanoflann.hpp:402-406, dim = 3: only the tail loop runs;
    // `result += diff0*diff0` contracted to fma by the reference's compiler flags.
    const float d0 = a[0] - b[0];
    const float d1 = a[1] - b[1];
    const float d2 = a[2] - b[2];
    float r = d0 * d0;
    r = fmaf(d1, d1, r);
    r = fmaf(d2, d2, r);
    return r;
}

// ---------------------------------------------------------------------------
// Ball query backends
// ---------------------------------------------------------------------------
struct BallIndex {
    const cloud_t* pts = nullptr;
#ifdef ORACLE_USE_REF_NANOFLANN
    typedef KDTreeVectorOfVectorsAdaptor<cloud_t, float> kd_tree_t;
    kd_tree_t* tree = nullptr;
#endif
    explicit BallIndex(const cloud_t& p) : pts(&p) {
#ifdef ORACLE_USE_REF_NANOFLANN
        // src/cvo.cpp:112-113: adaptor ctor builds the index, then buildIndex() again.
        tree = new kd_tree_t(3, p, 10);
        tree->index->buildIndex();
#endif
    }
    ~BallIndex() {
#ifdef ORACLE_USE_REF_NANOFLANN
        delete tree;
#endif
    }
    // Appends (index, d2) of every point with d2 < r2; ascending index order.
    void query(const float* q, float r2, std::vector<std::pair<int, float>>& out) const {
        out.clear();
#ifdef ORACLE_USE_REF_NANOFLANN
        std::vector<std::pair<size_t, float>> ret;
        nanoflann::SearchParams params;  // sorted by distance, src/cvo.cpp:122-125
        tree->index->radiusSearch(q, r2, ret, params);
        out.reserve(ret.size());
        for (auto& pr : ret) out.emplace_back((int)pr.first, pr.second);
        // setFromTriplets orders each row by column (src/cvo.cpp:159)
        std::sort(out.begin(), out.end(),
                  [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.first < b.first; });
#else
        const cloud_t& P = *pts;
        const int n = (int)P.size();
        for (int j = 0; j < n; ++j) {
            const float d2 = dist2_nanoflann(q, P[j].data());
            if (d2 < r2) out.emplace_back(j, d2);  // strict <, thirdparty/nanoflann.hpp:249-253
        }
#endif
    }
};

// ---------------------------------------------------------------------------
// se_kernel (src/cvo.cpp:99-161; generalised form src/adaptive_cvo.cpp:92-151)
// ---------------------------------------------------------------------------
struct KernelParams {
    float l, s2, sp_thres, c_ell, c_sigma, c_gate_thres;
};

inline void kernel_thresholds(const KernelParams& k, float& d2_thres, float& d2_c_thres) {
    // src/cvo.cpp:102-103 -- log() on a float argument is std::log(float).
    d2_thres = (float)(-2.0 * k.l * k.l * std::log(k.sp_thres / k.s2));
    d2_c_thres = (float)(-2.0 * k.c_ell * k.c_ell * std::log(k.c_gate_thres / k.c_sigma / k.c_sigma));
}

inline bool kernel_value(const KernelParams& kp, float d2, float d2_thres, float d2_c_thres,
                         const float* fa, const float* fb, float& a_out) {
    if (!(d2 < d2_thres)) return false;  // src/cvo.cpp:143
    float d2_color = 0.f;                // (feature_x-feature_y).squaredNorm(), src/cvo.cpp:146
    for (int t = 0; t < 5; ++t) {
        const float df = fa[t] - fb[t];
        d2_color = d2_color + df * df;
    }
    if (!(d2_color < d2_c_thres)) return false;  // src/cvo.cpp:148
    const float k = (float)(kp.s2 * std::exp(-d2 / (2.0 * kp.l * kp.l)));                          // :149
    const float ck = (float)(kp.c_sigma * kp.c_sigma * std::exp(-d2_color / (2.0 * kp.c_ell * kp.c_ell)));  // :150
    const float a = ck * k;                                                                        // :151
    if (!(a > kp.sp_thres)) return false;                                                          // :152
    a_out = a;
    return true;
}

void se_kernel(const cloud_t& a_pos, const float* a_feat, const cloud_t& b_pos, const float* b_feat,
               const KernelParams& kp, Csr& A, int64_t* n_in_ball) {
    const int na = (int)a_pos.size();
    float d2_thres, d2_c_thres;
    kernel_thresholds(kp, d2_thres, d2_c_thres);

    BallIndex index(b_pos);
    std::vector<std::vector<std::pair<int, float>>> rows(na);
    int64_t in_ball = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : in_ball)
    for (int i = 0; i < na; ++i) {
        std::vector<std::pair<int, float>> matches;
        index.query(a_pos[i].data(), d2_thres, matches);
        in_ball += (int64_t)matches.size();
        const float* fa = a_feat + (size_t)i * 5;
        std::vector<std::pair<int, float>>& row = rows[i];
        for (auto& m : matches) {
            float a;
            if (kernel_value(kp, m.second, d2_thres, d2_c_thres, fa, b_feat + (size_t)m.first * 5, a))
                row.emplace_back(m.first, a);
        }
    }
    if (n_in_ball) *n_in_ball = in_ball;
    A.rows = na;
    A.rowptr.assign(na + 1, 0);
    for (int i = 0; i < na; ++i) A.rowptr[i + 1] = A.rowptr[i] + (int64_t)rows[i].size();
    A.col.resize(A.rowptr[na]);
    A.val.resize(A.rowptr[na]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < na; ++i) {
        int64_t o = A.rowptr[i];
        for (auto& e : rows[i]) {
            A.col[o] = e.first;
            A.val[o] = e.second;
            ++o;
        }
    }
}

// ---------------------------------------------------------------------------
// small f32 3x3 helpers in Eigen's evaluation order
// ---------------------------------------------------------------------------
struct Mat3 {
    float m[9];  // row-major
};

inline Mat3 mat3_identity() {
    Mat3 I;
    for (int i = 0; i < 9; ++i) I.m[i] = 0.f;
    I.m[0] = I.m[4] = I.m[8] = 1.f;
    return I;
}
inline Mat3 mat3_mul(const Mat3& a, const Mat3& b) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r.m[i * 3 + j] = (a.m[i * 3 + 0] * b.m[0 * 3 + j] + a.m[i * 3 + 1] * b.m[1 * 3 + j]) + a.m[i * 3 + 2] * b.m[2 * 3 + j];
    return r;
}
inline vec3 mat3_vec(const Mat3& a, const vec3& v) {
    vec3 r;
    for (int i = 0; i < 3; ++i) r[i] = (a.m[i * 3 + 0] * v[0] + a.m[i * 3 + 1] * v[1]) + a.m[i * 3 + 2] * v[2];
    return r;
}
inline Mat3 skew(const vec3& v) {  // src/LieGroup.cpp:20-27
    Mat3 M;
    M.m[0] = 0.f;   M.m[1] = -v[2]; M.m[2] = v[1];
    M.m[3] = v[2];  M.m[4] = 0.f;   M.m[5] = -v[0];
    M.m[6] = -v[1]; M.m[7] = v[0];  M.m[8] = 0.f;
    return M;
}
inline vec3 cross(const vec3& a, const vec3& b) {
    return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
inline float dot3(const vec3& a, const vec3& b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline float norm3(const vec3& a) { return std::sqrt(dot3(a, a)); }

// Exp_SEK3, K = 1 (src/LieGroup.cpp:159-186)
void exp_sek3(const vec3& w, const vec3& v, float dt, Mat3& R, vec3& t) {
    const float TOLERANCE = 1e-6f;  // src/LieGroup.cpp:18
    const float theta = norm3(w);
    Mat3 Jl;
    const Mat3 I = mat3_identity();
    if (theta < TOLERANCE) {  // Q2: Jl = I, not dt*I
        R = I;
        Jl = I;
    } else {
        const Mat3 A = skew(w);
        const float theta2 = theta * theta;
        const float stheta = std::sin(dt * theta);
        const float ctheta = std::cos(dt * theta);
        const float oneMinusCosTheta2 = (1 - ctheta) / (theta2);
        const Mat3 A2 = mat3_mul(A, A);
        const float sa = stheta / theta;
        const float sj = (dt * theta - stheta) / (theta2 * theta);
        for (int i = 0; i < 9; ++i) {
            R.m[i] = (I.m[i] + sa * A.m[i]) + oneMinusCosTheta2 * A2.m[i];
            Jl.m[i] = (dt * I.m[i] + oneMinusCosTheta2 * A.m[i]) + sj * A2.m[i];
        }
    }
    t = mat3_vec(Jl, v);
}

// dist_se3 (src/cvo.cpp:71-81): ||logm([dR dT; 0 1])||_F, closed form for the
// matrix Exp_SEK3 produced from (w, v, s).
float dist_se3_closed(const vec3& w, const vec3& v, float s) {
    const float theta = norm3(w);
    if (theta < 1e-6f) {
        // dR = I, dT = v  =>  logm = [0 v; 0 0]
        return norm3(v);
    }
    // logm(Exp(s*[w^ v;0 0])) = s*[w^ v; 0 0] while s*theta < pi
    const float w2 = dot3(w, w), v2 = dot3(v, v);
    return s * std::sqrt(2.f * w2 + v2);
}

// poly_solver + root selection (src/cvo.cpp:53-69, 291-307)
float step_from_coeffs(double B, double C, double D, double E, float min_step, float max_step) {
    // p_coef << 4.0*float(E), 3.0*float(D), 2.0*float(C), float(B)   (f32 vector)
    const float p0 = (float)(4.0 * (float)E);
    const float p1 = (float)(3.0 * (float)D);
    const float p2 = (float)(2.0 * (float)C);
    const float p3 = (float)B;
    // companion first row: -(coef/coef(0)).segment(1,3)   (f32 division)
    const float a2f = p1 / p0, a1f = p2 / p0, a0f = p3 / p0;
    float temp_step = std::numeric_limits<float>::max();
    if (std::isfinite(a2f) && std::isfinite(a1f) && std::isfinite(a0f)) {
        // roots of x^3 + a2 x^2 + a1 x + a0, f64 closed form
        const double a2 = a2f, a1 = a1f, a0 = a0f;
        const double q = (3.0 * a1 - a2 * a2) / 9.0;
        const double r = (9.0 * a2 * a1 - 27.0 * a0 - 2.0 * a2 * a2 * a2) / 54.0;
        const double disc = q * q * q + r * r;
        double roots[3];
        int nroots = 0;
        if (disc > 0) {
            const double sd = std::sqrt(disc);
            const double s = std::cbrt(r + sd), t = std::cbrt(r - sd);
            roots[nroots++] = s + t - a2 / 3.0;
        } else if (disc == 0) {
            const double s = std::cbrt(r);
            roots[nroots++] = 2 * s - a2 / 3.0;
            roots[nroots++] = -s - a2 / 3.0;
        } else {
            const double th = std::acos(std::max(-1.0, std::min(1.0, r / std::sqrt(-q * q * q))));
            const double m = 2.0 * std::sqrt(-q);
            roots[nroots++] = m * std::cos(th / 3.0) - a2 / 3.0;
            roots[nroots++] = m * std::cos((th + 2.0 * M_PI) / 3.0) - a2 / 3.0;
            roots[nroots++] = m * std::cos((th + 4.0 * M_PI) / 3.0) - a2 / 3.0;
        }
        for (int i = 0; i < nroots; ++i) {
            double x = roots[i];
            for (int it = 0; it < 3; ++it) {  // Newton polish
                const double f = ((x + a2) * x + a1) * x + a0;
                const double fp = (3.0 * x + 2.0 * a2) * x + a1;
                if (fp == 0 || !std::isfinite(f)) break;
                const double xn = x - f / fp;
                if (!std::isfinite(xn)) break;
                x = xn;
            }
            const float xr = (float)x;
            if (xr > 0 && xr < temp_step) temp_step = xr;  // src/cvo.cpp:299-301
        }
    }
    float step = (temp_step == std::numeric_limits<float>::max()) ? min_step : temp_step;  // :304
    step = step > max_step ? max_step : step;                                                 // :307
    return step;
}

// ---------------------------------------------------------------------------
// registration state
// ---------------------------------------------------------------------------
struct State {
    const oracle_params* p;
    int num_fixed, num_moving;
    cloud_t x;           // fixed positions (cloud_x)
    cloud_t y0;          // moving positions, original
    cloud_t y;           // moving positions, transformed (cloud_y)
    const float* fx;     // N x 5 row-major
    const float* fy;     // M x 5 row-major
    Mat3 R;
    vec3 T;
    float ell, ell_max;
    Mat3 tf_lin;         // transform.linear()
    vec3 tf_trans;       // transform.translation()
    Csr A, Axx, Ayy;
    vec3 omega, v;
    double dl = 0, dl_num = 0;
    double B = 0, C = 0, D = 0, E = 0, sum_a = 0;
    float step = 0;
    int64_t n_in_ball = 0;
};

void update_tf(State& s) {  // src/cvo.cpp:83-87
    Mat3 Rt;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rt.m[i * 3 + j] = s.R.m[j * 3 + i];
    s.tf_lin = Rt;
    // -R.transpose()*T : (-R^T) * T
    Mat3 nRt;
    for (int i = 0; i < 9; ++i) nRt.m[i] = -Rt.m[i];
    s.tf_trans = mat3_vec(nRt, s.T);
}

void transform_pcd(State& s) {  // src/cvo.cpp:310-315
    const Mat3& L = s.tf_lin;
    const vec3& t = s.tf_trans;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < s.num_moving; ++j) {
        const vec3& p = s.y0[j];
        vec3 q;
        for (int i = 0; i < 3; ++i) {
            float acc = L.m[i * 3 + 0] * p[0];
            acc = fmaf(L.m[i * 3 + 1], p[1], acc);
            acc = fmaf(L.m[i * 3 + 2], p[2], acc);
            q[i] = acc + t[i];
        }
        s.y[j] = q;
    }
}

KernelParams kparams_xy(const State& s) {
    const oracle_params& p = *s.p;
    KernelParams k;
    k.l = s.ell;
    k.s2 = p.sigma * p.sigma;
    k.sp_thres = p.sp_thres;
    k.c_ell = p.c_ell;
    k.c_sigma = p.c_sigma;
    // cvo gates colour with sp_thres (src/cvo.cpp:103); acvo with c_sp_thres (src/adaptive_cvo.cpp:101)
    k.c_gate_thres = (p.mode == ORACLE_MODE_ACVO) ? p.c_sp_thres : p.sp_thres;
    return k;
}

void compute_flow(State& s) {
    const oracle_params& p = *s.p;
    const KernelParams kp = kparams_xy(s);
    const bool acvo = (p.mode == ORACLE_MODE_ACVO);
    se_kernel(s.x, s.fx, s.y, s.fy, kp, s.A, &s.n_in_ball);  // src/cvo.cpp:166 / src/adaptive_cvo.cpp:156
    if (acvo) {
        se_kernel(s.x, s.fx, s.x, s.fx, kp, s.Axx, nullptr);  // src/adaptive_cvo.cpp:159
        se_kernel(s.y, s.fy, s.y, s.fy, kp, s.Ayy, nullptr);  // :160 (transformed y)
    }

    const int N = s.num_fixed, M = s.num_moving;
    const float inv_c = 1 / p.c, inv_d = 1 / p.d;
    const float ell_3 = s.ell * s.ell * s.ell;          // src/adaptive_cvo.cpp:171
    const float inv_ell3 = 1 / ell_3;
    std::vector<std::array<double, 8>> part(N);

#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < N; ++i) {
        // (1/c*Ai*cross_xy): row vector (1/c*Ai) times nnz x 3, f32, column order
        float po[3] = {0, 0, 0}, pv[3] = {0, 0, 0}, pdl_xy = 0.f, psum = 0.f;
        const vec3& xi = s.x[i];
        for (int64_t e = s.A.rowptr[i]; e < s.A.rowptr[i + 1]; ++e) {
            const int idx = s.A.col[e];
            const float a = s.A.val[e];
            const vec3& yj = s.y[idx];
            const vec3 cr = cross(xi, yj);                                   // src/cvo.cpp:191
            const vec3 df = {yj[0] - xi[0], yj[1] - xi[1], yj[2] - xi[2]};   // :192
            const float ac = inv_c * a, ad = inv_d * a;
            for (int t = 0; t < 3; ++t) {
                po[t] = po[t] + ac * cr[t];
                pv[t] = pv[t] + ad * df[t];
            }
            psum = psum + a;
            if (acvo) {
                const float n2 = (df[0] * df[0] + df[1] * df[1]) + df[2] * df[2];  // src/adaptive_cvo.cpp:202
                pdl_xy = pdl_xy + (inv_ell3 * a) * n2;
            }
        }
        double partial_dl = 0;
        if (acvo) {
            // Q1: for rows i < num_moving the Ayy loop never fills sum_diff_yy_2
            // (src/adaptive_cvo.cpp:213-223) => contributes exactly 0.
            if (i < M) partial_dl += (double)(0.f);
            partial_dl -= (double)(2 * pdl_xy);                                   // :228
            float pdl_xx = 0.f;
            for (int64_t e = s.Axx.rowptr[i]; e < s.Axx.rowptr[i + 1]; ++e) {
                const int idx = s.Axx.col[e];
                const vec3& xj = s.x[idx];
                const vec3 df = {xj[0] - xi[0], xj[1] - xi[1], xj[2] - xi[2]};
                const float n2 = (df[0] * df[0] + df[1] * df[1]) + df[2] * df[2];
                pdl_xx = pdl_xx + (inv_ell3 * s.Axx.val[e]) * n2;
            }
            partial_dl += (double)pdl_xx;                                          // :231
        }
        part[i] = {(double)po[0], (double)po[1], (double)po[2], (double)pv[0], (double)pv[1], (double)pv[2],
                   partial_dl, (double)psum};
    }
    double dw[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, dl = 0, sum_a = 0;
    for (int i = 0; i < N; ++i) {
        for (int t = 0; t < 3; ++t) {
            dw[t] += part[i][t];
            dv[t] += part[i][3 + t];
        }
        dl += part[i][6];
        sum_a += part[i][7];
    }
    if (acvo && M > N) {  // src/adaptive_cvo.cpp:243-265
        std::vector<double> pyy(M, 0.0);
#pragma omp parallel for schedule(dynamic, 16)
        for (int i = N; i < M; ++i) {
            float acc = 0.f;
            const vec3& yi = s.y[i];
            for (int64_t e = s.Ayy.rowptr[i]; e < s.Ayy.rowptr[i + 1]; ++e) {
                const vec3& yj = s.y[s.Ayy.col[e]];
                const vec3 df = {yj[0] - yi[0], yj[1] - yi[1], yj[2] - yi[2]};
                const float n2 = (df[0] * df[0] + df[1] * df[1]) + df[2] * df[2];
                acc = acc + (inv_ell3 * s.Ayy.val[e]) * n2;
            }
            pyy[i] = (double)acc;
        }
        for (int i = N; i < M; ++i) dl += pyy[i];
    }
    s.omega = {(float)dw[0], (float)dw[1], (float)dw[2]};  // src/cvo.cpp:208-209
    s.v = {(float)dv[0], (float)dv[1], (float)dv[2]};
    s.sum_a = sum_a;
    if (acvo) {
        s.dl_num = dl;
        const long long den = (long long)s.Axx.nnz() + (long long)s.Ayy.nnz() - 2 * (long long)s.A.nnz();
        s.dl = dl / (double)den;  // src/adaptive_cvo.cpp:271
    } else {
        s.dl = 0;
        s.dl_num = 0;
    }
}

void compute_step_size(State& s) {  // src/cvo.cpp:213-308
    const oracle_params& p = *s.p;
    const int N = s.num_fixed, M = s.num_moving;
    const Mat3 W = skew(s.omega);
    const Mat3 W2 = mat3_mul(W, W);
    const Mat3 W3 = mat3_mul(W2, W);
    const Mat3 W4 = mat3_mul(W3, W);
    const vec3 Wv = mat3_vec(W, s.v);
    const vec3 W2v = mat3_vec(W2, s.v);
    const vec3 W3v = mat3_vec(W3, s.v);

    std::vector<vec3> xiz(M), xi2z(M), xi3z(M), xi4z(M);
    std::vector<float> normxiz2(M), xiz_dot_xi2z(M), epsil_const(M);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < M; ++j) {
        const vec3& y = s.y[j];
        const vec3 c1 = cross(s.omega, y);
        const vec3 a1 = {c1[0] + s.v[0], c1[1] + s.v[1], c1[2] + s.v[2]};          // :228
        const vec3 m2 = mat3_vec(W2, y), m3 = mat3_vec(W3, y), m4 = mat3_vec(W4, y);
        const vec3 a2 = {m2[0] + Wv[0], m2[1] + Wv[1], m2[2] + Wv[2]};             // :229-230
        const vec3 a3 = {m3[0] + W2v[0], m3[1] + W2v[1], m3[2] + W2v[2]};          // :231-232
        const vec3 a4 = {m4[0] + W3v[0], m4[1] + W3v[1], m4[2] + W3v[2]};          // :233-234
        xiz[j] = a1; xi2z[j] = a2; xi3z[j] = a3; xi4z[j] = a4;
        normxiz2[j] = dot3(a1, a1);                                                 // :235
        xiz_dot_xi2z[j] = -dot3(a1, a2);                                            // :236
        epsil_const[j] = dot3(a2, a2) + 2 * dot3(a1, a3);                           // :237
    }

    const float temp_coef = (float)(1 / (2.0 * s.ell * s.ell));  // :241
    const float m2t = (float)(-2.0 * temp_coef);                 // scalar folded to f32 by Eigen
    const float p2t = (float)(2.0 * temp_coef);
    std::vector<std::array<double, 4>> part(N);
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < N; ++i) {
        double Bi = 0, Ci = 0, Di = 0, Ei = 0;
        const vec3& xi = s.x[i];
        for (int64_t e = s.A.rowptr[i]; e < s.A.rowptr[i + 1]; ++e) {
            const int idx = s.A.col[e];
            const vec3& yj = s.y[idx];
            const vec3 r = {xi[0] - yj[0], xi[1] - yj[1], xi[2] - yj[2]};  // :260
            const vec3& z1 = xiz[idx]; const vec3& z2 = xi2z[idx];
            const vec3& z3 = xi3z[idx]; const vec3& z4 = xi4z[idx];
            const float beta_ij = ((m2t * z1[0]) * r[0] + (m2t * z1[1]) * r[1]) + (m2t * z1[2]) * r[2];           // :262
            const float g_in = ((2.f * z2[0]) * r[0] + (2.f * z2[1]) * r[1]) + (2.f * z2[2]) * r[2];
            const float gamma_ij = -temp_coef * (normxiz2[idx] + g_in);                                            // :264-265
            const float d_in = ((-z3[0]) * r[0] + (-z3[1]) * r[1]) + (-z3[2]) * r[2];
            const float delta_ij = p2t * (xiz_dot_xi2z[idx] + d_in);                                               // :267-268
            const float e_in = ((2.f * z4[0]) * r[0] + (2.f * z4[1]) * r[1]) + (2.f * z4[2]) * r[2];
            const float epsil_ij = -temp_coef * (epsil_const[idx] + e_in);                                         // :270-271
            const float A_ij = s.A.val[e];
            Bi += double(A_ij * beta_ij);                                                                           // :275
            Ci += double(A_ij * (gamma_ij + beta_ij * beta_ij / 2.0));                                              // :276
            Di += double(A_ij * (delta_ij + beta_ij * gamma_ij + beta_ij * beta_ij * beta_ij / 6.0));               // :277
            Ei += double(A_ij * (epsil_ij + beta_ij * delta_ij + 1 / 2.0 * beta_ij * beta_ij * gamma_ij +
                                 1 / 2.0 * gamma_ij * gamma_ij + 1 / 24.0 * beta_ij * beta_ij * beta_ij * beta_ij));  // :278-279
        }
        part[i] = {Bi, Ci, Di, Ei};
    }
    double B = 0, C = 0, D = 0, E = 0;
    for (int i = 0; i < N; ++i) {
        B += part[i][0]; C += part[i][1]; D += part[i][2]; E += part[i][3];
    }
    s.B = B; s.C = C; s.D = D; s.E = E;
    s.step = step_from_coeffs(B, C, D, E, p.min_step, p.max_step);
}

void init_state(State& s, const oracle_params* p, const float* x_pos, const float* x_feat, int n_fixed,
                const float* y_pos, const float* y_feat, int n_moving, const float* R, const float* T, float ell) {
    s.p = p;
    s.num_fixed = n_fixed;
    s.num_moving = n_moving;
    s.x.resize(n_fixed);
    s.y0.resize(n_moving);
    s.y.resize(n_moving);
    for (int i = 0; i < n_fixed; ++i) s.x[i] = {x_pos[3 * i], x_pos[3 * i + 1], x_pos[3 * i + 2]};
    for (int j = 0; j < n_moving; ++j) s.y0[j] = {y_pos[3 * j], y_pos[3 * j + 1], y_pos[3 * j + 2]};
    s.y = s.y0;  // src/cvo.cpp:351
    s.fx = x_feat;
    s.fy = y_feat;
    for (int i = 0; i < 9; ++i) s.R.m[i] = R[i];
    s.T = {T[0], T[1], T[2]};
    s.ell = ell;
    s.ell_max = p->ell_max;
    s.omega = {0, 0, 0};
    s.v = {0, 0, 0};
}

void write_tf44(const State& s, float* out) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) out[i * 4 + j] = s.tf_lin.m[i * 3 + j];
        out[i * 4 + 3] = s.tf_trans[i];
    }
    out[12] = out[13] = out[14] = 0.f;
    out[15] = 1.f;
}

}  // namespace

extern "C" {

void oracle_default_params_cvo(oracle_params* p) {  // src/cvo.cpp:18-48
    std::memset(p, 0, sizeof(*p));
    p->mode = ORACLE_MODE_CVO;
    p->ell_policy = ORACLE_ELL_SCHEDULE;
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

void oracle_default_params_acvo(oracle_params* p) {  // src/adaptive_cvo.cpp:18-50
    oracle_default_params_cvo(p);
    p->mode = ORACLE_MODE_ACVO;
    p->ell_policy = ORACLE_ELL_ADAPTIVE;
    p->ell_init = 0.1f;
    p->ell_min = 0.0391f;
    p->ell_max = 0.15f;
    p->dl_step = 0.3;
    p->sp_thres = 8.315e-3f;
    p->c_ell = 0.5f;
    p->c_sp_thres = 8.315e-3f;
}

int oracle_eval(const float* x_pos, const float* x_feat, int n_fixed, const float* y_pos, const float* y_feat,
                int n_moving, const float* R, const float* T, float ell, const oracle_params* p,
                oracle_eval_out* out) {
    if (n_fixed <= 0 || n_moving <= 0) return -1;
    State s;
    init_state(s, p, x_pos, x_feat, n_fixed, y_pos, y_feat, n_moving, R, T, ell);
    update_tf(s);
    transform_pcd(s);
    compute_flow(s);
    compute_step_size(s);
    std::memset(out, 0, sizeof(*out));
    out->nnz = s.A.nnz();
    out->sum_a = s.sum_a;
    for (int t = 0; t < 3; ++t) {
        out->omega[t] = s.omega[t];
        out->v[t] = s.v[t];
    }
    out->B = s.B; out->C = s.C; out->D = s.D; out->E = s.E;
    out->step = s.step;
    out->nnz_xx = s.Axx.nnz();
    out->nnz_yy = s.Ayy.nnz();
    out->dl = s.dl;
    out->dl_num = s.dl_num;
    out->n_in_ball = s.n_in_ball;
    return 0;
}

int oracle_align(const float* x_pos, const float* x_feat, int n_fixed, const float* y_pos, const float* y_feat,
                 int n_moving, const oracle_params* p, float* R, float* T, float* ell, float* transform_out,
                 float* prev_transform_out, int* iters_out, int* status_out, oracle_trace_rec* trace,
                 int trace_cap, int* trace_len) {
    if (n_fixed <= 0 || n_moving <= 0) return -1;
    State s;
    init_state(s, p, x_pos, x_feat, n_fixed, y_pos, y_feat, n_moving, R, T, *ell);
    const bool acvo = (p->mode == ORACLE_MODE_ACVO);
    const int max_iter = p->fixed_iters > 0 ? p->fixed_iters : p->max_iter;
    const bool stops = !(p->fixed_iters > 0);
    int iters = max_iter, status = 0, nrec = 0;
    update_tf(s);
    for (int k = 0; k < max_iter; ++k) {
        update_tf(s);         // src/cvo.cpp:368
        transform_pcd(s);     // :371
        compute_flow(s);      // :374
        compute_step_size(s); // :377

        oracle_trace_rec rec;
        std::memset(&rec, 0, sizeof(rec));
        rec.ell = s.ell;
        rec.step = s.step;
        for (int t = 0; t < 3; ++t) { rec.omega[t] = s.omega[t]; rec.v[t] = s.v[t]; }
        rec.B = s.B; rec.C = s.C; rec.D = s.D; rec.E = s.E;
        rec.sum_a = s.sum_a; rec.dl = s.dl;
        rec.nnz = s.A.nnz(); rec.nnz_xx = s.Axx.nnz(); rec.nnz_yy = s.Ayy.nnz();

        bool stop = false;
        if (stops) {
            bool small;
            if (acvo)  // src/adaptive_cvo.cpp:509 (norms in f64)
                small = std::sqrt((double)s.omega[0] * s.omega[0] + (double)s.omega[1] * s.omega[1] + (double)s.omega[2] * s.omega[2]) < p->eps &&
                        std::sqrt((double)s.v[0] * s.v[0] + (double)s.v[1] * s.v[1] + (double)s.v[2] * s.v[2]) < p->eps;
            else       // src/cvo.cpp:380
                small = norm3(s.omega) < p->eps && norm3(s.v) < p->eps;
            if (small) { iters = k; status = 1; stop = true; }
        }
        if (!stop) {
            Mat3 dR; vec3 dT;
            exp_sek3(s.omega, s.v, s.step, dR, dT);          // :391
            const vec3 RdT = mat3_vec(s.R, dT);
            s.T = {RdT[0] + s.T[0], RdT[1] + s.T[1], RdT[2] + s.T[2]};  // :398
            s.R = mat3_mul(s.R, dR);                          // :399
            if (stops && dist_se3_closed(s.omega, s.v, s.step) < p->eps_2) {  // :402
                iters = k; status = 2; stop = true;
            }
        }
        if (!stop) {
            if (p->ell_policy == ORACLE_ELL_SCHEDULE) {       // src/cvo.cpp:408-410
                s.ell = (k > 2) ? (float)0.10 : s.ell;
                s.ell = (k > 9) ? (float)0.06 : s.ell;
                s.ell = (k > 19) ? (float)0.03 : s.ell;
            } else if (p->ell_policy == ORACLE_ELL_ADAPTIVE) { // src/adaptive_cvo.cpp:538-545
                s.ell = (float)(s.ell + p->dl_step * s.dl);
                if (s.ell >= s.ell_max) {
                    s.ell = (float)(s.ell_max * 0.7);
                    s.ell_max = (float)(s.ell_max * 0.7);
                }
                s.ell = (s.ell < p->ell_min) ? p->ell_min : s.ell;
            }
        }
        for (int t = 0; t < 9; ++t) rec.R[t] = s.R.m[t];
        for (int t = 0; t < 3; ++t) rec.T[t] = s.T[t];
        if (trace && nrec < trace_cap) trace[nrec] = rec;
        ++nrec;
        if (stop) break;
    }
    // src/cvo.cpp:413-415: prev_transform = transform (stale, Q3); update_tf()
    if (prev_transform_out) write_tf44(s, prev_transform_out);
    update_tf(s);
    if (transform_out) write_tf44(s, transform_out);
    for (int i = 0; i < 9; ++i) R[i] = s.R.m[i];
    for (int i = 0; i < 3; ++i) T[i] = s.T[i];
    *ell = s.ell;
    if (iters_out) *iters_out = iters;
    if (status_out) *status_out = status;
    if (trace_len) *trace_len = nrec;
    return 0;
}

float oracle_inner_product(const float* a_pos, const float* a_feat, int n_a, const float* b_pos,
                           const float* b_feat, int n_b, float ell, const oracle_params* p, double* sum_a_out,
                           long long* count_out) {
    // src/adaptive_cvo.cpp:385-439
    cloud_t a(n_a), b(n_b);
    for (int i = 0; i < n_a; ++i) a[i] = {a_pos[3 * i], a_pos[3 * i + 1], a_pos[3 * i + 2]};
    for (int j = 0; j < n_b; ++j) b[j] = {b_pos[3 * j], b_pos[3 * j + 1], b_pos[3 * j + 2]};
    KernelParams kp;
    kp.l = ell;
    kp.s2 = p->sigma * p->sigma;  // :424 uses sigma*sigma; :391 divides by sigma twice
    kp.sp_thres = p->sp_thres;
    kp.c_ell = p->c_ell;
    kp.c_sigma = p->c_sigma;
    kp.c_gate_thres = p->sp_thres;  // :392
    float d2_thres = (float)(-2.0 * ell * ell * std::log(p->sp_thres / p->sigma / p->sigma));  // :391
    float d2_c_thres = (float)(-2.0 * p->c_ell * p->c_ell * std::log(p->sp_thres / p->c_sigma / p->c_sigma));
    BallIndex index(b);
    double sum_A = 0, sum = 0;
    std::vector<double> row_sum(n_a, 0.0);
    std::vector<long long> row_cnt(n_a, 0);
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n_a; ++i) {
        std::vector<std::pair<int, float>> matches;
        index.query(a[i].data(), d2_thres, matches);
        double rs = 0;
        long long rc = 0;
        for (auto& m : matches) {
            float av;
            if (kernel_value(kp, m.second, d2_thres, d2_c_thres, a_feat + (size_t)i * 5,
                             b_feat + (size_t)m.first * 5, av)) {
                rs += av;  // sum_A += a (f64 accumulator), :429
                rc += 1;
            }
        }
        row_sum[i] = rs;
        row_cnt[i] = rc;
    }
    for (int i = 0; i < n_a; ++i) { sum_A += row_sum[i]; sum += (double)row_cnt[i]; }
    if (sum_a_out) *sum_a_out = sum_A;
    if (count_out) *count_out = (long long)sum;
    return (float)(sum_A / sum);  // :438
}

void oracle_exp_sek3(const float* omega, const float* v, float dt, float* dR, float* dT) {
    Mat3 R; vec3 t;
    exp_sek3({omega[0], omega[1], omega[2]}, {v[0], v[1], v[2]}, dt, R, t);
    for (int i = 0; i < 9; ++i) dR[i] = R.m[i];
    for (int i = 0; i < 3; ++i) dT[i] = t[i];
}

float oracle_step_from_coeffs(double B, double C, double D, double E, float min_step, float max_step) {
    return step_from_coeffs(B, C, D, E, min_step, max_step);
}

int oracle_ball_query(const float* pts, int n, const float* q, float r2, int* idx_out, float* d2_out, int cap) {
    cloud_t P(n);
    for (int i = 0; i < n; ++i) P[i] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    BallIndex index(P);
    std::vector<std::pair<int, float>> m;
    index.query(q, r2, m);
    for (int i = 0; i < (int)m.size() && i < cap; ++i) {
        idx_out[i] = m[i].first;
        d2_out[i] = m[i].second;
    }
    return (int)m.size();
}

const char* oracle_backend(void) {
#ifdef ORACLE_USE_REF_NANOFLANN
    return "reference-nanoflann-kdtree";
#else
    return "brute-force";
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"

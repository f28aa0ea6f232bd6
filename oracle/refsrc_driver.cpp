// oracle/refsrc_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C interface around the REFERENCE'S OWN first-party sources, compiled unmodified where they lie:
//     /root/reference/cpp/rkhs_registration/src/cvo.cpp, src/adaptive_cvo.cpp, src/LieGroup.cpp
//     + thirdparty/nanoflann.hpp, thirdparty/KDTreeVectorOfVectorsAdaptor.h, include/*.h*
// against oracle/shim/ (stand-ins for the Eigen / TBB / OpenCV / PCL / Boost headers they include, none of which exists
// in this image).  The result, oracle/_ref/libcvo_refsrc.so (oracle/Makefile target `refsrc`), is what PINS the
// restatement oracle/cvo_oracle.cpp: tests/golden/make_refsrc_golden.py runs both on the same seeded inputs and commits
// the reference-side outputs as fixtures (tests/golden/refsrc_golden.json).
//
// What is reference code here and what is not:
//   * cvo::cvo / acvo::acvo -- constructors (all parameters), set_pcd(), align(), se_kernel(), compute_flow(),
//     compute_step_size(), transform_pcd(), update_tf(), poly_solver(), dist_se3(), function_inner_product(), and
//     LieGroup.cpp's skew() / Exp_SEK3(): the reference's own object code.
//   * NOT compiled: src/pcd_generator.cpp and thirdparty/PixelSelector2.cpp (the image front end needs the real
//     OpenCV).  This file supplies stand-ins for the four pcd_generator members set_pcd() calls; they hand the test
//     case's prepared cloud (N x 3 positions, N x 5 features) through the opaque cv::Mat of oracle/shim.
//   * The Eigen arithmetic is the shim's (oracle/shim/eigen/shim_eigen.hpp): same types, same promotion rules, scalar-type
//     accumulation; summation order inside a product and the 3 x 3 eigenvalue iteration are NOT bit-identical to any
//     particular Eigen release (the reference pins none: "Eigen3", README.md:11).
//   * Private members are reached with -fno-access-control (this translation unit only).  The per-iteration trace
//     (mode 1) and the fixed-ell benchmark schedule (mode 2) re-drive the reference's private functions from a loop in
//     this file that mirrors align()'s body; mode 0 is the reference's align() itself, and
//     tests/test_refsrc.py checks that mode 1 reproduces mode 0 bit for bit.
#include <cmath>
#include <cstring>
#include <iostream>
#include <limits>
#include <sstream>

#include "adaptive_cvo.hpp"
#include "cvo.hpp"
#include "cvo_oracle.h"

namespace {
struct CloudPayload {
    const float* xyz;   // n x 3
    const float* feat;  // n x 5 row-major
    int n;
};
void fill_cloud(const CloudPayload& p, cvo::point_cloud* pc) {
    pc->num_points = p.n;
    pc->positions.resize(p.n);
    pc->features.resize(p.n, NUM_FEATURES);
    for (int i = 0; i < p.n; ++i) {
        pc->positions[i] = Eigen::Vector3f(p.xyz[3 * i], p.xyz[3 * i + 1], p.xyz[3 * i + 2]);
        for (int j = 0; j < NUM_FEATURES; ++j) pc->features(i, j) = p.feat[NUM_FEATURES * i + j];
    }
}
struct QuietCout {  // the reference prints from set_pcd() / align()
    std::ostringstream sink;
    std::streambuf* old;
    QuietCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~QuietCout() { std::cout.rdbuf(old); }
};
}  // namespace

// ---- stand-ins for the image front end (src/pcd_generator.cpp is not compiled) ----------------------------------
namespace cvo {
pcd_generator::pcd_generator() : num_want(3000), dep_thres(20000), map(nullptr), cam_info{1000, 616.368f, 616.745f, 319.935f, 243.639f} {}
pcd_generator::~pcd_generator() {}
void pcd_generator::load_image(const cv::Mat& RGB_img, const cv::Mat& dep_img, frame* ptr_fr) {
    ptr_fr->image = RGB_img;  // carries the CloudPayload pointer
    ptr_fr->depth = dep_img;
}
void pcd_generator::create_pointcloud(const int, frame* ptr_fr, point_cloud* ptr_pcd) {
    fill_cloud(*static_cast<const CloudPayload*>(ptr_fr->image.payload), ptr_pcd);
}
}  // namespace cvo

namespace {

template <class M> void mat_to_rowmajor(const M& m, float* out) {
    for (int i = 0; i < m.rows(); ++i)
        for (int j = 0; j < m.cols(); ++j) out[i * m.cols() + j] = m(i, j);
}
double sparse_sum(const Eigen::SparseMatrix<float, Eigen::RowMajor>& A) {
    double s = 0.0;
    for (int i = 0; i < A.rows(); ++i)
        for (Eigen::SparseMatrix<float, Eigen::RowMajor>::InnerIterator it(A, i); it; ++it) s += (double)it.value();
    return s;
}

template <class Reg> struct Traits;
template <> struct Traits<cvo::cvo> {
    static constexpr bool adaptive = false;
};
template <> struct Traits<acvo::acvo> {
    static constexpr bool adaptive = true;
};

template <class Reg> void bind_pair(Reg& reg, const CloudPayload& x, const CloudPayload& y) {
    reg.set_pcd(1, cv::Mat(&x), cv::Mat(), "", "");  // first call: the fixed cloud (src/cvo.cpp:326-334)
    reg.set_pcd(1, cv::Mat(&y), cv::Mat(), "", "");  // second call: the moving cloud + sizes (:336-356)
}
template <class Reg> void set_state(Reg& reg, const float* R, const float* T, const float* ell) {
    if (R)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) reg.R(i, j) = R[3 * i + j];
    if (T)
        for (int i = 0; i < 3; ++i) reg.T(i) = T[i];
    if (ell) reg.ell = *ell;
}
template <class Reg> void get_state(Reg& reg, float* R, float* T, float* ell) {
    if (R) mat_to_rowmajor(reg.R, R);
    if (T)
        for (int i = 0; i < 3; ++i) T[i] = reg.T(i);
    if (ell) *ell = reg.ell;
}

void fill_rec(cvo::cvo& reg, float ell_used, oracle_trace_rec* r) {
    r->nnz = reg.A.nonZeros();
    r->sum_a = sparse_sum(reg.A);
    r->nnz_xx = r->nnz_yy = 0;
    r->dl = 0.0;
    r->ell = ell_used;
}
void fill_rec(acvo::acvo& reg, float ell_used, oracle_trace_rec* r) {
    r->nnz = reg.A.nonZeros();
    r->sum_a = sparse_sum(reg.A);
    r->nnz_xx = reg.Axx.nonZeros();
    r->nnz_yy = reg.Ayy.nonZeros();
    r->dl = reg.dl;
    r->ell = ell_used;
}
void ell_policy(cvo::cvo& reg, int k) {  // src/cvo.cpp:408-410
    reg.ell = (k > 2) ? 0.10 : reg.ell;
    reg.ell = (k > 9) ? 0.06 : reg.ell;
    reg.ell = (k > 19) ? 0.03 : reg.ell;
}
void ell_policy(acvo::acvo& reg, int) {  // src/adaptive_cvo.cpp:538-545
    reg.ell = reg.ell + reg.dl_step * reg.dl;
    if (reg.ell >= reg.ell_max) {
        reg.ell = reg.ell_max * 0.7;
        reg.ell_max = reg.ell_max * 0.7;
    }
    reg.ell = (reg.ell < reg.ell_min) ? reg.ell_min : reg.ell;
}
bool twist_small(cvo::cvo& reg) { return reg.omega.norm() < reg.eps && reg.v.norm() < reg.eps; }  // src/cvo.cpp:380
bool twist_small(acvo::acvo& reg) {                                                                 // src/adaptive_cvo.cpp:509
    return reg.omega.template cast<double>().norm() < reg.eps && reg.v.template cast<double>().norm() < reg.eps;
}

// align()'s loop (src/cvo.cpp:366-411, src/adaptive_cvo.cpp:495-546) re-driven from here, calling the reference's own
// private functions, so that every iteration can be recorded (mode 1) or the benchmark's fixed-ell schedule without
// stop tests can be run (mode 2: fixed_iters > 0, ell never changes).
template <class Reg>
void driven_loop(Reg& reg, int fixed_iters, int* iters, int* status, oracle_trace_rec* trace, int trace_cap, int* n_run) {
    const int max_iter = fixed_iters > 0 ? fixed_iters : reg.MAX_ITER;
    *iters = max_iter;
    *status = 0;
    int k = 0;
    for (; k < max_iter; ++k) {
        reg.update_tf();
        reg.transform_pcd();
        reg.compute_flow();
        reg.compute_step_size();
        oracle_trace_rec* rec = (trace && k < trace_cap) ? trace + k : nullptr;
        const float ell_used = reg.ell;
        if (rec) {
            std::memset(rec, 0, sizeof(*rec));
            rec->B = rec->C = rec->D = rec->E = std::numeric_limits<double>::quiet_NaN();  // locals of compute_step_size
            fill_rec(reg, ell_used, rec);
            rec->step = reg.step;
            for (int i = 0; i < 3; ++i) {
                rec->omega[i] = reg.omega(i);
                rec->v[i] = reg.v(i);
            }
        }
        bool stop = false;
        if (fixed_iters <= 0 && twist_small(reg)) {
            *iters = k;
            *status = 1;
            stop = true;
        }
        if (!stop) {
            Eigen::VectorXf vec_joined(reg.omega.size() + reg.v.size());
            vec_joined << reg.omega, reg.v;
            Eigen::MatrixXf dtrans = Exp_SEK3(vec_joined, reg.step);
            Eigen::Matrix3f dR = dtrans.block<3, 3>(0, 0);
            Eigen::Vector3f dT = dtrans.block<3, 1>(0, 3);
            reg.T = reg.R * dT + reg.T;
            reg.R = reg.R * dR;
            if (fixed_iters <= 0 && reg.dist_se3(dR, dT) < reg.eps_2) {
                *iters = k;
                *status = 2;
                stop = true;
            }
        }
        if (rec) {
            mat_to_rowmajor(reg.R, rec->R);
            for (int i = 0; i < 3; ++i) rec->T[i] = reg.T(i);
        }
        if (stop) {
            ++k;
            break;
        }
        if (fixed_iters <= 0) ell_policy(reg, k);
    }
    *n_run = k;
    // tail of align() (src/cvo.cpp:412-419)
    reg.prev_transform = reg.transform.matrix();
    reg.accum_transform = reg.accum_transform * reg.transform.matrix();
    reg.update_tf();
    reg.ptr_fixed_pcd = std::move(reg.ptr_moving_pcd);
    delete reg.cloud_y;
}

template <class Reg>
int run_align(const CloudPayload& x, const CloudPayload& y, int mode, int fixed_iters, int max_iter, float* R, float* T, float* ell,
              float* transform_out, float* prev_transform_out, int* iters_out, int* status_out, oracle_trace_rec* trace,
              int trace_cap, int* trace_len) {
    QuietCout quiet;
    Reg reg;
    if (max_iter > 0) const_cast<int&>(reg.MAX_ITER) = max_iter;
    bind_pair(reg, x, y);
    // set_pcd() of acvo re-arms ell (src/adaptive_cvo.cpp:476): carried-in values are applied afterwards
    set_state(reg, R, T, (ell && *ell > 0.f) ? ell : nullptr);
    int iters = -1, status = -1, n_run = -1;
    if (mode == 0) {
        reg.iter = -1;  // quirk Q5: `iter` keeps its previous value when MAX_ITER is hit (uninitialised in the reference)
        reg.align();
        iters = reg.iter;
    } else {
        driven_loop(reg, mode == 2 ? fixed_iters : 0, &iters, &status, trace, trace_cap, &n_run);
    }
    get_state(reg, R, T, ell);
    if (transform_out) mat_to_rowmajor(reg.transform.matrix(), transform_out);
    if (prev_transform_out) mat_to_rowmajor(reg.prev_transform.matrix(), prev_transform_out);
    if (iters_out) *iters_out = iters;
    if (status_out) *status_out = status;
    if (trace_len) *trace_len = n_run;
    return 0;
}

template <class Reg>
int run_eval(const CloudPayload& x, const CloudPayload& y, const float* R, const float* T, float ell, oracle_eval_out* out) {
    QuietCout quiet;
    Reg reg;
    bind_pair(reg, x, y);
    set_state(reg, R, T, &ell);
    reg.update_tf();
    reg.transform_pcd();
    reg.compute_flow();
    reg.compute_step_size();
    std::memset(out, 0, sizeof(*out));
    oracle_trace_rec rec;
    std::memset(&rec, 0, sizeof(rec));
    fill_rec(reg, ell, &rec);
    out->nnz = rec.nnz;
    out->sum_a = rec.sum_a;
    out->nnz_xx = rec.nnz_xx;
    out->nnz_yy = rec.nnz_yy;
    out->dl = rec.dl;
    out->B = out->C = out->D = out->E = std::numeric_limits<double>::quiet_NaN();
    out->dl_num = std::numeric_limits<double>::quiet_NaN();
    out->n_in_ball = -1;
    out->step = reg.step;
    for (int i = 0; i < 3; ++i) {
        out->omega[i] = reg.omega(i);
        out->v[i] = reg.v(i);
    }
    delete reg.cloud_y;
    return 0;
}

// The reference's driver loop (src/cvo_main.cpp:36-66) on ONE object: run_cvo() per frame, i.e. set_pcd() + align()
// with everything the object carries from pair to pair (quirks Q3 accum_transform, Q4 warm start, Q5 iter).
// xyz / feat: the frames' clouds back to back (counts[f] points each).  Per frame f: transform, accum_transform
// (row-major 4x4), iter, init.  Frame 0 only initialises (identity transforms).
template <class Reg>
int run_sequence(int n_frames, const float* xyz, const float* feat, const int* counts, float* transform_out, float* accum_out,
                 int* iter_out, float* ell_out) {
    QuietCout quiet;
    Reg reg;
    reg.iter = -1;
    size_t off = 0;
    for (int f = 0; f < n_frames; ++f) {
        const CloudPayload c{xyz + 3 * off, feat + NUM_FEATURES * off, counts[f]};
        reg.run_cvo(1, cv::Mat(&c), cv::Mat(), "", "");
        mat_to_rowmajor(reg.transform.matrix(), transform_out + 16 * f);
        mat_to_rowmajor(reg.accum_transform.matrix(), accum_out + 16 * f);
        iter_out[f] = reg.iter;
        ell_out[f] = reg.ell;
        off += counts[f];
    }
    return 0;
}

}  // namespace

extern "C" {

// kind: 0 = cvo::cvo, 1 = acvo::acvo.
// mode: 0 = the reference's own align(); 1 = the same loop driven from this file with a per-iteration trace;
//       2 = fixed ell (*ell), exactly fixed_iters iterations, no stop tests, no ell policy (benchmark config 2 / 5).
// max_iter > 0 overrides MAX_ITER.  *ell <= 0: keep the constructor's / set_pcd()'s length-scale.
int refsrc_align(int kind, const float* x_pos, const float* x_feat, int n_fixed, const float* y_pos, const float* y_feat, int n_moving,
                 int mode, int fixed_iters, int max_iter, float* R, float* T, float* ell, float* transform_out,
                 float* prev_transform_out, int* iters_out, int* status_out, oracle_trace_rec* trace, int trace_cap, int* trace_len) {
    const CloudPayload x{x_pos, x_feat, n_fixed}, y{y_pos, y_feat, n_moving};
    if (kind == 0)
        return run_align<cvo::cvo>(x, y, mode, fixed_iters, max_iter, R, T, ell, transform_out, prev_transform_out, iters_out,
                                   status_out, trace, trace_cap, trace_len);
    return run_align<acvo::acvo>(x, y, mode, fixed_iters, max_iter, R, T, ell, transform_out, prev_transform_out, iters_out,
                                 status_out, trace, trace_cap, trace_len);
}

// One pass of update_tf + transform_pcd + compute_flow + compute_step_size at (R, T, ell).  B..E are locals of the
// reference's compute_step_size and come back as NaN; `step` is what they produce.
int refsrc_eval(int kind, const float* x_pos, const float* x_feat, int n_fixed, const float* y_pos, const float* y_feat, int n_moving,
                const float* R, const float* T, float ell, oracle_eval_out* out) {
    const CloudPayload x{x_pos, x_feat, n_fixed}, y{y_pos, y_feat, n_moving};
    if (kind == 0) return run_eval<cvo::cvo>(x, y, R, T, ell, out);
    return run_eval<acvo::acvo>(x, y, R, T, ell, out);
}

int refsrc_run_sequence(int kind, int n_frames, const float* xyz, const float* feat, const int* counts, float* transform_out,
                        float* accum_out, int* iter_out, float* ell_out) {
    if (kind == 0) return run_sequence<cvo::cvo>(n_frames, xyz, feat, counts, transform_out, accum_out, iter_out, ell_out);
    return run_sequence<acvo::acvo>(n_frames, xyz, feat, counts, transform_out, accum_out, iter_out, ell_out);
}

// acvo::function_inner_product(cloud_a, cloud_b) (src/adaptive_cvo.cpp:385-439) at length-scale ell.
float refsrc_inner_product(const float* a_pos, const float* a_feat, int n_a, const float* b_pos, const float* b_feat, int n_b, float ell) {
    QuietCout quiet;
    acvo::acvo reg;
    reg.ell = ell;
    cvo::point_cloud a, b;
    fill_cloud(CloudPayload{a_pos, a_feat, n_a}, &a);
    fill_cloud(CloudPayload{b_pos, b_feat, n_b}, &b);
    return reg.function_inner_product(&a, &b);
}

// Exp_SEK3 (src/LieGroup.cpp:159-186), K = 1.
void refsrc_exp_sek3(const float* omega, const float* v, float dt, float* dR, float* dT) {
    Eigen::VectorXf x(6);
    x << omega[0], omega[1], omega[2], v[0], v[1], v[2];
    const Eigen::MatrixXf X = Exp_SEK3(x, dt);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) dR[3 * i + j] = X(i, j);
        dT[i] = X(i, 3);
    }
}

// poly_solver + root selection exactly as the tail of compute_step_size (src/cvo.cpp:53-69,291-307).  The selection
// loop is restated here (it lives in the middle of compute_step_size); poly_solver is the reference's.
float refsrc_step_from_coeffs(double B, double C, double D, double E, float min_step) {
    cvo::cvo reg;
    Eigen::VectorXf p_coef(4);
    p_coef << 4.0 * float(E), 3.0 * float(D), 2.0 * float(C), float(B);
    Eigen::VectorXcf rc = reg.poly_solver(p_coef);
    float temp_step = std::numeric_limits<float>::max();
    for (int i = 0; i < rc.real().size(); i++)
        if (rc(i, 0).real() > 0 && rc(i, 0).real() < temp_step && rc(i, 0).imag() == 0) temp_step = rc(i, 0).real();
    float step = temp_step == std::numeric_limits<float>::max() ? min_step : temp_step;
    step = step > 0.8 ? 0.8 : step;
    return step;
}

const char* refsrc_backend(void) {
    return "reference first-party sources (src/cvo.cpp, src/adaptive_cvo.cpp, src/LieGroup.cpp, thirdparty/nanoflann.hpp) "
           "compiled unmodified against oracle/shim (Eigen/TBB/OpenCV/PCL stand-ins)";
}
}

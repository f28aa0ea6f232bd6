/*
 * oracle/cvo_oracle.h -- C interface of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's
 * RKHS SE(3) registration hot path (cvo::align / acvo::align and their private
 * helpers).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load it.  The product (libcvo_b200.so) never links,
 * loads or calls anything in this directory.
 *
 * Parity status: the reference ships no golden vectors, KATs or tests for this
 * path and cannot be built in this image (Eigen3, TBB, OpenCV, PCL, Boost are
 * absent), so this oracle is pinned by (1) an independent NumPy float64
 * restatement (oracle/numpy_ref.py), (2) the reference's own nanoflann kd-tree
 * compiled from /root/reference (oracle/_ref, ball-query semantics), and
 * (3) algebraic invariants checked in tests/.  See DESIGN.md "Oracle".
 */
#ifndef CVO_ORACLE_H
#define CVO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_MODE_CVO = 0, ORACLE_MODE_ACVO = 1 };
enum { ORACLE_ELL_SCHEDULE = 0,   /* cvo: .10/.06/.03 after k>2/9/19 (src/cvo.cpp:408-410) */
       ORACLE_ELL_ADAPTIVE = 1,   /* acvo: ell += dl_step*dl ... (src/adaptive_cvo.cpp:538-545) */
       ORACLE_ELL_FIXED    = 2 }; /* bench config 2/5: ell never changes */

/* Every tunable the two reference constructors initialise
 * (src/cvo.cpp:18-48, src/adaptive_cvo.cpp:18-50). */
typedef struct oracle_params {
    int   mode;        /* ORACLE_MODE_*  (selects compute_flow variant)          */
    int   ell_policy;  /* ORACLE_ELL_*                                           */
    float ell_init;    /* cvo 0.15, acvo 0.1                                     */
    float ell_min;     /* acvo 0.0391                                            */
    float ell_max;     /* acvo 0.15                                              */
    double dl_step;    /* acvo 0.3 (double in inc/adaptive_cvo.hpp:78)           */
    float sigma;       /* 0.1                                                    */
    float sp_thres;    /* cvo 8e-3, acvo 8.315e-3                                */
    float c;           /* 7                                                      */
    float d;           /* 7                                                      */
    float c_ell;       /* cvo 200, acvo 0.5                                      */
    float c_sigma;     /* 1                                                      */
    float c_sp_thres;  /* acvo 8.315e-3; cvo uses sp_thres (src/cvo.cpp:103)     */
    int   max_iter;    /* 2000                                                   */
    float min_step;    /* 0.2                                                    */
    float max_step;    /* 0.8 (literal at src/cvo.cpp:307)                       */
    float eps;         /* 5e-5                                                   */
    float eps_2;       /* 1e-5                                                   */
    int   fixed_iters; /* >0: run exactly this many iterations, stop tests off   */
} oracle_params;

/* One evaluation of the hot path at a given (R, T, ell): what one outer
 * iteration computes before the pose update. */
typedef struct oracle_eval_out {
    long long nnz;         /* nonzeros of A (x,y)                                 */
    double    sum_a;       /* sum of A                                            */
    float     omega[3];
    float     v[3];
    double    B, C, D, E;
    float     step;
    /* acvo only */
    long long nnz_xx, nnz_yy;
    double    dl;          /* after division (src/adaptive_cvo.cpp:271)           */
    double    dl_num;      /* numerator before division                           */
    /* ball statistics */
    long long n_in_ball;   /* pairs with d2 < d2_thres                            */
} oracle_eval_out;

/* Per-iteration trace record (level-2 parity). */
typedef struct oracle_trace_rec {
    float     ell;
    float     step;
    float     omega[3];
    float     v[3];
    double    B, C, D, E;
    double    sum_a;
    double    dl;
    long long nnz, nnz_xx, nnz_yy;
    float     R[9];        /* state AFTER the update of this iteration (row-major) */
    float     T[3];
} oracle_trace_rec;

void oracle_default_params_cvo(oracle_params* p);
void oracle_default_params_acvo(oracle_params* p);

/* positions: n x 3 row-major f32; features: n x 5 row-major f32.
 * R row-major 3x3, T 3. Returns 0 on success. */
int oracle_eval(const float* x_pos, const float* x_feat, int n_fixed,
                const float* y_pos, const float* y_feat, int n_moving,
                const float* R, const float* T, float ell,
                const oracle_params* p, oracle_eval_out* out);

/* Runs align() (src/cvo.cpp:361-420 / src/adaptive_cvo.cpp:490-555).
 * R, T, ell: in = carried-in state (Q4 warm start), out = state at loop exit.
 * transform_out: row-major 4x4 `transform` as the reference leaves it after
 * align() (i.e. [R^T, -R^T T] of the FINAL R,T, src/cvo.cpp:415).
 * prev_transform_out: the stale transform multiplied into accum_transform
 * (quirk Q3, src/cvo.cpp:413-414). iters_out: k at exit (max_iter if the cap
 * was hit). status_out: 1 = stop-1 (twist), 2 = stop-2 (update), 0 = cap.
 * trace may be NULL; at most trace_cap records are written, *trace_len gets
 * the number of iterations executed. */
int oracle_align(const float* x_pos, const float* x_feat, int n_fixed,
                 const float* y_pos, const float* y_feat, int n_moving,
                 const oracle_params* p,
                 float* R, float* T, float* ell,
                 float* transform_out, float* prev_transform_out,
                 int* iters_out, int* status_out,
                 oracle_trace_rec* trace, int trace_cap, int* trace_len);

/* acvo::function_inner_product (src/adaptive_cvo.cpp:385-439): untransformed
 * clouds, colour gate from sp_thres. Returns sum_A/count as float; the parts
 * are returned too. */
float oracle_inner_product(const float* a_pos, const float* a_feat, int n_a,
                           const float* b_pos, const float* b_feat, int n_b,
                           float ell, const oracle_params* p,
                           double* sum_a, long long* count);

/* Exp_SEK3 for K=1 (src/LieGroup.cpp:159-186). dR row-major 3x3, dT 3. */
void oracle_exp_sek3(const float* omega, const float* v, float dt, float* dR, float* dT);

/* compute_step_size tail: cubic 4E t^3 + 3D t^2 + 2C t + B (src/cvo.cpp:291-307). */
float oracle_step_from_coeffs(double B, double C, double D, double E,
                              float min_step, float max_step);

/* Ball query used by se_kernel: indices j (ascending) with ||q - pts[j]||^2 < r2,
 * d2 formed like nanoflann's L2_Adaptor tail loop (tp/nanoflann.hpp:402-406).
 * Returns the count; writes at most cap indices/dists. Which backend answers
 * (brute force, or the reference's kd-tree when built as oracle/_ref) is
 * reported by oracle_backend(). */
int oracle_ball_query(const float* pts, int n, const float* q, float r2,
                      int* idx_out, float* d2_out, int cap);
const char* oracle_backend(void);
int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif

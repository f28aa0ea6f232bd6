// include/cvo_b200_frontend.hpp -- C++ host frontends above the C ABI (include/cvo_b200.h).
//
// cvo::cvo and acvo::acvo keep the PUBLIC surface of the reference classes
// (cpp/rkhs_registration/include/cvo.hpp:101-107,171-192 and include/adaptive_cvo.hpp:108-114,169-195):
//     bool init; int iter; transform, prev_transform, accum_transform;
//     set_pcd(...), align(), run_cvo(...), and for acvo function_inner_product(a, b)
// with the same observable behaviour, including the quirks of SURVEY.md section 8a:
//   Q3  accum_transform is multiplied by the transform of the TOP of the last iteration (src/cvo.cpp:413-414)
//   Q4  R, T are never reset between pairs; ell is never re-armed in cvo but is in acvo (src/adaptive_cvo.cpp:476-478)
//   Q5  `iter` is only assigned when a stop test fired (src/cvo.cpp:381,403)
// The reference's image overload set_pcd(dataset_seq, cv::Mat RGB, cv::Mat depth, ...) runs the image
// front end (pcd_generator + DSO PixelSelector2); this header offers both an
// image overload on plain pointers (the front end then runs on the device, cvo_b200_push_frame_images) and the
// raw-array overload that front end's OUTPUT feeds: N x 3 positions and N x 5 features.  INTEGRATION.md shows
// the three-line change that routes the reference's own set_pcd()/align() through this library.
//
// Header-only; needs only libcvo_b200.so.  No Eigen / OpenCV / PCL / TBB.
#ifndef CVO_B200_FRONTEND_HPP
#define CVO_B200_FRONTEND_HPP

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "cvo_b200.h"

namespace cvo_b200 {

// Minimal stand-in for Eigen::Affine3f (row-major 4x4).
struct Affine3f {
    float m[16];
    Affine3f() { set_identity(); }
    static Affine3f Identity() { return Affine3f(); }
    void set_identity() {
        std::memset(m, 0, sizeof(m));
        m[0] = m[5] = m[10] = m[15] = 1.f;
    }
    float& operator()(int r, int c) { return m[r * 4 + c]; }
    float operator()(int r, int c) const { return m[r * 4 + c]; }
    const float* data() const { return m; }
    Affine3f operator*(const Affine3f& o) const {
        Affine3f r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                float s = 0.f;
                for (int k = 0; k < 4; ++k) s += m[i * 4 + k] * o.m[k * 4 + j];
                r.m[i * 4 + j] = s;
            }
        return r;
    }
};

// The hot path's input contract: cvo::point_cloud without its pcl members (inc/data_type.h:59-71).
struct point_cloud {
    int num_points = 0;
    std::vector<float> positions;  // num_points x 3
    std::vector<float> features;   // num_points x 5, ROW-major (the reference's Eigen matrix is column-major)
};

namespace detail {

class registration {
  public:
    bool init = false;
    int iter = 0;
    Affine3f transform, prev_transform, accum_transform;

    explicit registration(bool adaptive, int device, int max_points) : adaptive_(adaptive) {
        if (adaptive) cvo_b200_default_params_acvo(&params_);
        else cvo_b200_default_params_cvo(&params_);
        if (cvo_b200_create(&ctx_, device, max_points, 2) != CVO_B200_OK)
            throw std::runtime_error("cvo_b200_create failed: no usable sm_100 device (there is no CPU fallback)");
        RT_[0] = RT_[4] = RT_[8] = 1.f;  // R = I, T = 0 (src/cvo.cpp:42-43)
        ell_ = params_.ell_init;
    }
    ~registration() { cvo_b200_destroy(ctx_); }
    registration(const registration&) = delete;
    registration& operator=(const registration&) = delete;

    cvo_b200_params& params() { return params_; }
    cvo_b200_ctx* context() { return ctx_; }

    // set_pcd (src/cvo.cpp:319-357): first call only stores the fixed cloud; later calls bind a new moving cloud.
    void set_pcd(const float* xyz, const float* feat, int n) {
        if (!init) {
            first_xyz_.assign(xyz, xyz + 3 * (size_t)n);
            first_feat_.assign(feat, feat + 5 * (size_t)n);
            first_n_ = n;
            init = true;
            return;
        }
        int rc;
        if (!pair_bound_) {
            rc = cvo_b200_set_pair(ctx_, 0, first_xyz_.data(), first_feat_.data(), first_n_, xyz, feat, n);
            pair_bound_ = (rc == CVO_B200_OK);
        } else if (aligned_since_set_) {
            rc = cvo_b200_push_frame(ctx_, 0, xyz, feat, n);  // fixed <- moving happened in align() (src/cvo.cpp:417)
        } else {
            // a second set_pcd() without an align() in between only replaces the moving cloud (src/cvo.cpp:336-351)
            rc = cvo_b200_replace_moving(ctx_, 0, xyz, feat, n);
        }
        check(rc);
        aligned_since_set_ = false;
        if (adaptive_) {  // src/adaptive_cvo.cpp:476-478
            ell_ = params_.ell_init;
        }
        have_moving_ = true;
    }
    void set_pcd(const point_cloud& pc) { set_pcd(pc.positions.data(), pc.features.data(), pc.num_points); }

    // set_pcd(dataset_seq, RGB, depth, ...) (src/cvo.cpp:319-357) with the image front end on the device
    // (pcd_generator::load_image + create_pointcloud: feature type 1 for cvo, src/cvo.cpp:329, 0 for acvo,
    // src/adaptive_cvo.cpp:451).  img3: height x width x 3 8-bit as cv::imread returns it (cv::Mat::data of a
    // continuous CV_8UC3), depth: height x width 16-bit (CV_16UC1).  Returns the number of points.
    int set_pcd(int dataset_seq, const unsigned char* img3, const unsigned short* depth, int width, int height) {
        int n = 0;
        if (prefetched_ == img3 && img3 != nullptr)  // the look-ahead of prefetch(): the cloud is (being) generated already
            check(cvo_b200_push_prefetched_frame(ctx_, 0, (!pair_bound_ || aligned_since_set_) ? 1 : 0, &n));
        else if (!pair_bound_ || aligned_since_set_)
            check(cvo_b200_push_frame_images(ctx_, 0, img3, depth, width, height, dataset_seq, adaptive_ ? 0 : 1, &n));
        else  // no align() since the last frame: only the moving cloud is replaced (src/cvo.cpp:336-351)
            check(cvo_b200_replace_moving_images(ctx_, 0, img3, depth, width, height, dataset_seq, adaptive_ ? 0 : 1, &n));
        prefetched_ = nullptr;
        aligned_since_set_ = false;
        if (!init) {
            init = true;
            return n;
        }
        pair_bound_ = true;
        if (adaptive_) ell_ = params_.ell_init;  // src/adaptive_cvo.cpp:476-478
        have_moving_ = true;
        return n;
    }
    // Look-ahead for a sequence loop (src/cvo_main.cpp:36-66 has frame k + 1 on disk while it aligns frame k): starts the
    // front end for the frame the NEXT set_pcd / run_cvo will be given (the same pointers, valid until then), so that
    // it overlaps the align() in between (cvo_b200_prefetch_frame_images).
    void prefetch(int dataset_seq, const unsigned char* img3, const unsigned short* depth, int width, int height) {
        check(cvo_b200_prefetch_frame_images(ctx_, img3, depth, width, height, dataset_seq, adaptive_ ? 0 : 1));
        prefetched_ = img3;
    }
    void run_cvo(int dataset_seq, const unsigned char* img3, const unsigned short* depth, int width, int height) {
        const bool first = !init;  // src/cvo.cpp:422-435
        set_pcd(dataset_seq, img3, depth, width, height);
        if (!first) align();
    }

    // align (src/cvo.cpp:361-420)
    void align() {
        if (!have_moving_) throw std::runtime_error("align() called before a moving cloud was set");
        const int slot = 0;
        int iters = 0, status = 0;
        check(cvo_b200_align(ctx_, &slot, 1, &params_, RT_, &ell_, transform.m, prev_transform.m, &iters, &status));
        if (status != CVO_B200_STATUS_MAX_ITER) iter = iters;      // Q5
        accum_transform = accum_transform * prev_transform;        // Q3 (src/cvo.cpp:413-414)
        last_status_ = status;
        have_moving_ = false;
        aligned_since_set_ = true;  // ptr_fixed_pcd = std::move(ptr_moving_pcd) (src/cvo.cpp:417): the next set_pcd promotes
    }

    // align() with the reference driver's look-ahead (src/cvo_main.cpp:36-66 has frame k + 1 on disk while frame k is
    // aligned): the kernel is launched, the front end of the NEXT frame is enqueued while it runs (prefetch), then the
    // result is collected.  The next set_pcd / run_cvo must be given the same image pointers.
    void align(int next_dataset_seq, const unsigned char* next_img3, const unsigned short* next_depth, int width, int height) {
        if (!have_moving_) throw std::runtime_error("align() called before a moving cloud was set");
        const int slot = 0;
        int iters = 0, status = 0;
        check(cvo_b200_align_begin(ctx_, &slot, 1, &params_, RT_, &ell_));
        prefetch(next_dataset_seq, next_img3, next_depth, width, height);
        check(cvo_b200_align_finish(ctx_, RT_, &ell_, transform.m, prev_transform.m, &iters, &status));
        if (status != CVO_B200_STATUS_MAX_ITER) iter = iters;      // Q5
        accum_transform = accum_transform * prev_transform;        // Q3 (src/cvo.cpp:413-414)
        last_status_ = status;
        have_moving_ = false;
        aligned_since_set_ = true;
    }

    void run_cvo(const float* xyz, const float* feat, int n) {  // src/cvo.cpp:422-435
        if (!init) {
            set_pcd(xyz, feat, n);
        } else {
            set_pcd(xyz, feat, n);
            align();
        }
    }
    void run_cvo(const point_cloud& pc) { run_cvo(pc.positions.data(), pc.features.data(), pc.num_points); }

    int last_status() const { return last_status_; }
    float ell() const { return ell_; }

  protected:
    void check(int rc) {
        if (rc != CVO_B200_OK) throw std::runtime_error(std::string("libcvo_b200: ") + cvo_b200_last_error(ctx_));
    }
    bool adaptive_;
    cvo_b200_params params_;
    cvo_b200_ctx* ctx_ = nullptr;
    float RT_[12] = {0};
    float ell_ = 0.f;
    std::vector<float> first_xyz_, first_feat_;
    int first_n_ = 0;
    bool pair_bound_ = false, have_moving_ = false, aligned_since_set_ = false;
    const unsigned char* prefetched_ = nullptr;  // image of the frame whose front end was started by prefetch()
    int last_status_ = 0;
};

}  // namespace detail
}  // namespace cvo_b200

namespace cvo {
using cvo_b200::Affine3f;
using cvo_b200::point_cloud;
class cvo : public cvo_b200::detail::registration {
  public:
    explicit cvo(int device = 0, int max_points = 16384) : registration(false, device, max_points) {}
};
}  // namespace cvo

namespace acvo {
using cvo_b200::Affine3f;
using cvo_b200::point_cloud;
class acvo : public cvo_b200::detail::registration {
  public:
    explicit acvo(int device = 0, int max_points = 16384) : registration(true, device, max_points) {}

    // function_inner_product(cloud_a, cloud_b) (src/adaptive_cvo.cpp:385-439): mean surviving kernel value of two
    // UNtransformed clouds at the current ell.
    float function_inner_product(const point_cloud* a, const point_cloud* b) {
        check(cvo_b200_set_pair(ctx_, 1, a->positions.data(), a->features.data(), a->num_points,
                                b->positions.data(), b->features.data(), b->num_points));
        float value = 0.f;
        check(cvo_b200_inner_product(ctx_, 1, ell_, &params_, &value, nullptr, nullptr));
        return value;
    }
};
}  // namespace acvo

#endif

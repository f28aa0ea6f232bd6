// examples/frontend_example.cpp -- the reference's driver loop (src/cvo_main.cpp:36-66) on the C++ frontends of
// include/cvo_b200_frontend.hpp: a short synthetic "sequence" of three frames is registered frame to frame with
// cvo::cvo and acvo::acvo; prints the per-pair transform and the accumulated pose, returns non-zero if the known
// inter-frame motion is not recovered.  Built by __graft_entry__.build(); run by tests/test_gpu_frontend.py.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "../include/cvo_b200_frontend.hpp"

namespace {

struct Frame {
    std::vector<float> xyz, feat_raw, feat_norm;
    int n;
};

// points on a desk plane, a back wall and a box edge pattern, seen from a camera translated by (tx, 0, tz) and
// yawed by `yaw`; colours piecewise constant per surface.
Frame make_frame(int n, float tx, float tz, float yaw, unsigned seed) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::normal_distribution<float> N(0.f, 1.f);
    Frame f;
    f.n = n;
    const float c = std::cos(yaw), s = std::sin(yaw);
    for (int i = 0; i < n; ++i) {
        float p[3], bgr[3];
        const int surf = i % 3;
        if (surf == 0) {  // desk
            p[0] = -0.6f + 1.2f * U(rng); p[1] = 0.35f; p[2] = 0.9f + 0.7f * U(rng);
            if (i % 2) p[0] = std::floor(p[0] * 8.f) / 8.f;  // texture lines
            bgr[0] = 60; bgr[1] = 120; bgr[2] = 200;
        } else if (surf == 1) {  // wall
            p[0] = -0.6f + 1.2f * U(rng); p[1] = -0.4f + 0.75f * U(rng); p[2] = 1.7f;
            if (i % 2) p[1] = std::floor(p[1] * 8.f) / 8.f;
            bgr[0] = 200; bgr[1] = 180; bgr[2] = 40;
        } else {  // box front
            p[0] = -0.1f + 0.3f * U(rng); p[1] = 0.1f + 0.25f * U(rng); p[2] = 1.2f;
            bgr[0] = 90; bgr[1] = 40; bgr[2] = 150;
        }
        // world -> camera: R^T (p - t)
        const float qx = p[0] - tx, qz = p[2] - tz;
        const float x = c * qx - s * qz, z = s * qx + c * qz;
        f.xyz.push_back(x + 0.001f * N(rng));
        f.xyz.push_back(p[1] + 0.001f * N(rng));
        f.xyz.push_back(z + 0.001f * N(rng));
        for (int k = 0; k < 3; ++k) {
            const float v = bgr[k] + 4.f * N(rng);
            f.feat_raw.push_back(v);
            f.feat_norm.push_back(v / 255.f);
        }
        for (int k = 0; k < 2; ++k) {
            const float g = 20.f * N(rng);
            f.feat_raw.push_back(g);
            f.feat_norm.push_back(g * 2.f / 255.f);
        }
    }
    return f;
}

void print44(const char* name, const cvo_b200::Affine3f& T) {
    std::printf("%s\n", name);
    for (int r = 0; r < 4; ++r) std::printf("  % .6f % .6f % .6f % .6f\n", T(r, 0), T(r, 1), T(r, 2), T(r, 3));
}

}  // namespace

int main() {
    const int n = 2000;
    // camera moves +1 cm in x, +0.5 cm in z and yaws 0.5 deg per frame
    std::vector<Frame> frames;
    for (int k = 0; k < 3; ++k) frames.push_back(make_frame(n, 0.01f * k, 0.005f * k, 0.0087f * k, 100 + k));
    int bad = 0;
    try {
        cvo::cvo reg;
        for (int k = 0; k < 3; ++k) {
            reg.run_cvo(frames[k].xyz.data(), frames[k].feat_raw.data(), frames[k].n);  // src/cvo_main.cpp:52
            if (k == 0) continue;
            std::printf("cvo pair %d: iter=%d status=%d ell=%.3f\n", k, reg.iter, reg.last_status(), reg.ell());
            print44(" transform", reg.transform);
            if (std::fabs(reg.transform(0, 3) - 0.01f) > 4e-3f || std::fabs(reg.transform(2, 3) - 0.005f) > 4e-3f) ++bad;
        }
        print44("cvo accum_transform", reg.accum_transform);

        acvo::acvo areg;
        for (int k = 0; k < 3; ++k) {
            areg.run_cvo(frames[k].xyz.data(), frames[k].feat_norm.data(), frames[k].n);
            if (k == 0) continue;
            std::printf("acvo pair %d: iter=%d status=%d ell=%.4f\n", k, areg.iter, areg.last_status(), areg.ell());
            if (std::fabs(areg.transform(0, 3) - 0.01f) > 4e-3f || std::fabs(areg.transform(2, 3) - 0.005f) > 4e-3f) ++bad;
        }
        print44("acvo accum_transform", areg.accum_transform);
        cvo_b200::point_cloud a, b;
        a.num_points = b.num_points = n;
        a.positions = frames[0].xyz; a.features = frames[0].feat_norm;
        b.positions = frames[1].xyz; b.features = frames[1].feat_norm;
        const float ip = areg.function_inner_product(&a, &b);
        std::printf("acvo function_inner_product(frame0, frame1) = %.6f\n", ip);
        if (!(ip > 8.315e-3f && ip < 1e-2f)) ++bad;  // mean of values in (sp_thres, sigma^2]
    } catch (const std::exception& e) {
        std::fprintf(stderr, "frontend_example: %s\n", e.what());
        return 2;
    }
    std::printf(bad ? "FAILED (%d checks)\n" : "OK\n", bad);
    return bad ? 1 : 0;
}

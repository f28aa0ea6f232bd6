// examples/cvo_batch_multi_gpu.cpp -- the native multi-GPU batch driver above the C ABI (BASELINE config 4).
//
// The reference registers one pair at a time on the host (src/cvo_main.cpp:36-66).  Independent pairs shard trivially:
// this driver creates one context per visible GPU and hands a batch of frame pairs to cvo_b200_align_multi, which deals
// pair q to GPU q mod W, uploads and aligns every share on its own host thread and writes the 4x4 poses back at index q.
//
//   cvo_batch_multi_gpu [n_pairs = 64] [points = 3000] [n_gpus = all]
//
// Input here is synthetic (two noisy samplings of one random box scene per pair, the second moved by a small rigid
// motion); a caller with real data fills the same arrays from its point clouds (include/cvo_b200_io.hpp reads PCD files).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <cuda_runtime_api.h>

#include "../include/cvo_b200.h"

int main(int argc, char** argv) {
    const int n_pairs = argc > 1 ? std::atoi(argv[1]) : 64;
    const int n = argc > 2 ? std::atoi(argv[2]) : 3000;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) {
        std::fprintf(stderr, "no CUDA device (there is no CPU fallback)\n");
        return 1;
    }
    const int W = argc > 3 ? std::min(std::atoi(argv[3]), n_dev) : n_dev;
    std::vector<float> fx((size_t)n_pairs * n * 3), ff((size_t)n_pairs * n * 5), mx(fx.size()), mf(ff.size());
    std::vector<int> cnt(n_pairs, n);
    std::mt19937 rng(12345);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::normal_distribution<float> N(0.f, 1.f);
    std::vector<float> truth((size_t)n_pairs * 3);
    for (int q = 0; q < n_pairs; ++q) {
        const float tx = 0.01f * N(rng), ty = 0.01f * N(rng), tz = 0.01f * N(rng), yaw = 0.01f * N(rng);
        truth[3 * q] = tx; truth[3 * q + 1] = ty; truth[3 * q + 2] = tz;
        for (int cloud = 0; cloud < 2; ++cloud) {
            float* X = (cloud ? mx : fx).data() + (size_t)q * n * 3;
            float* F = (cloud ? mf : ff).data() + (size_t)q * n * 5;
            for (int i = 0; i < n; ++i) {
                // points on the edges of a few boxes: CVO's inputs are high-gradient pixels
                const int edge = (int)(U(rng) * 12) % 12, box = (int)(U(rng) * 4) % 4;
                const float t = U(rng), s = 0.25f + 0.1f * box;
                float p[3] = {(edge & 1) ? s : -s, (edge & 2) ? s : -s, (edge & 4) ? s : -s};
                p[edge % 3] = s * (2.f * t - 1.f);
                float x = p[0] + 0.3f * box - 0.4f, y = p[1] * 0.6f, z = 1.3f + p[2] * 0.5f;
                x += 0.002f * N(rng); y += 0.002f * N(rng); z += 0.002f * N(rng);
                if (cloud) {  // the moving cloud: the same scene seen after a small motion
                    const float c = std::cos(yaw), sn = std::sin(yaw);
                    const float xr = c * x - sn * y, yr = sn * x + c * y;
                    x = xr - tx; y = yr - ty; z = z - tz;
                }
                X[3 * i] = x; X[3 * i + 1] = y; X[3 * i + 2] = z;
                const float base = 40.f + 50.f * box;
                F[5 * i] = base + 8.f * N(rng); F[5 * i + 1] = base * 0.7f + 8.f * N(rng); F[5 * i + 2] = 200.f - base + 8.f * N(rng);
                F[5 * i + 3] = 20.f * N(rng); F[5 * i + 4] = 20.f * N(rng);
            }
        }
    }
    std::vector<cvo_b200_ctx*> ctx(W, nullptr);
    const int slots = (n_pairs + W - 1) / W;
    for (int d = 0; d < W; ++d)
        if (cvo_b200_create(&ctx[d], d, n, slots) != CVO_B200_OK) {
            std::fprintf(stderr, "cvo_b200_create failed on device %d\n", d);
            return 1;
        }
    cvo_b200_params p;
    cvo_b200_default_params_cvo(&p);
    std::vector<float> tf((size_t)n_pairs * 16), ms(W);
    std::vector<int> iters(n_pairs), status(n_pairs);
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = cvo_b200_align_multi(ctx.data(), W, n_pairs, fx.data(), ff.data(), cnt.data(), mx.data(), mf.data(), cnt.data(), n,
                                            &p, tf.data(), iters.data(), status.data(), ms.data());
        if (rc != CVO_B200_OK) {
            std::fprintf(stderr, "cvo_b200_align_multi: %d (%s)\n", rc, cvo_b200_last_error(ctx[0]));
            return 1;
        }
        best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    int converged = 0;
    double iter_sum = 0;
    for (int q = 0; q < n_pairs; ++q) {
        converged += status[q] == CVO_B200_STATUS_CONVERGED_TWIST || status[q] == CVO_B200_STATUS_CONVERGED_UPDATE;
        iter_sum += iters[q];
    }
    std::printf("%d pairs of %d x %d points on %d GPU(s): %.1f pairs/s end to end (host arrays -> poses), %d converged, %.1f iterations on average\n",
                n_pairs, n, n, W, n_pairs / best, converged, iter_sum / n_pairs);
    for (int d = 0; d < W; ++d) {
        std::printf("  GPU %d: align kernels %.2f ms\n", d, ms[d]);
        cvo_b200_destroy(ctx[d]);
    }
    std::puts(converged == n_pairs ? "OK" : "NOT ALL CONVERGED");
    return converged == n_pairs ? 0 : 1;
}

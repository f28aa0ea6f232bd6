// examples/cvo_sequence.cpp -- the reference's sequence driver (src/cvo_main.cpp / src/adaptive_cvo_main.cpp) over
// point-cloud files instead of TUM images: for every line of assoc.txt it loads <folder>/pcd_ds/<rgb_name>.pcd,
// sub-samples it to at most `num_want` points (the reference selects num_want = 3000, src/pcd_generator.cpp:22),
// calls run_cvo() and appends `name tx ty tz qx qy qz qw` of accum_transform to <folder>/{cvo,acvo}_poses_qt.txt.
//
//   cvo_sequence <folder/> <cvo|acvo> [num_want=3000] [max_frames=0 (all)]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <numeric>
#include <random>

#include "../include/cvo_b200_io.hpp"

template <class Reg>
int run(Reg& reg, const std::string& folder, const std::string& out_name, bool adaptive, int num_want, int max_frames) {
    const auto assoc = cvo_b200::read_assoc(folder + "assoc.txt");
    std::cout << "num images: " << assoc.size() << std::endl;
    cvo_b200::pose_writer poses(folder + out_name);
    const auto t0 = std::chrono::steady_clock::now();
    int done = 0;
    for (size_t i = 0; i < assoc.size() && (max_frames <= 0 || (int)i < max_frames); ++i) {
        cvo_b200::pcd_cloud pc = cvo_b200::read_pcd_ascii(folder + "pcd_ds/" + assoc[i].rgb_name + ".pcd");
        if (pc.n > num_want) {  // seeded sub-sample, a pure function of the frame index
            std::vector<int> idx(pc.n);
            std::iota(idx.begin(), idx.end(), 0);
            std::mt19937 rng(3141592u + (unsigned)i);
            std::shuffle(idx.begin(), idx.end(), rng);
            idx.resize(num_want);
            std::sort(idx.begin(), idx.end());
            cvo_b200::pcd_cloud sub;
            for (int k : idx) {
                for (int a = 0; a < 3; ++a) {
                    sub.xyz.push_back(pc.xyz[3 * k + a]);
                    sub.rgb.push_back(pc.rgb[3 * k + a]);
                }
            }
            sub.n = num_want;
            pc = sub;
        }
        const bool had_fixed = reg.init;
        reg.run_cvo(cvo_b200::make_point_cloud(pc, adaptive));
        if (had_fixed)
            std::cout << "frame " << i << ": " << assoc[i - 1].rgb_name << " -> " << assoc[i].rgb_name << "  iterations "
                      << reg.iter << std::endl;
        // the reference writes accum_transform for EVERY frame once cvo.init is set, the identity line of frame 0
        // included (src/cvo_main.cpp:58-65)
        if (reg.init) poses.write(assoc[i].rgb_name, reg.accum_transform);
        ++done;
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "Total time for " << done << " frames is: " << s << " s" << std::endl;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <folder/> <cvo|acvo> [num_want] [max_frames]\n", argv[0]);
        return 2;
    }
    const std::string folder = argv[1], kind = argv[2];
    const int num_want = argc > 3 ? std::atoi(argv[3]) : 3000;
    const int max_frames = argc > 4 ? std::atoi(argv[4]) : 0;
    try {
        if (kind == "acvo") {
            acvo::acvo reg(0, 16384);
            return run(reg, folder, "acvo_poses_qt.txt", true, num_want, max_frames);
        }
        cvo::cvo reg(0, 16384);
        return run(reg, folder, "cvo_poses_qt.txt", false, num_want, max_frames);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}

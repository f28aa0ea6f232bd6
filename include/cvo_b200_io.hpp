// include/cvo_b200_io.hpp -- file front door of the frontends, header-only C++17.
//
// The reference's drivers read TUM images and run its image front-end (pcd_generator, out of scope).  This header
// gives the same drivers a way in from point-cloud files instead:
//   read_pcd_ascii   the reference's sample clouds data/rgbd_dataset/freiburg1_desk/pcd_ds/*.pcd
//                    (FIELDS x y z rgb, SIZE 8 8 8 4, DATA ascii, colour packed into a float's bit pattern)
//   make_features    N x 5 feature rows of the two flavours (src/pcd_generator.cpp:336-381); gradients are 0 for a PCD
//   read_assoc       TUM assoc.txt exactly as load_file_name() reads it (src/cvo_main.cpp:75-101)
//   pose_writer      `name tx ty tz qx qy qz qw` of accum_transform per aligned frame (src/cvo_main.cpp:58-65)
#ifndef CVO_B200_IO_HPP
#define CVO_B200_IO_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cvo_b200_frontend.hpp"

namespace cvo_b200 {

struct pcd_cloud {
    std::vector<float> xyz;         // n x 3
    std::vector<unsigned char> rgb; // n x 3, R G B
    int n = 0;
};

inline pcd_cloud read_pcd_ascii(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::vector<std::string> fields;
    long points = -1;
    std::string line;
    bool data = false;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string key;
        if (!(ss >> key) || key[0] == '#') continue;
        if (key == "FIELDS") {
            std::string f;
            while (ss >> f) fields.push_back(f);
        } else if (key == "POINTS") {
            ss >> points;
        } else if (key == "DATA") {
            std::string kind;
            ss >> kind;
            if (kind != "ascii") throw std::runtime_error(path + ": only DATA ascii is supported");
            data = true;
            break;
        }
    }
    if (!data || fields.size() < 3 || fields[0] != "x" || fields[1] != "y" || fields[2] != "z")
        throw std::runtime_error(path + ": not a PCD file with FIELDS x y z ...");
    const int rgb_at = (int)(std::find(fields.begin(), fields.end(), "rgb") - fields.begin());
    pcd_cloud pc;
    long seen = 0;
    while (std::getline(in, line) && (points < 0 || seen < points)) {
        std::istringstream ss(line);
        std::vector<std::string> tok;
        std::string t;
        while (ss >> t) tok.push_back(t);
        if (tok.size() < 3) continue;
        ++seen;
        const double x = std::strtod(tok[0].c_str(), nullptr), y = std::strtod(tok[1].c_str(), nullptr),
                     z = std::strtod(tok[2].c_str(), nullptr);
        if (!(std::isfinite(x) && std::isfinite(y) && std::isfinite(z))) continue;  // invalid depth
        pc.xyz.push_back((float)x);
        pc.xyz.push_back((float)y);
        pc.xyz.push_back((float)z);
        std::uint32_t bits = 0;
        if (rgb_at < (int)fields.size() && rgb_at < (int)tok.size()) {
            const float packed = std::strtof(tok[rgb_at].c_str(), nullptr);  // 0x00RRGGBB in the float's bits
            std::memcpy(&bits, &packed, 4);
        }
        pc.rgb.push_back((unsigned char)((bits >> 16) & 255));
        pc.rgb.push_back((unsigned char)((bits >> 8) & 255));
        pc.rgb.push_back((unsigned char)(bits & 255));
    }
    pc.n = (int)(pc.xyz.size() / 3);
    return pc;
}

// 8-bit HSV of cv::cvtColor(COLOR_RGB2HSV) with (c0, c1, c2) taken as (R, G, B)
inline void hsv_u8(double r, double g, double b, double& h, double& s, double& v) {
    v = std::max(r, std::max(g, b));
    const double mn = std::min(r, std::min(g, b)), d = v - mn;
    s = v > 0 ? 255.0 * d / v : 0.0;
    if (d == 0) h = 0;
    else if (v == r) h = 60.0 * (g - b) / d;
    else if (v == g) h = 120.0 + 60.0 * (b - r) / d;
    else h = 240.0 + 60.0 * (r - g) / d;
    if (h < 0) h += 360.0;
    h = std::fmod(std::nearbyint(h / 2.0), 180.0);
    s = std::nearbyint(s);
}

// adaptive = false: feature_type 1 (cvo): B, G, R raw, gradients 0.  adaptive = true: feature_type 0 (acvo):
// H/180, S/255, V/255 where the reference feeds its BGR image to COLOR_RGB2HSV (src/pcd_generator.cpp:389) -- reproduced.
inline point_cloud make_point_cloud(const pcd_cloud& pc, bool adaptive) {
    point_cloud out;
    out.num_points = pc.n;
    out.positions = pc.xyz;
    out.features.assign((size_t)pc.n * 5, 0.f);
    for (int i = 0; i < pc.n; ++i) {
        const double r = pc.rgb[3 * i], g = pc.rgb[3 * i + 1], b = pc.rgb[3 * i + 2];
        float* f = &out.features[(size_t)5 * i];
        if (!adaptive) {
            f[0] = (float)b; f[1] = (float)g; f[2] = (float)r;
        } else {
            double h, s, v;
            hsv_u8(b, g, r, h, s, v);
            f[0] = (float)(h / 180.0); f[1] = (float)(s / 255.0); f[2] = (float)(v / 255.0);
        }
    }
    return out;
}

struct assoc_entry {
    std::string rgb_name, rgb_path, depth_name, depth_path;
};

inline std::vector<assoc_entry> read_assoc(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::vector<assoc_entry> out;
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        std::istringstream ss(line);
        assoc_entry e;
        if (!(ss >> e.rgb_name)) continue;
        ss >> e.rgb_path >> e.depth_name >> e.depth_path;
        out.push_back(e);
    }
    return out;
}

// Eigen::Quaternionf(Matrix3f) branch structure; returns x, y, z, w
inline void rotation_to_quaternion(const Affine3f& T, double q[4]) {
    const double m[3][3] = {{T(0, 0), T(0, 1), T(0, 2)}, {T(1, 0), T(1, 1), T(1, 2)}, {T(2, 0), T(2, 1), T(2, 2)}};
    double t = m[0][0] + m[1][1] + m[2][2];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[2][1] - m[1][2]) * t;
        q[1] = (m[0][2] - m[2][0]) * t;
        q[2] = (m[1][0] - m[0][1]) * t;
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[k][j] - m[j][k]) * t;
        q[j] = (m[j][i] + m[i][j]) * t;
        q[k] = (m[k][i] + m[i][k]) * t;
    }
}

class pose_writer {
  public:
    explicit pose_writer(const std::string& path) : out_(path) {
        if (!out_) throw std::runtime_error("cannot open " + path);
        out_.precision(9);
    }
    void write(const std::string& name, const Affine3f& accum) {
        double q[4];
        rotation_to_quaternion(accum, q);
        out_ << name << " " << accum(0, 3) << " " << accum(1, 3) << " " << accum(2, 3) << " " << q[0] << " " << q[1]
             << " " << q[2] << " " << q[3] << "\n";
    }

  private:
    std::ofstream out_;
};

}  // namespace cvo_b200
#endif

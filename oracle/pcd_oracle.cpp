// oracle/pcd_oracle.cpp -- CPU restatement of the reference's image front end (TEST INFRASTRUCTURE ONLY).
//
// Restates, function by function, what turns an 8-bit colour image + a 16-bit depth image into a cvo::point_cloud
// in MaaniGhaffari/cvo-rgbd (paths relative to /root/reference/cpp/rkhs_registration/):
//   pcd_generator::load_image            src/pcd_generator.cpp:384-396   (cv::cvtColor RGB2GRAY / RGB2HSV)
//   pcd_generator::make_pyramid          src/pcd_generator.cpp:33-120
//   dso::PixelSelector (ctor)            thirdparty/PixelSelector2.cpp:34-49   (srand(3141592); rand() & 0xFF)
//   dso::PixelSelector::makeHists        thirdparty/PixelSelector2.cpp:71-136
//   dso::PixelSelector::makeMaps         thirdparty/PixelSelector2.cpp:137-282
//   dso::PixelSelector::select           thirdparty/PixelSelector2.cpp:286-435
//   pcd_generator::select_point          src/pcd_generator.cpp:122-164   (incl. the Canny top-up, cv::blur + cv::Canny restated)
//   pcd_generator::get_points_from_pixels src/pcd_generator.cpp:233-327
//   pcd_generator::get_features          src/pcd_generator.cpp:329-382
// Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may load this library; the product path never does.
//
// OpenCV is a third-party dependency of the reference that is absent from /root/reference and unpinned ("OpenCV >= 3",
// CMakeLists.txt:46).  The two colour conversions are restated here from OpenCV 4.x's published 8-bit fixed-point
// formulas (modules/imgproc/src/color_rgb / color_hsv: 15-bit luma weights 9798/19235/3735; HSV with hsv_shift = 12
// and the sdiv / hdiv180 tables), cv::blur(3x3) and cv::Canny(0, 25, aperture 3, L1 gradient) from OpenCV's published
// algorithm (modules/imgproc/src/canny.cpp: Sobel with replicated border, 15-bit tan(22.5 deg) sector test with its
// asymmetric > / >= comparisons, 8-connected hysteresis); all four are PINNED against the cv2 4.13 of this image in
// tests/test_pcd_oracle.py.
//
// Two reads of uninitialised heap memory in the reference are given a defined value here (and in the CUDA path):
//   U1  abs_squared_grad[l] is `new float[]` and its first and last image rows are never written
//       (src/pcd_generator.cpp:42,94-109); select() reads level 2 row h/4-1 for pixels of row h-4
//       (thirdparty/PixelSelector2.cpp:395).  Defined as 0 (the pixel can then never be a level-3 pick).
//   U2  dI_pyr[l][.][1..2] of the same border rows: never read for selected pixels (4 <= y <= h-4).
// The threshold lookup thsSmoothed[(x>>5) + (y>>5)*thsStep] (PixelSelector2.cpp:367) is only in range when the image
// size is a multiple of 32; other sizes are refused.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Pyr {
    int w[3], h[3];
    std::vector<float> I[3], dx[3], dy[3], g2[3];  // dI[.][0], dI[.][1], dI[.][2], abs_squared_grad
};

// src/pcd_generator.cpp:33-120
void make_pyramid(const uint8_t* gray, int w, int h, Pyr& p) {
    int wl = w, hl = h;
    for (int l = 0; l < 3; ++l) {
        p.w[l] = wl; p.h[l] = hl;
        p.I[l].assign((size_t)wl * hl, 0.f);
        p.dx[l].assign((size_t)wl * hl, 0.f);   // U2: defined as 0
        p.dy[l].assign((size_t)wl * hl, 0.f);
        p.g2[l].assign((size_t)wl * hl, 0.f);   // U1: defined as 0
        wl /= 2; hl /= 2;
    }
    for (int i = 0; i < w * h; ++i) p.I[0][i] = gray[i];  // :55-59
    for (int l = 0; l < 3; ++l) {
        wl = p.w[l]; hl = p.h[l];
        if (l > 0) {  // :78-92
            const int pw = wl * 2;
            const float* q = p.I[l - 1].data();
            for (int y = 0; y < hl; ++y)
                for (int x = 0; x < wl; ++x)
                    p.I[l][x + y * wl] = 0.25f * (q[2 * x + 2 * y * pw] + q[2 * x + 1 + 2 * y * pw] + q[2 * x + 2 * y * pw + pw] +
                                                  q[2 * x + 1 + 2 * y * pw + pw]);
        }
        const float* I = p.I[l].data();
        for (int idx = wl; idx < wl * (hl - 1); ++idx) {  // :96-112 (row ends wrap into the neighbouring rows, as there)
            float dx = 0.5f * (I[idx + 1] - I[idx - 1]);
            float dy = 0.5f * (I[idx + wl] - I[idx - wl]);
            if (!std::isfinite(dx)) dx = 0;
            if (!std::isfinite(dy)) dy = 0;
            p.dx[l][idx] = dx;
            p.dy[l][idx] = dy;
            p.g2[l][idx] = dx * dx + dy * dy;
        }
    }
}

// thirdparty/PixelSelector2.cpp:58-67
int hist_quantile(const int* hist, float below) {
    int th = (int)(hist[0] * below + 0.5f);
    for (int i = 0; i < 90; ++i) {
        th -= hist[i + 1];
        if (th < 0) return i;
    }
    return 90;
}

struct Selector {
    int w, h, w32, h32;
    std::vector<unsigned char> randomPattern;
    std::vector<float> ths, thsSmoothed;
    int currentPotential = 3;

    Selector(int w_, int h_) : w(w_), h(h_), w32(w_ / 32), h32(h_ / 32) {  // :34-49
        randomPattern.resize((size_t)w * h);
        std::srand(3141592);
        for (int i = 0; i < w * h; ++i) randomPattern[i] = std::rand() & 0xFF;
        ths.assign((size_t)w32 * h32 + 100, 0.f);
        thsSmoothed.assign((size_t)w32 * h32 + 100, 0.f);
    }

    void makeHists(const Pyr& p) {  // :71-136
        const float* mapmax0 = p.g2[0].data();
        for (int y = 0; y < h32; ++y)
            for (int x = 0; x < w32; ++x) {
                const float* map0 = mapmax0 + 32 * x + 32 * y * w;
                int hist0[100];
                memset(hist0, 0, sizeof(hist0));
                for (int j = 0; j < 32; ++j)
                    for (int i = 0; i < 32; ++i) {
                        const int it = i + 32 * x, jt = j + 32 * y;
                        if (it > w - 2 || jt > h - 2 || it < 1 || jt < 1) continue;
                        int g = (int)sqrtf(map0[i + j * w]);
                        if (g > 48) g = 48;
                        hist0[g + 1]++;
                        hist0[0]++;
                    }
                ths[x + y * w32] = (float)(hist_quantile(hist0, 0.5f) + 7);  // setting_minGradHistCut / Add
            }
        for (int y = 0; y < h32; ++y)
            for (int x = 0; x < w32; ++x) {
                float sum = 0, num = 0;
                if (x > 0) {
                    if (y > 0) { num++; sum += ths[x - 1 + (y - 1) * w32]; }
                    if (y < h32 - 1) { num++; sum += ths[x - 1 + (y + 1) * w32]; }
                    num++; sum += ths[x - 1 + y * w32];
                }
                if (x < w32 - 1) {
                    if (y > 0) { num++; sum += ths[x + 1 + (y - 1) * w32]; }
                    if (y < h32 - 1) { num++; sum += ths[x + 1 + (y + 1) * w32]; }
                    num++; sum += ths[x + 1 + y * w32];
                }
                if (y > 0) { num++; sum += ths[x + (y - 1) * w32]; }
                if (y < h32 - 1) { num++; sum += ths[x + (y + 1) * w32]; }
                num++; sum += ths[x + y * w32];
                thsSmoothed[x + y * w32] = (sum / num) * (sum / num);
            }
    }

    // :286-435 with setting_selectDirectionDistribution == false (dirNorm = the gradient magnitude itself, so the
    // random direction, and with it the dependence on n2, drops out)
    void select(const Pyr& p, float* map_out, int pot, float thFactor, int n[3]) {
        const float* mapmax0 = p.g2[0].data();
        const float* mapmax1 = p.g2[1].data();
        const float* mapmax2 = p.g2[2].data();
        const int w1 = w / 2, w2 = w / 4;
        memset(map_out, 0, sizeof(float) * w * h);
        const float dw1 = 0.75f, dw2 = dw1 * dw1;  // setting_gradDownweightPerLevel
        int n3 = 0, n2 = 0, n4 = 0;
        for (int y4 = 0; y4 < h; y4 += 4 * pot)
            for (int x4 = 0; x4 < w; x4 += 4 * pot) {
                const int my3 = std::min(4 * pot, h - y4), mx3 = std::min(4 * pot, w - x4);
                int bestIdx4 = -1;
                float bestVal4 = 0;
                for (int y3 = 0; y3 < my3; y3 += 2 * pot)
                    for (int x3 = 0; x3 < mx3; x3 += 2 * pot) {
                        const int x34 = x3 + x4, y34 = y3 + y4;
                        const int my2 = std::min(2 * pot, h - y34), mx2 = std::min(2 * pot, w - x34);
                        int bestIdx3 = -1;
                        float bestVal3 = 0;
                        for (int y2 = 0; y2 < my2; y2 += pot)
                            for (int x2 = 0; x2 < mx2; x2 += pot) {
                                const int x234 = x2 + x34, y234 = y2 + y34;
                                const int my1 = std::min(pot, h - y234), mx1 = std::min(pot, w - x234);
                                int bestIdx2 = -1;
                                float bestVal2 = 0;
                                for (int y1 = 0; y1 < my1; ++y1)
                                    for (int x1 = 0; x1 < mx1; ++x1) {
                                        const int xf = x1 + x234, yf = y1 + y234;
                                        const int idx = xf + w * yf;
                                        if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
                                        const float pixelTH0 = thsSmoothed[(xf >> 5) + (yf >> 5) * w32];
                                        const float pixelTH1 = pixelTH0 * dw1;
                                        const float pixelTH2 = pixelTH1 * dw2;
                                        const float ag0 = mapmax0[idx];
                                        if (ag0 > pixelTH0 * thFactor) {
                                            if (ag0 > bestVal2) { bestVal2 = ag0; bestIdx2 = idx; bestIdx3 = -2; bestIdx4 = -2; }
                                        }
                                        if (bestIdx3 == -2) continue;
                                        const float ag1 = mapmax1[(int)(xf * 0.5f + 0.25f) + (int)(yf * 0.5f + 0.25f) * w1];
                                        if (ag1 > pixelTH1 * thFactor) {
                                            if (ag1 > bestVal3) { bestVal3 = ag1; bestIdx3 = idx; bestIdx4 = -2; }
                                        }
                                        if (bestIdx4 == -2) continue;
                                        const float ag2 = mapmax2[(int)(xf * 0.25f + 0.125) + (int)(yf * 0.25f + 0.125) * w2];
                                        if (ag2 > pixelTH2 * thFactor) {
                                            if (ag2 > bestVal4) { bestVal4 = ag2; bestIdx4 = idx; }
                                        }
                                    }
                                if (bestIdx2 > 0) { map_out[bestIdx2] = 1; bestVal3 = 1e10f; n2++; }
                            }
                        if (bestIdx3 > 0) { map_out[bestIdx3] = 2; bestVal4 = 1e10f; n3++; }
                    }
                if (bestIdx4 > 0) { map_out[bestIdx4] = 4; n4++; }
            }
        n[0] = n2; n[1] = n3; n[2] = n4;
    }

    // :137-282
    int makeMaps(const Pyr& p, float* map_out, float density, int recursionsLeft, float thFactor, int* pots, int& npots) {
        float numHave = 0, numWant = density, quotia;
        int idealPotential = currentPotential;
        int n[3];
        pots[npots++] = currentPotential;
        select(p, map_out, currentPotential, thFactor, n);
        numHave = (float)(n[0] + n[1] + n[2]);
        quotia = numWant / numHave;
        const float K = numHave * (currentPotential + 1) * (currentPotential + 1);
        idealPotential = (int)(sqrtf(K / numWant) - 1);
        if (idealPotential < 1) idealPotential = 1;
        if (recursionsLeft > 0 && quotia > 1.25 && currentPotential > 1) {
            if (idealPotential >= currentPotential) idealPotential = currentPotential - 1;
            currentPotential = idealPotential;
            return makeMaps(p, map_out, density, recursionsLeft - 1, thFactor, pots, npots);
        } else if (recursionsLeft > 0 && quotia < 0.25) {
            if (idealPotential <= currentPotential) idealPotential = currentPotential + 1;
            currentPotential = idealPotential;
            return makeMaps(p, map_out, density, recursionsLeft - 1, thFactor, pots, npots);
        }
        int numHaveSub = (int)numHave;
        if (quotia < 0.95) {
            const int wh = w * h;
            int rn = 0;
            const unsigned char charTH = (unsigned char)(255 * quotia);
            for (int i = 0; i < wh; ++i)
                if (map_out[i] != 0) {
                    if (randomPattern[rn] > charTH) { map_out[i] = 0; numHaveSub--; }
                    rn++;
                }
        }
        currentPotential = idealPotential;
        return numHaveSub;
    }
};

struct Cam { float scaling_factor, fx, fy, cx, cy; };
Cam camera(int dataset_seq) {  // src/pcd_generator.cpp:241-302
    switch (dataset_seq) {
        case 1: return {5000.0f, 517.3f, 516.5f, 318.6f, 255.3f};
        case 2: return {5000.0f, 520.9f, 521.0f, 325.1f, 249.7f};
        case 3: return {5000.0f, 535.4f, 539.2f, 320.1f, 247.6f};
        case 4: return {2000.0f, 718.856f, 718.856f, 607.1928f, 185.2157f};
        case 5: return {2000.0f, 707.0912f, 707.0912f, 601.8873f, 183.1104f};
        default: return {1000.0f, 616.368f, 616.745f, 319.935f, 243.639f};
    }
}

}  // namespace

extern "C" {

// cv::cvtColor(image, intensity, COLOR_RGB2GRAY) on an 8UC3 image, channel 0 weighted as "R" (src/pcd_generator.cpp:390).
void pcd_oracle_rgb2gray(const uint8_t* img3, int n, uint8_t* gray) {
    for (int i = 0; i < n; ++i) {
        const int c0 = img3[3 * i], c1 = img3[3 * i + 1], c2 = img3[3 * i + 2];
        gray[i] = (uint8_t)((c0 * 9798 + c1 * 19235 + c2 * 3735 + (1 << 14)) >> 15);
    }
}

// cv::cvtColor(image, image_hsv, COLOR_RGB2HSV), 8-bit, H in [0, 180) (src/pcd_generator.cpp:391).
void pcd_oracle_rgb2hsv(const uint8_t* img3, int n, uint8_t* hsv3) {
    static int sdiv[256], hdiv[256];
    static bool init = false;
    if (!init) {
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; ++i) {
            sdiv[i] = (int)lrint((255 << 12) / (1. * i));
            hdiv[i] = (int)lrint((180 << 12) / (6. * i));
        }
        init = true;
    }
    for (int i = 0; i < n; ++i) {
        const int r = img3[3 * i], g = img3[3 * i + 1], b = img3[3 * i + 2];
        int v = b, vmin = b;
        if (g > v) v = g;
        if (r > v) v = r;
        if (g < vmin) vmin = g;
        if (r < vmin) vmin = r;
        const int diff = v - vmin;
        const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
        const int s = (diff * sdiv[v] + (1 << 11)) >> 12;
        int hh = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
        hh = (hh * hdiv[diff] + (1 << 11)) >> 12;
        hh += hh < 0 ? 180 : 0;
        hsv3[3 * i] = (uint8_t)hh;
        hsv3[3 * i + 1] = (uint8_t)s;
        hsv3[3 * i + 2] = (uint8_t)v;
    }
}

// cv::blur(src, dst, Size(3,3)) on 8-bit (BORDER_REFLECT_101, rounded mean; sum/9 never ends in .5) followed by
// cv::Canny(dst, dst, 0, 25, 3) (L2gradient = false).  edge: w*h bytes, 0 or 255.
void pcd_oracle_blur_canny(const uint8_t* gray, int w, int h, uint8_t* blurred, uint8_t* edge) {
    auto refl = [](int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); };
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int sum = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) sum += gray[refl(y + dy, h) * w + refl(x + dx, w)];
            blurred[y * w + x] = (uint8_t)((sum + 4) / 9);
        }
    auto clampi = [](int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); };
    std::vector<int> gx((size_t)w * h), gy((size_t)w * h), mag((size_t)(w + 2) * (h + 2), 0);
    auto B = [&](int x, int y) { return (int)blurred[clampi(y, h) * w + clampi(x, w)]; };
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int dx = (B(x + 1, y - 1) + 2 * B(x + 1, y) + B(x + 1, y + 1)) - (B(x - 1, y - 1) + 2 * B(x - 1, y) + B(x - 1, y + 1));
            const int dy = (B(x - 1, y + 1) + 2 * B(x, y + 1) + B(x + 1, y + 1)) - (B(x - 1, y - 1) + 2 * B(x, y - 1) + B(x + 1, y - 1));
            gx[y * w + x] = dx;
            gy[y * w + x] = dy;
            mag[(y + 1) * (w + 2) + x + 1] = std::abs(dx) + std::abs(dy);
        }
    const int low = 0, high = 25;
    const long long TG22 = (long long)(0.4142135623730950488016887242097 * (1 << 15) + 0.5);
    std::vector<uint8_t> cls((size_t)w * h, 0);  // 0 no edge, 1 candidate, 2 edge
    std::vector<int> stack;
    const int ms = w + 2;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int* m = &mag[(y + 1) * ms + x + 1];
            const int v = *m;
            if (!(v > low)) continue;
            const int xs = gx[y * w + x], ys = gy[y * w + x];
            const long long ax = std::abs(xs), ay = (long long)std::abs(ys) << 15;
            const long long tg22x = ax * TG22, tg67x = tg22x + (ax << 16);
            bool ismax;
            if (ay < tg22x) ismax = v > m[-1] && v >= m[1];
            else if (ay > tg67x) ismax = v > m[-ms] && v >= m[ms];
            else {
                const int sgn = (xs ^ ys) < 0 ? -1 : 1;
                ismax = v > m[-ms - sgn] && v > m[ms + sgn];
            }
            if (!ismax) continue;
            if (v > high) { cls[y * w + x] = 2; stack.push_back(y * w + x); }
            else cls[y * w + x] = 1;
        }
    while (!stack.empty()) {  // 8-connected hysteresis
        const int i = stack.back();
        stack.pop_back();
        const int x = i % w, y = i / w;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx, yy = y + dy;
                if (xx < 0 || yy < 0 || xx >= w || yy >= h) continue;
                if (cls[yy * w + xx] == 1) { cls[yy * w + xx] = 2; stack.push_back(yy * w + xx); }
            }
    }
    for (int i = 0; i < w * h; ++i) edge[i] = cls[i] == 2 ? 255 : 0;
}

// the Canny top-up of select_point (src/pcd_generator.cpp:135-163): in every 8 x 8 block the first edge pixel (rows
// outer, columns inner) that is not selected yet becomes a selected pixel.  Returns 1 if it ran.
int pcd_oracle_canny_topup(const uint8_t* gray, int w, int h, int num_want, int num_selected, float* map) {
    if (!(num_selected < num_want / 3)) return 0;
    std::vector<uint8_t> blurred((size_t)w * h), edge((size_t)w * h);
    pcd_oracle_blur_canny(gray, w, h, blurred.data(), edge.data());
    const int block_size = 8;
    for (int y = 0; y < h; y += block_size)
        for (int x = 0; x < w; x += block_size) {
            bool point_got = false;
            for (int j = 0; j < block_size; ++j) {
                for (int i = 0; i < block_size; ++i)
                    if (edge[(y + j) * w + x + i] != 0 && map[(y + j) * w + x + i] == 0) {
                        map[(y + j) * w + x + i] = 1;
                        point_got = true;
                        break;
                    }
                if (point_got) break;
            }
        }
    return 1;
}

// select_point up to (not including) the Canny top-up: map_out (w*h floats, 0 / 1 / 2 / 4), returns num_selected.
// pots[0..*npots) receives the potentials the selector tried (1 or 2 entries).
int pcd_oracle_select(const uint8_t* gray, int w, int h, int num_want, float* map_out, float* dx0, float* dy0, int* pots,
                      int* npots) {
    if (w % 32 || h % 32 || w < 64 || h < 64) return -1;
    Pyr p;
    make_pyramid(gray, w, h, p);
    Selector sel(w, h);
    sel.makeHists(p);
    *npots = 0;
    const int n = sel.makeMaps(p, map_out, (float)num_want, 1, 1.0f, pots, *npots);
    memcpy(dx0, p.dx[0].data(), sizeof(float) * w * h);
    memcpy(dy0, p.dy[0].data(), sizeof(float) * w * h);
    return n;
}

// get_points_from_pixels + get_features (src/pcd_generator.cpp:233-382): raster order over the selected pixels with a
// non-zero depth.  xyz: n x 3, feat: n x 5 row-major.  Returns the number of points.
int pcd_oracle_points(const float* map, const uint16_t* depth, const uint8_t* img3, const uint8_t* hsv3, const float* dx0,
                      const float* dy0, int w, int h, int dataset_seq, int feature_type, float* xyz, float* feat) {
    const Cam cam = camera(dataset_seq);
    int idx = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint16_t dep = depth[y * w + x];
            if (map[y * w + x] != 0 && dep != 0) {
                const float z = dep / cam.scaling_factor;
                xyz[3 * idx + 2] = z;
                xyz[3 * idx + 0] = (x - cam.cx) * z / cam.fx;
                xyz[3 * idx + 1] = (y - cam.cy) * z / cam.fy;
                const int i = y * w + x;
                if (feature_type == 0) {  // :336-358
                    feat[5 * idx + 0] = (float)(hsv3[3 * i] / 180.0);
                    feat[5 * idx + 1] = (float)(hsv3[3 * i + 1] / 255.0);
                    feat[5 * idx + 2] = (float)(hsv3[3 * i + 2] / 255.0);
                    feat[5 * idx + 3] = (float)(dx0[i] / 255.0 * 2);
                    feat[5 * idx + 4] = (float)(dy0[i] / 255.0 * 2);
                } else {  // :359-381
                    feat[5 * idx + 0] = img3[3 * i];
                    feat[5 * idx + 1] = img3[3 * i + 1];
                    feat[5 * idx + 2] = img3[3 * i + 2];
                    feat[5 * idx + 3] = dx0[i];
                    feat[5 * idx + 4] = dy0[i];
                }
                ++idx;
            }
        }
    return idx;
}

}  // extern "C"

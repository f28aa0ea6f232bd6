// pcd_kernels.cuh -- sm_100a kernels for the image front end that feeds the registration loop
// (SURVEY.md 8f row 2): 8-bit colour image + 16-bit depth image -> selected pixels -> packed point cloud, all on
// the device, stream-ordered, without a host round trip before the final point count.
//
// Replaces, in the reference (paths relative to cpp/rkhs_registration/):
//   pcd_generator::load_image              src/pcd_generator.cpp:384-396   (cv::cvtColor RGB2GRAY / RGB2HSV, 8-bit)
//   pcd_generator::make_pyramid            src/pcd_generator.cpp:33-120
//   dso::PixelSelector::makeHists          thirdparty/PixelSelector2.cpp:71-136
//   dso::PixelSelector::select / makeMaps  thirdparty/PixelSelector2.cpp:137-282, 286-435
//   pcd_generator::get_points_from_pixels  src/pcd_generator.cpp:233-327
//   pcd_generator::get_features            src/pcd_generator.cpp:329-382
// Every integer / byte step is bit-exact; the float steps use explicitly rounded operations (no FMA contraction) in
// the reference's order, so the generated cloud equals the CPU restatement used by the tests bit for bit.
//
// What makes it parallel: with setting_selectDirectionDistribution == false (thirdparty/PixelSelector2.h:31) the
// selector's "random direction" never enters a comparison, so a 4pot x 4pot block's picks depend on that block
// only -- one thread walks one block with the reference's own loop nest; the two raster-order passes of the
// reference (the random sub-sampling, PixelSelector2.cpp:226-243, indexed by the RANK of a selected pixel, and the
// point / feature emission) are exclusive scans over the pixel flags.
//
// The Canny top-up for low-texture frames (src/pcd_generator.cpp:135-163: cv::blur 3x3 + cv::Canny(0, 25, 3), then
// one extra pixel per 8 x 8 block) is decided and executed on the device as well: its five kernels are always
// enqueued and return at once unless the selector kept fewer than num_want/3 pixels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cvo_b200 {

enum { PCD_STATUS_OK = 0, PCD_STATUS_NEEDS_CANNY = 1, PCD_STATUS_TOO_MANY_POINTS = 2 };

struct CamInfo {
    float scaling_factor, fx, fy, cx, cy;
};

// device-resident control block of one image push (decisions of makeMaps are taken on the device)
struct SelCtl {
    int pot[2];       // potential of the first / second select()
    int run[2];       // whether that select() runs
    int n[2][3];      // picks per level of each select()
    float quotia;     // numWant / numHave of the select() that counts
    int num_have;     // picks of that select()
    int num_selected; // after the random sub-sampling (makeMaps' return value)
    int num_points;   // selected pixels with a depth reading: the cloud size
    int status;
    int num_want;
    int canny_used;   // the low-texture top-up ran for this frame
    int num_dropped;  // pixels the random sub-sampling removed
};

struct PcdBuffers {
    int w, h;
    const uint8_t* img3;     // h x w x 3
    const uint16_t* depth;   // h x w
    float* I[3];             // dI[.][0] per pyramid level
    float* dx[3];
    float* dy[3];
    float* g2[3];            // abs_squared_grad
    float* ths;              // (w/32) * (h/32)
    float* thsSmoothed;
    uint8_t* map;            // 0 / 1 / 2 / 4 per pixel
    const uint8_t* randomPattern;
    SelCtl* ctl;
};

__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];

// cv::cvtColor(..., COLOR_RGB2GRAY), 8-bit: 15-bit fixed-point luma, channel 0 weighted as "R"
__device__ __forceinline__ int rgb2gray_u8(int c0, int c1, int c2) { return (c0 * 9798 + c1 * 19235 + c2 * 3735 + (1 << 14)) >> 15; }

// cv::cvtColor(..., COLOR_RGB2HSV), 8-bit, H in [0, 180)
__device__ __forceinline__ void rgb2hsv_u8(int r, int g, int b, int& hh, int& s, int& v) {
    v = max(b, max(g, r));
    const int vmin = min(b, min(g, r));
    const int diff = v - vmin;
    const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
    s = (diff * c_sdiv[v] + (1 << 11)) >> 12;
    hh = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    hh = (hh * c_hdiv[diff] + (1 << 11)) >> 12;
    hh += hh < 0 ? 180 : 0;
}

// level 0 intensity (src/pcd_generator.cpp:52-60) + reset of the control block
__global__ void pcd_gray_kernel(PcdBuffers b, int num_want) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        SelCtl& c = *b.ctl;
        c.pot[0] = 3; c.pot[1] = 0;   // currentPotential = 3 of a fresh PixelSelector (PixelSelector2.cpp:40)
        c.run[0] = 1; c.run[1] = 0;
        for (int s = 0; s < 2; ++s) c.n[s][0] = c.n[s][1] = c.n[s][2] = 0;
        c.quotia = 0.f; c.num_have = 0; c.num_selected = 0; c.num_points = 0; c.status = PCD_STATUS_OK;
        c.num_want = num_want;
        c.canny_used = 0;
        c.num_dropped = 0;
    }
    if (i >= b.w * b.h) return;
    b.I[0][i] = (float)rgb2gray_u8(b.img3[3 * i], b.img3[3 * i + 1], b.img3[3 * i + 2]);
}

// 2x2 box downsampling (src/pcd_generator.cpp:78-92), summed left to right
__global__ void pcd_down_kernel(const float* prev, float* cur, int wl, int hl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wl * hl) return;
    const int x = i % wl, y = i / wl, pw = wl * 2;
    const float* q = prev + 2 * x + 2 * y * pw;
    cur[i] = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(q[0], q[1]), q[pw]), q[pw + 1]));
}

// central differences and squared gradient magnitude (src/pcd_generator.cpp:94-112); the first and last rows, which
// the reference leaves uninitialised, are defined as 0
__global__ void pcd_grad_kernel(const float* I, float* dx, float* dy, float* g2, int wl, int hl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wl * hl) return;
    float gx = 0.f, gy = 0.f, g = 0.f;
    if (i >= wl && i < wl * (hl - 1)) {
        gx = __fmul_rn(0.5f, __fsub_rn(I[i + 1], I[i - 1]));
        gy = __fmul_rn(0.5f, __fsub_rn(I[i + wl], I[i - wl]));
        if (!isfinite(gx)) gx = 0.f;
        if (!isfinite(gy)) gy = 0.f;
        g = __fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy));
    }
    dx[i] = gx;
    dy[i] = gy;
    g2[i] = g;
}

// one CTA per 32 x 32 block: gradient-magnitude histogram -> median-based threshold (PixelSelector2.cpp:84-108)
__global__ void __launch_bounds__(1024) pcd_hist_kernel(PcdBuffers b) {
    __shared__ int hist[52];
    const int w = b.w, h = b.h, w32 = w / 32;
    const int bx = blockIdx.x % w32, by = blockIdx.x / w32;
    if (threadIdx.x < 52) hist[threadIdx.x] = 0;
    __syncthreads();
    const int i = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int it = i + 32 * bx, jt = j + 32 * by;
    if (!(it > w - 2 || jt > h - 2 || it < 1 || jt < 1)) {
        int g = (int)sqrtf(b.g2[0][it + jt * w]);
        if (g > 48) g = 48;
        atomicAdd(&hist[g + 1], 1);
        atomicAdd(&hist[0], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // computeHistQuantil(hist, 0.5) + setting_minGradHistAdd (:58-67,105)
        int th = (int)(hist[0] * 0.5f + 0.5f);
        int q = 90;
        for (int k = 0; k < 90; ++k) {
            th -= (k + 1 < 52) ? hist[k + 1] : 0;
            if (th < 0) { q = k; break; }
        }
        b.ths[bx + by * w32] = (float)(q + 7);
    }
}

// 3 x 3 smoothing of the block thresholds, squared (PixelSelector2.cpp:110-133)
__global__ void pcd_smooth_kernel(PcdBuffers b) {
    const int w32 = b.w / 32, h32 = b.h / 32;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= w32 * h32) return;
    const int x = t % w32, y = t / w32;
    const float* ths = b.ths;
    float sum = 0.f, num = 0.f;
    if (x > 0) {
        if (y > 0) { num += 1.f; sum = __fadd_rn(sum, ths[x - 1 + (y - 1) * w32]); }
        if (y < h32 - 1) { num += 1.f; sum = __fadd_rn(sum, ths[x - 1 + (y + 1) * w32]); }
        num += 1.f; sum = __fadd_rn(sum, ths[x - 1 + y * w32]);
    }
    if (x < w32 - 1) {
        if (y > 0) { num += 1.f; sum = __fadd_rn(sum, ths[x + 1 + (y - 1) * w32]); }
        if (y < h32 - 1) { num += 1.f; sum = __fadd_rn(sum, ths[x + 1 + (y + 1) * w32]); }
        num += 1.f; sum = __fadd_rn(sum, ths[x + 1 + y * w32]);
    }
    if (y > 0) { num += 1.f; sum = __fadd_rn(sum, ths[x + (y - 1) * w32]); }
    if (y < h32 - 1) { num += 1.f; sum = __fadd_rn(sum, ths[x + (y + 1) * w32]); }
    num += 1.f; sum = __fadd_rn(sum, ths[x + y * w32]);
    const float m = __fdiv_rn(sum, num);
    b.thsSmoothed[t] = __fmul_rn(m, m);
}

// clears the selection map if stage `s` runs
__global__ void pcd_clear_map_kernel(PcdBuffers b, int s) {
    if (!b.ctl->run[s]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.w * b.h) b.map[i] = 0;
}

// select() (PixelSelector2.cpp:286-435).  The reference walks every 4pot x 4pot block with one loop nest whose
// early-outs (bestIdx3 / bestIdx4 == -2) make the result equivalent to:
//   level 1  every pot x pot cell picks its pixel of largest ag0 among those above the block threshold;
//   level 2  a 2pot x 2pot group WITHOUT any level-1 pass picks its pixel of largest ag1 among those above th1;
//   level 3  a block without any level-1 or level-2 pass picks its pixel of largest ag2 among those above th2;
// ties go to the pixel that comes first in the loop order (strict >).  One thread evaluates one cell; the 4 cells
// of a group and the 16 cells of a block sit in adjacent lanes (in loop order) and are combined with shuffles.
struct SelCand {
    float val;  // < 0: none
    int idx;
    int ord;    // position in the loop order, for ties
};
__device__ __forceinline__ SelCand sel_better(const SelCand& a, const SelCand& b) {
    if (b.val > a.val || (b.val == a.val && b.ord < a.ord)) return b;
    return a;
}
__device__ __forceinline__ SelCand sel_shfl_xor(const SelCand& c, int m) {
    SelCand r;
    r.val = __shfl_xor_sync(0xffffffffu, c.val, m);
    r.idx = __shfl_xor_sync(0xffffffffu, c.idx, m);
    r.ord = __shfl_xor_sync(0xffffffffu, c.ord, m);
    return r;
}

__global__ void __launch_bounds__(256) pcd_select_kernel(PcdBuffers b, int s) {
    SelCtl& c = *b.ctl;
    if (!c.run[s]) return;
    const int pot = c.pot[s];
    const int w = b.w, h = b.h, w1 = w / 2, w2 = w / 4, w32 = w / 32;
    const int nbx = (w + 4 * pot - 1) / (4 * pot), nby = (h + 4 * pot - 1) / (4 * pot);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int blk = t >> 4, sub = t & 15;
    const bool in_grid = blk < nbx * nby;  // (whole 16-lane groups are in or out together)
    const int x4 = (blk % nbx) * 4 * pot, y4 = (blk / nbx) * 4 * pot;
    const int grp = sub >> 2, cell = sub & 3;
    const int x234 = x4 + (grp & 1) * 2 * pot + (cell & 1) * pot, y234 = y4 + (grp >> 1) * 2 * pot + (cell >> 1) * pot;
    const float* mapmax0 = b.g2[0];
    const float* mapmax1 = b.g2[1];
    const float* mapmax2 = b.g2[2];
    const float dw1 = 0.75f, dw2 = __fmul_rn(dw1, dw1);  // setting_gradDownweightPerLevel
    const float thFactor = 1.f;
    SelCand c1 = {-1.f, -1, sub}, c2 = {-1.f, -1, sub}, c3 = {-1.f, -1, sub};
    if (in_grid && x234 < w && y234 < h) {
        const int my1 = min(pot, h - y234), mx1 = min(pot, w - x234);
        for (int y1 = 0; y1 < my1; ++y1)
            for (int x1 = 0; x1 < mx1; ++x1) {
                const int xf = x1 + x234, yf = y1 + y234;
                const int idx = xf + w * yf;
                if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
                const float pixelTH0 = b.thsSmoothed[(xf >> 5) + (yf >> 5) * w32];
                const float pixelTH1 = __fmul_rn(pixelTH0, dw1);
                const float pixelTH2 = __fmul_rn(pixelTH1, dw2);
                const float ag0 = mapmax0[idx];
                if (ag0 > __fmul_rn(pixelTH0, thFactor) && ag0 > c1.val && ag0 > 0.f) { c1.val = ag0; c1.idx = idx; }
                const float ag1 = mapmax1[(int)__fadd_rn(__fmul_rn((float)xf, 0.5f), 0.25f) +
                                          (int)__fadd_rn(__fmul_rn((float)yf, 0.5f), 0.25f) * w1];
                if (ag1 > __fmul_rn(pixelTH1, thFactor) && ag1 > c2.val && ag1 > 0.f) { c2.val = ag1; c2.idx = idx; }
                // (int)(xf*0.25f+0.125): the literal 0.125 is a double there; exact either way
                const float ag2 = mapmax2[(int)((double)__fmul_rn((float)xf, 0.25f) + 0.125) +
                                          (int)((double)__fmul_rn((float)yf, 0.25f) + 0.125) * w2];
                if (ag2 > __fmul_rn(pixelTH2, thFactor) && ag2 > c3.val && ag2 > 0.f) { c3.val = ag2; c3.idx = idx; }
            }
    }
    // level 1: this cell's pick (bestIdx2 > 0)
    const bool p1 = c1.idx > 0;
    if (p1) b.map[c1.idx] = 1;
    // level 2: the group's pick, if no cell of the group has a level-1 pass
    bool g1 = c1.idx >= 0;  // "a pixel passed level 1" (idx 0 cannot pass: it lies in the excluded border)
    g1 |= __shfl_xor_sync(0xffffffffu, g1, 1);
    g1 |= __shfl_xor_sync(0xffffffffu, g1, 2);
    SelCand gc = c2;
    gc = sel_better(gc, sel_shfl_xor(gc, 1));
    gc = sel_better(gc, sel_shfl_xor(gc, 2));
    const bool p2 = cell == 0 && !g1 && gc.idx > 0;
    if (p2) b.map[gc.idx] = 2;
    // level 3: the block's pick, if no pixel of the block passed level 1 or level 2
    bool b12 = g1 || (c2.idx >= 0);
    b12 |= __shfl_xor_sync(0xffffffffu, b12, 1);
    b12 |= __shfl_xor_sync(0xffffffffu, b12, 2);
    b12 |= __shfl_xor_sync(0xffffffffu, b12, 4);
    b12 |= __shfl_xor_sync(0xffffffffu, b12, 8);
    SelCand bc = c3;
#pragma unroll
    for (int m = 1; m < 16; m <<= 1) bc = sel_better(bc, sel_shfl_xor(bc, m));
    const bool p3 = sub == 0 && !b12 && bc.idx > 0;
    if (p3) b.map[bc.idx] = 4;
    const int n2 = __popc(__ballot_sync(0xffffffffu, p1)), n3 = __popc(__ballot_sync(0xffffffffu, p2)),
              n4 = __popc(__ballot_sync(0xffffffffu, p3));
    if ((threadIdx.x & 31) == 0) {
        if (n2) atomicAdd(&c.n[s][0], n2);
        if (n3) atomicAdd(&c.n[s][1], n3);
        if (n4) atomicAdd(&c.n[s][2], n4);
    }
}

// makeMaps' decisions after a select() (PixelSelector2.cpp:181-224), one thread.  Stage 0 may schedule ONE
// re-selection (recursionsLeft = 1); stage 1 only records its result.
__global__ void pcd_decide_kernel(PcdBuffers b, int s) {
    SelCtl& c = *b.ctl;
    if (!c.run[s]) return;
    const int cur = c.pot[s];
    const float numHave = (float)(c.n[s][0] + c.n[s][1] + c.n[s][2]);
    const float numWant = (float)c.num_want;
    const float quotia = __fdiv_rn(numWant, numHave);
    const float K = __fmul_rn(__fmul_rn(numHave, (float)(cur + 1)), (float)(cur + 1));
    int ideal = (int)__fsub_rn(sqrtf(__fdiv_rn(K, numWant)), 1.f);
    if (ideal < 1) ideal = 1;
    c.quotia = quotia;
    c.num_have = (int)numHave;
    if (s == 0) {
        if ((double)quotia > 1.25 && cur > 1) {
            if (ideal >= cur) ideal = cur - 1;
            c.pot[1] = ideal;
            c.run[1] = 1;
        } else if ((double)quotia < 0.25) {
            if (ideal <= cur) ideal = cur + 1;
            c.pot[1] = ideal;
            c.run[1] = 1;
        }
    }
}

// Raster-order rank of the flagged pixels, spread over ceil(npix / 1024) CTAs: a counting kernel leaves one count
// per 1024-pixel block; the consuming kernel's CTA sums the counts of the blocks before it (<= a few hundred ints)
// and ranks its own pixels with warp ballots.  Pixel i is handled by thread i % 1024 of CTA i / 1024: coalesced.
__device__ __forceinline__ int block_count_flags(bool flag) {  // total over the CTA, valid in every thread
    __shared__ int wcnt[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) wcnt[warp] = __popc(bal);
    __syncthreads();
    int v = wcnt[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    return v;
}
// rank of this thread's pixel among the flagged pixels of the whole image (only meaningful where flag is set);
// *block_total / *before = flagged pixels in this CTA / in the CTAs before it
__device__ __forceinline__ int raster_rank(bool flag, const int* blockCnt, int* before, int* block_total) {
    __shared__ int wbase[32];
    __shared__ int sbefore;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) wbase[warp] = __popc(bal);
    int part = 0;  // counts of the blocks before this one
    for (int bidx = threadIdx.x; bidx < (int)blockIdx.x; bidx += blockDim.x) part += blockCnt[bidx];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (threadIdx.x == 0) sbefore = 0;
    __syncthreads();
    if (lane == 0 && part) atomicAdd(&sbefore, part);
    if (warp == 0) {
        const int v = wbase[lane];
        int iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += u;
        }
        wbase[lane] = iv - v;
        if (lane == 31) *block_total = iv;  // caller passes a __shared__ int
    }
    __syncthreads();
    *before = sbefore;
    return sbefore + wbase[warp] + __popc(bal & ((1u << lane) - 1u));
}

enum { FLAG_SELECTED = 0, FLAG_POINT = 1 };
template <int WHAT>
__global__ void __launch_bounds__(1024) pcd_count_kernel(PcdBuffers b, int* blockCnt) {
    if (WHAT == FLAG_SELECTED && !((double)b.ctl->quotia < 0.95)) return;  // no sub-sampling: nobody reads the counts
    const int i = blockIdx.x * 1024 + threadIdx.x, npix = b.w * b.h;
    bool flag = false;
    if (i < npix) flag = (WHAT == FLAG_SELECTED) ? b.map[i] != 0 : (b.map[i] != 0 && b.depth[i] != 0);
    const int c = block_count_flags(flag);
    if (threadIdx.x == 0) blockCnt[blockIdx.x] = c;
}

// the random sub-sampling of makeMaps (PixelSelector2.cpp:226-243): the rn-th selected pixel in raster order is
// dropped if randomPattern[rn] > 255 * quotia
__global__ void __launch_bounds__(1024) pcd_subsample_kernel(PcdBuffers b, const int* blockCnt) {
    SelCtl& c = *b.ctl;
    const float quotia = c.quotia;
    if (!((double)quotia < 0.95)) return;  // float against the double literal, as there
    __shared__ int stotal, sdrop;
    const int i = blockIdx.x * 1024 + threadIdx.x, npix = b.w * b.h;
    const bool flag = i < npix && b.map[i] != 0;
    int before;
    if (threadIdx.x == 0) sdrop = 0;
    const int rn = raster_rank(flag, blockCnt, &before, &stotal);
    const unsigned char charTH = (unsigned char)__fmul_rn(255.f, quotia);
    const bool drop = flag && b.randomPattern[rn] > charTH;
    if (drop) b.map[i] = 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, drop);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&sdrop, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0 && sdrop) atomicAdd(&c.num_dropped, sdrop);
}

// makeMaps' return value and the Canny condition of select_point (src/pcd_generator.cpp:135)
__global__ void pcd_after_subsample_kernel(PcdBuffers b) {
    SelCtl& c = *b.ctl;
    c.num_selected = c.num_have - c.num_dropped;
    if (c.num_selected < c.num_want / 3) c.status = PCD_STATUS_NEEDS_CANNY;
}

// ---------------------------------------------------------------------------------------------------------
// Canny top-up (src/pcd_generator.cpp:135-163).  cv::blur and cv::Canny restated from OpenCV's algorithm
// (modules/imgproc/src/canny.cpp); every kernel returns unless the selection raised PCD_STATUS_NEEDS_CANNY.
// ---------------------------------------------------------------------------------------------------------
struct CannyScratch {
    uint8_t* blurred;  // w * h
    uint16_t* mag;     // w * h: |dx| + |dy| of the 3 x 3 Sobel (<= 2040)
    uint8_t* cls;      // w * h: 0 no edge, 1 candidate (local maximum above the low threshold), 2 edge
};

// cv::blur(intensity, edge, Size(3,3)): BORDER_REFLECT_101, rounded mean (sum / 9 never ends in .5)
__global__ void pcd_blur_kernel(PcdBuffers b, CannyScratch cs) {
    if (b.ctl->status != PCD_STATUS_NEEDS_CANNY) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, w = b.w, h = b.h;
    if (i >= w * h) return;
    const int x = i % w, y = i / w;
    int sum = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        int yy = y + dy;
        yy = yy < 0 ? -yy : (yy >= h ? 2 * h - 2 - yy : yy);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            int xx = x + dx;
            xx = xx < 0 ? -xx : (xx >= w ? 2 * w - 2 - xx : xx);
            sum += (int)b.I[0][yy * w + xx];  // level 0 of the pyramid holds the 8-bit intensity exactly
        }
    }
    cs.blurred[i] = (uint8_t)((sum + 4) / 9);
}

__device__ __forceinline__ void sobel3(const uint8_t* img, int w, int h, int x, int y, int& dx, int& dy) {
    const int xm = max(x - 1, 0), xp = min(x + 1, w - 1), ym = max(y - 1, 0), yp = min(y + 1, h - 1);  // BORDER_REPLICATE
    const int a = img[ym * w + xm], bb = img[ym * w + x], c = img[ym * w + xp];
    const int d = img[y * w + xm], f = img[y * w + xp];
    const int g = img[yp * w + xm], hh = img[yp * w + x], k = img[yp * w + xp];
    dx = (c + 2 * f + k) - (a + 2 * d + g);
    dy = (g + 2 * hh + k) - (a + 2 * bb + c);
}

__global__ void pcd_sobel_mag_kernel(PcdBuffers b, CannyScratch cs) {
    if (b.ctl->status != PCD_STATUS_NEEDS_CANNY) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, w = b.w, h = b.h;
    if (i >= w * h) return;
    int dx, dy;
    sobel3(cs.blurred, w, h, i % w, i / w, dx, dy);
    cs.mag[i] = (uint16_t)(abs(dx) + abs(dy));  // L2gradient = false
}

// non-maximum suppression with the 15-bit tan(22.5 deg) sector test and its asymmetric comparisons; low = 0, high = 25
__global__ void pcd_nms_kernel(PcdBuffers b, CannyScratch cs) {
    if (b.ctl->status != PCD_STATUS_NEEDS_CANNY) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, w = b.w, h = b.h;
    if (i >= w * h) return;
    const int x = i % w, y = i / w;
    auto M = [&](int xx, int yy) { return (xx < 0 || yy < 0 || xx >= w || yy >= h) ? 0 : (int)cs.mag[yy * w + xx]; };
    const int v = cs.mag[i];
    uint8_t c = 0;
    if (v > 0) {
        int xs, ys;
        sobel3(cs.blurred, w, h, x, y, xs, ys);
        const long long TG22 = 13573;  // (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)
        const long long ax = abs(xs), ay = (long long)abs(ys) << 15;
        const long long tg22x = ax * TG22, tg67x = tg22x + (ax << 16);
        bool ismax;
        if (ay < tg22x) ismax = v > M(x - 1, y) && v >= M(x + 1, y);
        else if (ay > tg67x) ismax = v > M(x, y - 1) && v >= M(x, y + 1);
        else {
            const int sgn = (xs ^ ys) < 0 ? -1 : 1;
            ismax = v > M(x - sgn, y - 1) && v > M(x + sgn, y + 1);
        }
        if (ismax) c = v > 25 ? 2 : 1;
    }
    cs.cls[i] = c;
}

// 8-connected hysteresis: candidates touching an edge become edges, to the fixed point (which is unique, so the
// races between threads reading and promoting neighbours do not matter).  One CTA; thread t owns a contiguous
// pixel range and sweeps it forwards and backwards, so that chains along the raster order close in one pass.
__global__ void __launch_bounds__(1024) pcd_hysteresis_kernel(PcdBuffers b, CannyScratch cs) {
    if (b.ctl->status != PCD_STATUS_NEEDS_CANNY) return;
    __shared__ int changed;
    const int w = b.w, h = b.h, npix = w * h, per = (npix + 1023) / 1024;
    const int lo = min((int)threadIdx.x * per, npix), hi = min(lo + per, npix);
    volatile uint8_t* cls = cs.cls;
    auto promote = [&](int i) {
        if (cls[i] != 1) return 0;
        const int x = i % w, y = i / w;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx, yy = y + dy;
                if (xx < 0 || yy < 0 || xx >= w || yy >= h) continue;
                if (cls[yy * w + xx] == 2) {
                    cls[i] = 2;
                    return 1;
                }
            }
        return 0;
    };
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        int any = 0;
        for (int i = lo; i < hi; ++i) any |= promote(i);
        for (int i = hi - 1; i >= lo; --i) any |= promote(i);
        if (any) changed = 1;
        __threadfence_block();
        __syncthreads();
        if (!changed) break;
    }
}

// the top-up itself: in every 8 x 8 block the first edge pixel (rows outer, columns inner) that is not selected yet
// becomes a selected pixel (src/pcd_generator.cpp:144-162).  Clears the condition: the frame proceeds normally.
__global__ void pcd_topup_kernel(PcdBuffers b, CannyScratch cs) {
    if (b.ctl->status != PCD_STATUS_NEEDS_CANNY) return;
    const int w = b.w, h = b.h, nbx = w / 8, nby = h / 8;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbx * nby) return;
    const int x0 = (t % nbx) * 8, y0 = (t / nbx) * 8;
    for (int j = 0; j < 8; ++j)
        for (int i = 0; i < 8; ++i) {
            const int p = (y0 + j) * w + x0 + i;
            if (cs.cls[p] == 2 && b.map[p] == 0) {
                b.map[p] = 1;
                return;
            }
        }
}
__global__ void pcd_topup_done_kernel(PcdBuffers b) {
    if (b.ctl->status == PCD_STATUS_NEEDS_CANNY) {
        b.ctl->status = PCD_STATUS_OK;
        b.ctl->canny_used = 1;
    }
}

// get_points_from_pixels + get_features (src/pcd_generator.cpp:304-381): the selected pixels with a depth reading,
// in raster order, as n x 3 positions and n x 5 features (row-major), plus the point count for the pack job.
__global__ void __launch_bounds__(1024) pcd_points_kernel(PcdBuffers b, const int* blockCnt, CamInfo cam, int feature_type,
                                                          float* xyz, float* feat, int max_points, int* job_n) {
    SelCtl& c = *b.ctl;
    __shared__ int stotal;
    const int i = blockIdx.x * 1024 + threadIdx.x, npix = b.w * b.h, w = b.w;
    const bool flag = i < npix && b.map[i] != 0 && b.depth[i] != 0;
    int before;
    const int idx = raster_rank(flag, blockCnt, &before, &stotal);
    if (flag && idx < max_points) {
        const int x = i % w, y = i / w;
        const float z = __fdiv_rn((float)b.depth[i], cam.scaling_factor);
        xyz[3 * idx + 2] = z;
        xyz[3 * idx + 0] = __fdiv_rn(__fmul_rn(__fsub_rn((float)x, cam.cx), z), cam.fx);
        xyz[3 * idx + 1] = __fdiv_rn(__fmul_rn(__fsub_rn((float)y, cam.cy), z), cam.fy);
        const int c0 = b.img3[3 * i], c1 = b.img3[3 * i + 1], c2 = b.img3[3 * i + 2];
        if (feature_type == 0) {  // HSV / [180, 255, 255] and gradient * 2 / 255, evaluated in double (:336-358)
            int hh, s, v;
            rgb2hsv_u8(c0, c1, c2, hh, s, v);
            feat[5 * idx + 0] = (float)((double)hh / 180.0);
            feat[5 * idx + 1] = (float)((double)s / 255.0);
            feat[5 * idx + 2] = (float)((double)v / 255.0);
            feat[5 * idx + 3] = (float)((double)b.dx[0][i] / 255.0 * 2.0);
            feat[5 * idx + 4] = (float)((double)b.dy[0][i] / 255.0 * 2.0);
        } else {  // raw channels and raw gradient (:359-381)
            feat[5 * idx + 0] = (float)c0;
            feat[5 * idx + 1] = (float)c1;
            feat[5 * idx + 2] = (float)c2;
            feat[5 * idx + 3] = b.dx[0][i];
            feat[5 * idx + 4] = b.dy[0][i];
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {  // the last block sees the total
        const int total = before + stotal;
        c.num_points = total;
        if (total > max_points) c.status = PCD_STATUS_TOO_MANY_POINTS;
        // Too many points: the pack job gets n = 0, so pack_sort_kernel returns at once and the slot's planes (for a
        // bound slot: the current FIXED cloud) are left exactly as they were -- the error leaves the slot unchanged.
        *job_n = total > max_points ? 0 : total;
    }
}

}  // namespace cvo_b200

// cvo_pack.cuh -- upload-time packing: Morton sort + packed planes (pack_sort_kernel)
// (included by cvo_kernels.cuh inside namespace cvo_b200; see that file for the overall design)
#pragma once

// --------------------------------------------------------------------------------------------
// upload-time packing: Morton sort + 32-byte rows  (replaces the tail of set_pcd, src/cvo.cpp:343-356)
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

constexpr int kPackThreads = 1024;

// One CTA per cloud: bounding box -> 30-bit Morton key -> bitonic sort of (key, index) in shared
// memory -> gather into {x,y,z,f0} / {f1..f4} rows.  Ties break on the original index, so the
// packed order is a pure function of the input.
__global__ void __launch_bounds__(kPackThreads, 1) pack_sort_kernel(const PackJob* jobs, int sort_points) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    __shared__ float sbox[6][32];
    __shared__ float bb[6];
    const PackJob job = jobs[blockIdx.x];
    const int n = job.n;
    if (n <= 0) return;
    int npad = 1;
    while (npad < n) npad <<= 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int i = threadIdx.x; i < n; i += kPackThreads) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = job.xyz[3 * i + a];
            if (!isfinite(v)) continue;  // keeps the Morton grid of the finite points intact
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float l = warp_min(lo[a]), h = warp_max(hi[a]);
        if (lane == 0) {
            sbox[a][warp] = l;
            sbox[3 + a][warp] = h;
        }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float l = warp_min(sbox[a][lane]), h = warp_max(sbox[3 + a][lane]);
            if (lane == 0) {
                bb[a] = l;
                bb[3 + a] = h;
            }
        }
    }
    __syncthreads();
    float scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float ext = bb[3 + a] - bb[a];
        scale[a] = ext > 0.f ? 1023.999f / ext : 0.f;
    }
    for (int i = threadIdx.x; i < npad; i += kPackThreads) {
        unsigned long long key = ~0ull;
        if (i < n) {
            uint32_t code = 0;
            if (sort_points) {
                const uint32_t qx = (uint32_t)((job.xyz[3 * i + 0] - bb[0]) * scale[0]);
                const uint32_t qy = (uint32_t)((job.xyz[3 * i + 1] - bb[1]) * scale[1]);
                const uint32_t qz = (uint32_t)((job.xyz[3 * i + 2] - bb[2]) * scale[2]);
                code = spread10(qx) | (spread10(qy) << 1) | (spread10(qz) << 2);
            }
            key = ((unsigned long long)code << 32) | (unsigned long long)(uint32_t)i;
        }
        keys[i] = key;
    }
    __syncthreads();
    if (sort_points) {
        for (int k = 2; k <= npad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (npad >> 1); t += kPackThreads) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int p = i | j;
                    const unsigned long long a = keys[i], b = keys[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        keys[i] = b;
                        keys[p] = a;
                    }
                }
                __syncthreads();
            }
        }
    }
    for (int i = threadIdx.x; i < n; i += kPackThreads) {
        const int src = (int)(uint32_t)(keys[i] & 0xffffffffull);
        const float* p = job.xyz + 3 * src;
        const float* f = job.feat + 5 * src;
        job.out_g[i] = make_float4(p[0], p[1], p[2], __int_as_float(src));
        job.out_f[i] = make_float4(f[0], f[1], f[2], f[3]);
        job.out_f4[i] = f[4];
    }
}

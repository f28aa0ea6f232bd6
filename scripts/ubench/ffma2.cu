// Micro-benchmark (tuning aid, not product): issue rate of scalar FFMA against the packed FFMA2 / FADD2 / FMUL2 of sm_100
// (PTX fma.rn.f32x2), and a mixed FFMA + LDS stream, per SM.  Prints FMA lane-operations per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_ffma(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__global__ void k_ffma2(float* out, int iters, float a, float b) {
    unsigned long long x[8], aa, bb;
    float2 t = make_float2(a, a), u = make_float2(b, b);
    aa = *reinterpret_cast<unsigned long long*>(&t);
    bb = *reinterpret_cast<unsigned long long*>(&u);
    for (int j = 0; j < 8; ++j) {
        float2 v = make_float2(threadIdx.x + j, threadIdx.x - j);
        x[j] = *reinterpret_cast<unsigned long long*>(&v);
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u2 = 0; u2 < 8; ++u2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fma2(x[j], aa, bb);
        }
    }
    float s = 0.f;
    for (int j = 0; j < 8; ++j) {
        float2 v = *reinterpret_cast<float2*>(&x[j]);
        s += v.x + v.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FFMA interleaved with shared-memory loads (1 LDS per 4 FFMA): does the packed form free issue slots for them?
template <bool PACKED> __global__ void k_mix(float* out, int iters, float a, float b) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    unsigned long long x[4], aa, bb;
    float2 t = make_float2(a, a), u = make_float2(b, b);
    aa = *reinterpret_cast<unsigned long long*>(&t);
    bb = *reinterpret_cast<unsigned long long*>(&u);
    float y[8];
    for (int j = 0; j < 8; ++j) y[j] = threadIdx.x + j;
    for (int j = 0; j < 4; ++j) {
        float2 v = make_float2(threadIdx.x + j, threadIdx.x - j);
        x[j] = *reinterpret_cast<unsigned long long*>(&v);
    }
    float acc = 0.f;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u2 = 0; u2 < 8; ++u2) {
            if (PACKED) {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = fma2(x[j], aa, bb);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = fmaf(y[j], a, b);
            }
            acc += sm[(idx + u2 * 33) & 1023];
            acc += sm[(idx + u2 * 65 + 7) & 1023];
        }
        idx += 1;
    }
    float s = acc;
    for (int j = 0; j < 8; ++j) s += y[j];
    for (int j = 0; j < 4; ++j) {
        float2 v = *reinterpret_cast<float2*>(&x[j]);
        s += v.x + v.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 4 * 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32, blocks = sms;
        float t1 = time_ms([&] { k_ffma<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        float t2 = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        float t3 = time_ms([&] { k_mix<false><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        float t4 = time_ms([&] { k_mix<true><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
        const double cyc = (double)clk * 1e3;  // cycles per second (nominal max clock)
        const double fma1 = (double)iters * 64 * threads / (t1 * 1e-3 * cyc);
        const double fma2v = (double)iters * 64 * 2 * threads / (t2 * 1e-3 * cyc);
        const double mix1 = (double)iters * 64 * threads / (t3 * 1e-3 * cyc);
        const double mix2 = (double)iters * 64 * threads / (t4 * 1e-3 * cyc);
        printf("warps/SM %2d: FFMA %.1f fma/clk/SM | FFMA2 %.1f fma/clk/SM | 8 FMA + 2 LDS: scalar %.1f  packed %.1f fma/clk/SM\n", warps,
               fma1, fma2v, mix1, mix2);
    }
    return 0;
}

// Micro-benchmark (tuning aid): issue cost of FP32 instructions by OPERAND FORM on sm_100 -- three distinct register
// operands vs repeated operands vs immediates, scalar FFMA vs packed FFMA2 / FMUL2 / FADD2 (register-file bandwidth).
// Prints warp-instructions per clock per SM sub-partition and FP32 lane-operations per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define REP8(X) X X X X X X X X
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// 16 independent accumulators; operands rotate so that every instruction reads three DIFFERENT registers
template <int MODE> __global__ void k(float* out, int iters, float s) {
    float a[16], b[16];
    float2 p[16], q[16];
    for (int j = 0; j < 16; ++j) { a[j] = threadIdx.x + j; b[j] = 1.0f + 1e-7f * (threadIdx.x + 3 * j); p[j] = f2(a[j], a[j] + 1); q[j] = f2(b[j], b[j]); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (MODE == 0) a[j] = fmaf(a[j], b[(j + 1) & 15], b[(j + 5) & 15]);              // FFMA, 3 distinct regs
                if (MODE == 1) a[j] = fmaf(a[j], s, 0.5f);                                        // FFMA, reg * uniform + imm
                if (MODE == 2) p[j] = __ffma2_rn(p[j], q[(j + 1) & 15], q[(j + 5) & 15]);         // FFMA2, 3 distinct pairs
                if (MODE == 3) p[j] = __ffma2_rn(p[j], f2(s, s), f2(0.5f, 0.5f));                 // FFMA2, broadcast + imm
                if (MODE == 4) p[j] = __fmul2_rn(p[j], q[(j + 1) & 15]);                          // FMUL2, 2 distinct pairs
                if (MODE == 5) p[j] = __fadd2_rn(p[j], q[(j + 1) & 15]);                          // FADD2, 2 distinct pairs
                if (MODE == 6) a[j] = a[j] * b[(j + 1) & 15];                                     // FMUL, 2 distinct regs
                if (MODE == 7) p[j] = __ffma2_rn(p[j], q[(j + 1) & 15], p[j]);                    // FFMA2, a*b+a (2 distinct)
                if (MODE == 8) p[j] = __ffma2_rn(p[j], f2(b[j], b[j]), q[(j + 5) & 15]);          // FFMA2, pair * scalar-broadcast reg + pair
            }
        }
    }
    float r = 0.f;
    for (int j = 0; j < 16; ++j) r += a[j] + p[j].x + p[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char* name, float* out, int sms, int clk_khz, int lanes_per_inst) {
    const int iters = 4000, threads = 512;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, threads>>>(out, iters, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<sms, threads>>>(out, iters, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    const double winst = (double)iters * 64 * (threads / 32);  // warp instructions per SM
    printf("%-46s %.2f warp-inst/clk/SMSP   %.1f fp32 lane-ops/clk/SM\n", name, winst / cycles / 4, winst * 32 * lanes_per_inst / cycles);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 512);
    run<0>("FFMA   a = a*b[j+1] + b[j+5] (3 regs)", out, p.multiProcessorCount, clk, 1);
    run<1>("FFMA   a = a*uniform + imm", out, p.multiProcessorCount, clk, 1);
    run<6>("FMUL   a = a*b[j+1] (2 regs)", out, p.multiProcessorCount, clk, 1);
    run<2>("FFMA2  p = p*q[j+1] + q[j+5] (3 pairs)", out, p.multiProcessorCount, clk, 2);
    run<7>("FFMA2  p = p*q[j+1] + p (2 pairs)", out, p.multiProcessorCount, clk, 2);
    run<8>("FFMA2  p = p*bcast(b[j]) + q[j+5]", out, p.multiProcessorCount, clk, 2);
    run<3>("FFMA2  p = p*bcast(uniform) + imm", out, p.multiProcessorCount, clk, 2);
    run<4>("FMUL2  p = p*q[j+1] (2 pairs)", out, p.multiProcessorCount, clk, 2);
    run<5>("FADD2  p = p+q[j+1] (2 pairs)", out, p.multiProcessorCount, clk, 2);
    return 0;
}

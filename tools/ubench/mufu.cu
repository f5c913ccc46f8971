// Micro-benchmark: per-SM throughput of MUFU-class ops on sm_100a (results/clk/SM), to size the softmax / GELU epilogues.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#define ITER 4096
template <int OP> __device__ __forceinline__ float op(float x) {
    float y;
    if constexpr (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if constexpr (OP == 1) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if constexpr (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if constexpr (OP == 3) { uint32_t a = __float_as_uint(x), b; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(b) : "r"(a)); y = __uint_as_float(b); }
    else if constexpr (OP == 4) { uint32_t a = __float_as_uint(x), b; asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(b) : "r"(a)); y = __uint_as_float(b); }
    else if constexpr (OP == 5) {   // software exp2 on the FMA/ALU pipes (Cody-Waite + cubic)
        x = fmaxf(x, -126.0f);
        const float t = x + 12582912.0f;
        const float f = x - (t - 12582912.0f);
        float p = fmaf(0.055008930605317384f, f, 0.24221095955059072f);
        p = fmaf(p, f, 0.6932829271997277f);
        p = fmaf(p, f, 1.0f);
        y = __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
    } else if constexpr (OP == 6) { y = fmaf(x, 1.0001f, 0.5f); }
    else if constexpr (OP == 7) { uint32_t a = __float_as_uint(x), b; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(b) : "r"(a)); y = __uint_as_float(b); }
    return y;
}
// one MUFU.EX2 + N independent FFMA per step, 8 independent chains per thread: is the issue of a MUFU overlapped with FMA issue?
template <int NF> __global__ void kmix(float* out, float seed) {
    float v[8], w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = seed * (threadIdx.x + j) * 1e-3f - 1.0f; w[j] = v[j] * 0.5f; }
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
#pragma unroll
            for (int f = 0; f < NF; ++f) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(w[j]) : "f"(1.0001f), "f"(0.5f));
        }
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += v[j] + w[j];
    if (s == 12345.678f) out[0] = s;
}
template <int NF> void runmix(int warps_per_sm) {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int blocks = p.multiProcessorCount, threads = warps_per_sm * 32;
    kmix<NF><<<blocks, threads>>>(d, 0.5f); cudaDeviceSynchronize();
    cudaEventRecord(a); kmix<NF><<<blocks, threads>>>(d, 0.5f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double steps = (double)warps_per_sm * ITER * 8;        // warp-level (MUFU + NF FFMA) groups per SM
    printf("mix 1 MUFU + %2d FFMA, %2d warps/SM: %8.3f ms  %6.2f cycles per group per scheduler\n", NF, warps_per_sm, ms,
           ms * 1e-3 * khz * 1e3 / (steps / 4.0));
    cudaFree(d);
}
template <int OP> __global__ void k(float* out, float seed) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = seed * (threadIdx.x + j) * 1e-3f - 1.0f;
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = op<OP>(v[j]);
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += v[j];
    if (s == 12345.678f) out[0] = s;
}
template <int OP> void run(const char* name, int per_lane) {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int blocks = p.multiProcessorCount, threads = 1024;
    k<OP><<<blocks, threads>>>(d, 0.5f); cudaDeviceSynchronize();
    cudaEventRecord(a); k<OP><<<blocks, threads>>>(d, 0.5f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double ops = (double)threads * ITER * 8 * per_lane;          // per SM
    printf("%-28s %8.3f ms  %7.2f results/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3 * khz * 1e3), khz / 1000);
    cudaFree(d);
}
int main() {
    run<0>("ex2.approx.ftz.f32", 1); run<1>("tanh.approx.f32", 1); run<2>("rcp.approx.ftz.f32", 1);
    run<3>("ex2.approx.ftz.bf16x2", 2); run<7>("ex2.approx.f16x2", 2); run<4>("tanh.approx.bf16x2", 2);
    run<5>("exp2 poly (fma pipe)", 1); run<6>("ffma", 1);
    // a warp-wide MUFU occupies its pipe for 8 cycles: do N FFMAs of the same warp (or of other warps) fit underneath?
    runmix<0>(8); runmix<4>(8); runmix<8>(8); runmix<12>(8); runmix<0>(32); runmix<4>(32); runmix<8>(32); runmix<12>(32);
    return 0;
}

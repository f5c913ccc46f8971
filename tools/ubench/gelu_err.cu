// Accuracy of candidate bf16-epilogue GELUs against the exact erf GELU, on a dense grid (device vs double on host).
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__device__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ float tanha(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ float gelu_sig(float x) {      // shipped: x * sigmoid(2u), 2 MUFU
    const float xc = fminf(fmaxf(x, -6.0f), 6.0f), x2 = xc * xc;
    const float p = fmaf(x2, fmaf(x2, -0.0010142630198970437f, 0.10677572339773178f), 2.301121234893799f);
    const float e = ex2a(xc * p);
    return fmaf(-x, rcpa(1.0f + e), x);
}
__device__ float gelu_tanh(float x) {     // candidate: 0.5 x (1 + tanh(u)), 1 MUFU
    const float xc = fminf(fmaxf(x, -6.0f), 6.0f), x2 = xc * xc;
    const float p = fmaf(x2, fmaf(x2, -0.00035151677629392575f, 0.03700564581269318f), 0.7975078480466281f);
    const float t = tanha(xc * p);
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}
__global__ void k(const float* x, float* a, float* b, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = gelu_sig(x[i]); b[i] = gelu_tanh(x[i]); }
}
int main() {
    const int n = 1 << 22;
    std::vector<float> hx(n), ha(n), hb(n);
    for (int i = 0; i < n; ++i) hx[i] = -10.0f + 20.0f * i / (n - 1);
    float *dx, *da, *db;
    cudaMalloc(&dx, n * 4); cudaMalloc(&da, n * 4); cudaMalloc(&db, n * 4);
    cudaMemcpy(dx, hx.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(dx, da, db, n);
    cudaMemcpy(ha.data(), da, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), db, n * 4, cudaMemcpyDeviceToHost);
    double ea = 0, eb = 0, xa = 0, xb = 0, ra = 0, rb = 0, ebn = 0, xbn = 0;
    for (int i = 0; i < n; ++i) {
        const double x = hx[i], ref = 0.5 * x * (1.0 + erf(x / sqrt(2.0)));
        const double da_ = fabs(ha[i] - ref), db_ = fabs(hb[i] - ref);
        if (da_ > ea) { ea = da_; xa = x; }
        if (db_ > eb) { eb = db_; xb = x; }
        if (x < -1.5 && db_ > ebn) { ebn = db_; xbn = x; }
        const double ulp = fabs(ref) * pow(2.0, -8);        // one bf16 ulp of the result
        if (fabs(ref) > 1e-3) { ra = fmax(ra, da_ / ulp); rb = fmax(rb, db_ / ulp); }
    }
    printf("sigmoid form (2 MUFU): max abs err %.3e at x=%.3f, max err / bf16 ulp (|y|>1e-3) %.3f\n", ea, xa, ra);
    printf("tanh form    (1 MUFU): max abs err %.3e at x=%.3f, max err / bf16 ulp (|y|>1e-3) %.3f; negative tail (x<-1.5) max abs %.3e at x=%.3f\n", eb, xb, rb, ebn, xbn);
    return 0;
}

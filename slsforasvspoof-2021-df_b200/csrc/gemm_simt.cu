// fp32 CUDA-core GEMM used by the fp32 verification mode (<= 1e-4 parity with the CPU oracle; tensor-core
// "fp32" would be TF32 and cannot hold 1e-4 through 24 layers).
//   out[z][m][n] = act(sum_k A[z][m][k] * W[g][n][k] + bias[n]) (+ residual[z][m][n])
// A rows may overlap (lda < K) and k may be two-level strided, which is how the strided Conv1d layers and the
// grouped positional conv run as implicit GEMMs over channels-last activations without an im2col buffer.
#include "common.cuh"
#include "kernels.h"

namespace slsb {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

struct SimtDev {
    const float* A; long long lda, a_batch_stride, a_group_offset; int a_kinner; long long a_kouter;
    const float* W; long long ldw, w_group_stride;
    int groups, n_per_group;
    int M, N, K;
    float* out; long long ldc, out_batch_stride;
    const float* bias;
    const float* residual; long long ldr, res_batch_stride;
    int act, exact_gelu;
};

__global__ void __launch_bounds__(256) simt_gemm_kernel(const SimtDev p) {
    __shared__ float As[2][BK][BM + PAD];
    __shared__ float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int z = blockIdx.z;
    const int batch = z / p.groups, grp = z - batch * p.groups;
    const float* A = p.A + batch * p.a_batch_stride + grp * p.a_group_offset;
    const float* W = p.W + grp * p.w_group_stride;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int col_off = grp * p.n_per_group;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // loader mapping: 2 float4 of A and 2 of W per thread per k-tile
    int lrow[2], lk[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { int idx = tid + i * 256; lrow[i] = idx >> 2; lk[i] = (idx & 3) * 4; }

    auto a_koff = [&](int k) -> long long {
        if (p.a_kinner > 0) { int o = k / p.a_kinner; return (long long)o * p.a_kouter + (k - o * p.a_kinner); }
        return k;
    };
    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int m = m0 + lrow[i];
            ra[i] = (m < p.M) ? *reinterpret_cast<const float4*>(A + (long long)m * p.lda + a_koff(k0 + lk[i])) : make_float4(0, 0, 0, 0);
            const int n = n0 + lrow[i];
            rb[i] = (n < p.N) ? *reinterpret_cast<const float4*>(W + (long long)n * p.ldw + k0 + lk[i]) : make_float4(0, 0, 0, 0);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            As[buf][lk[i] + 0][lrow[i]] = ra[i].x; As[buf][lk[i] + 1][lrow[i]] = ra[i].y;
            As[buf][lk[i] + 2][lrow[i]] = ra[i].z; As[buf][lk[i] + 3][lrow[i]] = ra[i].w;
            Bs[buf][lk[i] + 0][lrow[i]] = rb[i].x; Bs[buf][lk[i] + 1][lrow[i]] = rb[i].y;
            Bs[buf][lk[i] + 2][lrow[i]] = rb[i].z; Bs[buf][lk[i] + 3][lrow[i]] = rb[i].w;
        }
    };

    const int nk = p.K / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

    float* out = p.out + batch * p.out_batch_stride + col_off;
    const float* res = p.residual ? p.residual + batch * p.res_batch_stride + col_off : nullptr;
    const float* bias = p.bias + col_off;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tx * 8 + j;
            if (n >= p.N) continue;
            float v = acc[i][j] + bias[n];
            if (p.act == ACT_GELU) v = p.exact_gelu ? gelu_exact(v) : gelu_fast(v);
            else if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
            if (res) v += res[(long long)m * p.ldr + n];
            out[(long long)m * p.ldc + n] = v;
        }
    }
}

}  // namespace

int simt_gemm(const SimtGemmArgs& g, cudaStream_t stream) {
    if (g.a_bf16 || g.out_bf16) { set_error("simt_gemm is the fp32 path only"); return -1; }
    if (g.K % BK != 0 || (g.a_kinner > 0 && g.a_kinner % BK != 0)) { set_error("simt_gemm: K=%d / kinner=%d must be multiples of %d", g.K, g.a_kinner, BK); return -1; }
    if (g.lda % 4 != 0 || g.ldw % 4 != 0 || g.a_kouter % 4 != 0 || g.a_group_offset % 4 != 0 || g.a_batch_stride % 4 != 0 || g.w_group_stride % 4 != 0) {
        set_error("simt_gemm: strides must be multiples of 4 floats"); return -1;
    }
    if (g.M <= 0 || g.N <= 0 || g.batches <= 0) return 0;
    SimtDev p{};
    p.A = static_cast<const float*>(g.A); p.lda = g.lda; p.a_batch_stride = g.a_batch_stride; p.a_group_offset = g.a_group_offset;
    p.a_kinner = g.a_kinner; p.a_kouter = g.a_kouter;
    p.W = g.W; p.ldw = g.ldw; p.w_group_stride = g.w_group_stride;
    p.groups = g.groups < 1 ? 1 : g.groups; p.n_per_group = g.n_per_group;
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.out = static_cast<float*>(g.out); p.ldc = g.ldc; p.out_batch_stride = g.out_batch_stride;
    p.bias = g.bias; p.residual = g.residual; p.ldr = g.ldr; p.res_batch_stride = g.res_batch_stride;
    p.act = g.act; p.exact_gelu = g.exact_gelu;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.batches * p.groups);
    if (grid.z > 65535 || grid.y > 65535) { set_error("simt_gemm: grid too large"); return -1; }
    simt_gemm_kernel<<<grid, 256, 0, stream>>>(p);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace slsb

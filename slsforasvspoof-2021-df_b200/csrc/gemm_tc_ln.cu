// Feature-extractor layer as ONE tensor-core kernel:  out = GELU(LayerNorm_512(A * W^T + bias))  (bf16 out)
// (wav2vec2.py:785-822: Conv1d -> Fp32LayerNorm over the 512 channels -> GELU, extractor_mode="layer_norm").
//
// The CTA tile is 128 rows x all 512 channels, so a whole LayerNorm row lives in one CTA's TMEM (128 lanes x 512 fp32
// columns = the full 256 KB) and the normalisation + GELU run in the epilogue straight out of TMEM: the pre-norm
// activations never touch HBM (saves one write + one read of every conv output and six stand-alone LN launches).
//   * A operand: A_PLAIN ([M, K] bf16; used for conv0 after a hi/lo bf16 split of the raw audio, see frontend.cu)
//                or A_CONV (implicit GEMM over channels-last activations through a 4-D tensor map, see gemm_tc.cu).
//   * B operand: W [512, K] bf16, loaded as two 256-row TMA boxes per k-block; two tcgen05.mma (N = 256) per UMMA_K step.
//   * 2-stage smem ring (80 KB / stage); the single 512-column accumulator means MMA and epilogue of one CTA alternate.
//   * epilogue: 8 warps, thread == row; pass 1 reads TMEM for sum / sum-of-squares (two column halves combine through
//     smem), pass 2 re-reads TMEM, normalises, applies affine + GELU and stores bf16.
#include "common.cuh"
#include "kernels.h"

namespace slsb {
namespace {

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16, NCH = 512;
constexpr int kStages = 2;
constexpr int kStageA = BLOCK_M * BLOCK_K * 2;          // 16 KB
constexpr int kStageB = NCH * BLOCK_K * 2;              // 64 KB
constexpr int kStage = kStageA + kStageB;
constexpr int kStoreOffset = kStages * kStage;          // 2 x [128 rows x 64 bf16] SWIZZLE_128B staging tiles (one per column half)
constexpr int kParamOffset = kStoreOffset + 2 * 16384;  // bias | ln_w | ln_b, 512 floats each (constant per launch)
constexpr int kPartOffset = kParamOffset + 3 * NCH * 4; // float2 part[2 buffers][2 halves][128 rows]
constexpr int kBarOffset = kPartOffset + 2 * 2 * 128 * 8;
constexpr int kSmemBytes = kBarOffset + 128;
constexpr int kThreads = 384;

struct LnDev {
    int M, K, batches, m_tiles;
    int conv_cin, conv_stride;
    bf16* out; long long out_batch_stride;
    const float* bias; const float* ln_w; const float* ln_b;
    float eps;
};

template <int A_MODE>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_out, const LnDev p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("slsb: dynamic smem base not 1024-aligned\n"); __trap(); }
    float* par = reinterpret_cast<float*>(smem + kParamOffset);
    float2* part = reinterpret_cast<float2*>(smem + kPartOffset);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.batches * p.m_tiles;
    const int num_kb = p.K / BLOCK_K;

    griddep_launch();
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); tma_prefetch_desc(&tmap_out); }
    for (int i = threadIdx.x; i < NCH; i += kThreads) { par[i] = p.bias[i]; par[NCH + i] = p.ln_w[i]; par[2 * NCH + i] = p.ln_b[i]; }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1); mbar_init(tmem_empty, 8);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    griddep_wait();                     // prologue (params are static weights) overlapped the previous kernel's tail

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % p.m_tiles, b = tile / p.m_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kStage;
                    uint8_t* sb = sa + kStageA;
                    mbar_expect_tx(&full_bar[stage], kStage);
                    if constexpr (A_MODE == A_PLAIN) {
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
                    } else {
                        const int k0 = kb * BLOCK_K;
                        const int tap = k0 / p.conv_cin, c = k0 - tap * p.conv_cin;
                        tma_load_4d(sa, &tmap_a, &full_bar[stage], c, tap % p.conv_stride, m_blk * BLOCK_M + tap / p.conv_stride, b);
                    }
                    tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, 0);
                    tma_load_2d(sb + kStageB / 2, &tmap_b, &full_bar[stage], kb * BLOCK_K, 256);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, 256);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                mbar_wait(tmem_empty, (it & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStage);
                    const uint32_t sb = sa + kStageA;
                    const uint64_t da = make_smem_desc_sw128(sa, 0, 1024);
                    const uint64_t db0 = make_smem_desc_sw128(sb, 0, 1024);
                    const uint64_t db1 = make_smem_desc_sw128(sb + kStageB / 2, 0, 1024);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                        tc_mma_f16(tmem_base, da + uint64_t(k * 2), db0 + uint64_t(k * 2), idesc, accum);
                        tc_mma_f16(tmem_base + 256, da + uint64_t(k * 2), db1 + uint64_t(k * 2), idesc, accum);
                    }
                    tc_commit(&empty_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(tmem_full);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + half * 256;
        const int col_base = half * 256;
        const float* bias_s = par + col_base;
        const float* g_s = par + NCH + col_base;
        const float* h_s = par + 2 * NCH + col_base;
        uint8_t* stage_tile = smem + kStoreOffset + half * 16384;
        uint8_t* srow = stage_tile + r * 128;
        const int bar_id = 2 + half;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_blk = tile % p.m_tiles, b = tile / p.m_tiles;
            mbar_wait(tmem_full, it & 1);
            tc_fence_after();
            // pass 1: statistics over this warp's 256 columns; TMEM loads double-buffered in registers
            float s = 0.f, ss = 0.f;
            {
                uint32_t acc[2][32];
                tmem_ld_32x32b_x32(taddr, acc[0]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c + 1 < 8) tmem_ld_32x32b_x32(taddr + (c + 1) * 32, acc[(c + 1) & 1]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bb = *reinterpret_cast<const float4*>(bias_s + c * 32 + 4 * j);
                        const float v0 = __uint_as_float(acc[c & 1][4 * j + 0]) + bb.x, v1 = __uint_as_float(acc[c & 1][4 * j + 1]) + bb.y;
                        const float v2 = __uint_as_float(acc[c & 1][4 * j + 2]) + bb.z, v3 = __uint_as_float(acc[c & 1][4 * j + 3]) + bb.w;
                        s += (v0 + v1) + (v2 + v3);
                        ss = fmaf(v0, v0, ss); ss = fmaf(v1, v1, ss); ss = fmaf(v2, v2, ss); ss = fmaf(v3, v3, ss);
                    }
                    tmem_ld_wait();
                }
            }
            float2* pbuf = part + (it & 1) * 256;
            pbuf[half * 128 + r] = make_float2(s, ss);
            asm volatile("bar.sync 1, 256;" ::: "memory");            // the 8 epilogue warps only
            const float2 other = pbuf[(half ^ 1) * 128 + r];
            const float mean = (s + other.x) * (1.0f / NCH);
            const float var = fmaxf((ss + other.y) * (1.0f / NCH) - mean * mean, 0.0f);
            const float rstd = rsqrtf(var + p.eps);
            // pass 2: normalise + affine + GELU -> bf16 staging tile (64 columns at a time) -> TMA store
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                uint32_t acc[2][32];
                tmem_ld_32x32b_x32(taddr + g * 64, acc[0]);
                tmem_ld_32x32b_x32(taddr + g * 64 + 32, acc[1]);
                if (r == 0) tma_store_wait_read<0>();
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                tmem_ld_wait();
                if (g == 3) {                                          // accumulator fully read -> MMA warp may start the next tile
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c0 = g * 64 + h * 32;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * j);
                        const float4 gg = *reinterpret_cast<const float4*>(g_s + c0 + 4 * j);
                        const float4 hh = *reinterpret_cast<const float4*>(h_s + c0 + 4 * j);
                        v[4 * j + 0] = gelu_fast(fmaf((__uint_as_float(acc[h][4 * j + 0]) + bb.x - mean) * rstd, gg.x, hh.x));
                        v[4 * j + 1] = gelu_fast(fmaf((__uint_as_float(acc[h][4 * j + 1]) + bb.y - mean) * rstd, gg.y, hh.y));
                        v[4 * j + 2] = gelu_fast(fmaf((__uint_as_float(acc[h][4 * j + 2]) + bb.z - mean) * rstd, gg.z, hh.z));
                        v[4 * j + 3] = gelu_fast(fmaf((__uint_as_float(acc[h][4 * j + 3]) + bb.w - mean) * rstd, gg.w, hh.w));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 w;
                        w.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); w.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                        w.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); w.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                        *reinterpret_cast<uint4*>(srow + (((h * 4 + j) ^ (r & 7)) << 4)) = w;
                    }
                }
                fence_proxy_async_smem();
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (r == 0) {
                    if constexpr (A_MODE == A_CONV) tma_store_3d(&tmap_out, stage_tile, col_base + g * 64, m_blk * BLOCK_M, b);
                    else tma_store_2d(&tmap_out, stage_tile, col_base + g * 64, m_blk * BLOCK_M);
                    tma_store_commit();
                }
            }
        }
        if (r == 0) tma_store_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int A_MODE>
int launch_ln(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const LnDev& dp, int num_sms, cudaStream_t stream) {
    static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
    auto kern = tc_gemm_ln_kernel<A_MODE>;
    if (first_use_on_device(&configured_on)) {
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    }
    const int tiles = dp.batches * dp.m_tiles;
    SLSB_CUDA_CHECK(launch_pdl(kern, dim3(tiles < num_sms ? tiles : num_sms), dim3(kThreads), kSmemBytes, stream, ta, tb, to, dp));
    return 0;
}

// raw audio -> hi/lo bf16 im2col rows for conv0: [x_hi(10) 0.. | x_hi(10) 0.. | x_lo(10) 0.. | 0 (16)]
// Eight threads build one 128-byte row, one 16-byte unit each (unit u holds elements [8u, 8u + 8)), so a warp writes 512
// contiguous bytes per store instruction; the 10 audio samples of a row come from L1 (8 threads share them).
__global__ void __launch_bounds__(256) conv0_im2col_kernel(const float* __restrict__ wav, bf16* __restrict__ out, int S, int L0, int k, int stride,
                                                           long long rows) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = gid >> 3;
    const int u = (int)(gid & 7);
    if (row >= rows) return;
    const int b = (int)(row / L0), f = (int)(row - (long long)b * L0);
    const float* x = wav + (long long)b * S + (long long)f * stride;
    const int seg = u >> 1;                  // 16-element segment: 0 = x_hi, 1 = x_hi, 2 = x_lo, 3 = zeros
    const int t0 = (u & 1) * 8;              // first tap of this unit inside its segment
    __align__(16) bf16 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = t0 + i;
        float val = 0.f;
        if (seg < 3 && t < k) {
            const float xv = __ldg(x + t);
            const bf16 hi = __float2bfloat16_rn(xv);
            val = seg == 2 ? xv - __bfloat162float(hi) : __bfloat162float(hi);
        }
        v[i] = __float2bfloat16_rn(val);
    }
    *reinterpret_cast<uint4*>(out + row * 64 + u * 8) = *reinterpret_cast<const uint4*>(v);
}

// conv0 weights [C, k] fp32 -> [C, 64] bf16: [w_hi | w_lo | w_hi | 0] so that A.W^T = x_hi w_hi + x_hi w_lo + x_lo w_hi
__global__ void conv0_pack_w_kernel(const float* __restrict__ w, bf16* __restrict__ out, int C, int k) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int i = 0; i < 64; ++i) out[c * 64 + i] = __float2bfloat16_rn(0.f);
    for (int t = 0; t < k; ++t) {
        const float wv = w[c * k + t];
        const bf16 hi = __float2bfloat16_rn(wv);
        const bf16 lo = __float2bfloat16_rn(wv - __bfloat162float(hi));
        out[c * 64 + t] = hi; out[c * 64 + 16 + t] = lo; out[c * 64 + 32 + t] = hi;
    }
}

}  // namespace

int tc_gemm_ln_gelu(const TcLnGemmArgs& g, int num_sms, cudaStream_t stream) {
    if (g.N != NCH) { set_error("tc_gemm_ln_gelu: N must be 512 (got %d)", g.N); return -1; }
    if (g.K % BLOCK_K != 0 || g.K <= 0) { set_error("tc_gemm_ln_gelu: K=%d must be a positive multiple of 64", g.K); return -1; }
    if (g.M <= 0 || g.batches <= 0) return 0;
    LnDev dp{};
    dp.M = g.M; dp.K = g.K; dp.batches = g.batches; dp.m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
    dp.conv_cin = g.conv_cin; dp.conv_stride = g.conv_stride;
    dp.out = static_cast<bf16*>(g.out); dp.out_batch_stride = g.out_batch_stride;
    dp.bias = g.bias; dp.ln_w = g.ln_w; dp.ln_b = g.ln_b; dp.eps = g.eps;
    CUtensorMap ta, tb, to;
    if (g.a_mode == A_PLAIN) {
        uint64_t dims[2] = {(uint64_t)NCH, (uint64_t)g.M};
        uint64_t strides[1] = {(uint64_t)NCH * 2};
        uint32_t box[2] = {64, BLOCK_M};
        if (encode_tmap_bf16(&to, g.out, 2, dims, strides, box)) return -1;
    } else {
        uint64_t dims[3] = {(uint64_t)NCH, (uint64_t)g.M, (uint64_t)g.batches};
        uint64_t strides[2] = {(uint64_t)NCH * 2, (uint64_t)g.out_batch_stride * 2};
        uint32_t box[3] = {64, BLOCK_M, 1};
        if (encode_tmap_bf16(&to, g.out, 3, dims, strides, box)) return -1;
    }
    {
        uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)NCH};
        uint64_t strides[1] = {(uint64_t)g.K * 2};
        uint32_t box[2] = {BLOCK_K, 256};
        if (encode_tmap_bf16(&tb, g.W, 2, dims, strides, box)) return -1;
    }
    if (g.a_mode == A_PLAIN) {
        uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.M};
        uint64_t strides[1] = {(uint64_t)g.lda * 2};
        uint32_t box[2] = {BLOCK_K, BLOCK_M};
        if (encode_tmap_bf16(&ta, g.A, 2, dims, strides, box)) return -1;
        return launch_ln<A_PLAIN>(ta, tb, to, dp, num_sms, stream);
    }
    const uint64_t C = g.conv_cin, s = g.conv_stride, Lin = g.conv_lin;
    uint64_t dims[4] = {C, s, (Lin + s - 1) / s, (uint64_t)g.batches};
    uint64_t strides[3] = {C * 2, s * C * 2, Lin * C * 2};
    uint32_t box[4] = {BLOCK_K, 1, BLOCK_M, 1};
    if (encode_tmap_bf16(&ta, g.A, 4, dims, strides, box)) return -1;
    return launch_ln<A_CONV>(ta, tb, to, dp, num_sms, stream);
}

int conv0_im2col(const float* wav, void* out, int B, int S, int L0, int k, int stride, cudaStream_t stream) {
    if (k > 16) { set_error("conv0_im2col: k=%d > 16", k); return -1; }
    const long long rows = (long long)B * L0;
    conv0_im2col_kernel<<<(unsigned)((rows * 8 + 255) / 256), 256, 0, stream>>>(wav, static_cast<bf16*>(out), S, L0, k, stride, rows);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int conv0_pack_weights(const float* w, void* out, int C, int k, cudaStream_t stream) {
    conv0_pack_w_kernel<<<(C + 127) / 128, 128, 0, stream>>>(w, static_cast<bf16*>(out), C, k);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace slsb

// conv0 of the feature extractor as ONE kernel:  raw audio -> GELU(LayerNorm_512(Conv1d(1 -> 512, k = 10, stride 5)))  (bf16 out)
// (wav2vec2.py:785-822 with extractor_mode="layer_norm", first block: C_in = 1.)
//
// conv0 is epilogue work, not math (0.066 GMAC / utterance against 6.6 M LayerNorm + GELU elements), so the kernel is built around
// the epilogue's instruction count:
//   * The LayerNorm statistics of a row never need the 512 conv outputs.  With a = [x_0 .. x_{k-1}, 1] (the row's audio window and a
//     one for the bias) and W' = [w | b] the outputs are y_n = a . W'_n, hence
//         mean = a . s / 512,  s_i = sum_n W'_ni          E[y^2] = a^T G a / 512,  G_ij = sum_n W'_ni W'_nj
//     (an 11 x 11 Gram matrix, built once per weight load in double precision).  The producer threads evaluate both forms in fp32
//     from the exact audio samples: 132 FMAs per row instead of a first pass over 512 accumulator columns.
//   * Because the statistics do not come from the accumulator, a tile does not have to hold whole rows: the accumulator is
//     128 rows x 256 channels, double-buffered in TMEM, and the epilogue of one half overlaps everything else.
//   * The conv bias rides in two spare K columns (bf16 hi + lo parts against a 1.0 in A), the weights (512 x 64 bf16 = 64 KB) stay
//     resident in shared memory for the life of the CTA, and the A operand (hi/lo bf16 split of the audio window, see below) is
//     built by four producer warps straight into the SWIZZLE_128B shared-memory tile: no im2col buffer, no per-tile weight fetch.
//   * 16 epilogue warps (4 per scheduler), thread == accumulator row, 64 channels per warp and half; per element:
//     2 FFMA (normalise, affine) + the one-MUFU GELU + half a pack; bf16 tiles leave through 8 KB SWIZZLE_64B staging buffers
//     in ping-pong and cp.async.bulk.tensor stores.
// Arithmetic: x = hi + lo (bf16 each), w = w_hi + w_lo; acc = hi.w_hi + hi.w_lo + lo.w_hi + b_hi + b_lo in fp32 on the tensor core
// (the dropped lo.w_lo term is 2^-16 relative), the same split as the round-1 kernel (gemm_tc_ln.cu), which stays as the A/B
// reference (SLSB_CONV0_V1=1).
#include "common.cuh"
#include "kernels.h"

namespace slsb {
namespace {

constexpr int C0_BLOCK_M = 128, C0_K = 64, C0_N = 512, C0_UMMA_K = 16;
constexpr int C0_STAGES = 4;                             // A tiles in flight
constexpr int C0_STAT_RING = 8;                          // row statistics: deeper than the A ring (read by the epilogue, not the MMA)
constexpr int C0_W_BYTES = C0_N * C0_K * 2;              // 64 KB, two 256-row SWIZZLE_128B boxes
constexpr int C0_A_BYTES = C0_BLOCK_M * C0_K * 2;        // 16 KB
constexpr int C0_OFF_A = C0_W_BYTES;
constexpr int C0_OFF_STORE = C0_OFF_A + C0_STAGES * C0_A_BYTES;      // 4 column groups x 2 buffers x [128 rows x 32 bf16] SWIZZLE_64B
constexpr int C0_OFF_PAR = C0_OFF_STORE + 4 * 2 * 8192;              // ln_w | ln_b
constexpr int C0_OFF_STAT = C0_OFF_PAR + 2 * C0_N * 4;               // float2 [ring][128]: (rstd, -mean * rstd)
constexpr int C0_OFF_GRAM = C0_OFF_STAT + C0_STAT_RING * 128 * 8;    // s[16] | G[16][16] (row stride 16 floats)
constexpr int C0_OFF_BAR = C0_OFF_GRAM + (16 + 256) * 4;
constexpr int C0_SMEM = C0_OFF_BAR + 256;
constexpr int C0_PROD_WARPS = 4, C0_EPI_WARPS = 16;
constexpr int C0_THREADS = 32 * (4 + C0_PROD_WARPS + C0_EPI_WARPS);  // warp 0: W load + MMA issue, warp 1: TMEM alloc, warps 2-3 idle

struct Conv0Dev {
    const float* wav; const float* gram; const float* ln_w; const float* ln_b;
    long long rows;          // B * L0
    int S, L0, k, stride, m_tiles;
    float eps;
};

__global__ void __launch_bounds__(C0_THREADS, 1)
conv0_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out, const Conv0Dev p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("slsb: dynamic smem base not 1024-aligned\n"); __trap(); }
    float* par = reinterpret_cast<float*>(smem + C0_OFF_PAR);
    float2* stat = reinterpret_cast<float2*>(smem + C0_OFF_STAT);
    float* gram = reinterpret_cast<float*>(smem + C0_OFF_GRAM);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C0_OFF_BAR);
    uint64_t* empty_bar = full_bar + C0_STAGES;
    uint64_t* tmem_full = empty_bar + C0_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* w_bar = tmem_empty + 2;
    uint64_t* stat_bar = w_bar + 1;                        // [C0_STAT_RING]: the tile's row statistics are published
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(stat_bar + C0_STAT_RING);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    griddep_launch();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w); tma_prefetch_desc(&tmap_out);
        for (int s = 0; s < C0_STAGES; ++s) { mbar_init(&full_bar[s], C0_PROD_WARPS); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], C0_EPI_WARPS); }
        for (int s = 0; s < C0_STAT_RING; ++s) mbar_init(&stat_bar[s], C0_PROD_WARPS);
        mbar_init(w_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    for (int i = threadIdx.x; i < C0_N; i += C0_THREADS) { par[i] = p.ln_w[i]; par[C0_N + i] = p.ln_b[i]; }
    for (int i = threadIdx.x; i < 16 + 256; i += C0_THREADS) gram[i] = p.gram[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== weights once, then the MMA issue loop (one thread) =====================
        if (lane == 0) {
            mbar_expect_tx(w_bar, C0_W_BYTES);
            tma_load_2d(smem, &tmap_w, w_bar, 0, 0);
            tma_load_2d(smem + C0_W_BYTES / 2, &tmap_w, w_bar, 0, 256);
            mbar_wait(w_bar, 0);
            constexpr uint32_t idesc = make_idesc_bf16(C0_BLOCK_M, 256);
            const uint32_t sw = smem_u32(smem);
            int stage = 0; uint32_t phase = 0;
            int unit = 0;
            for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t da = make_smem_desc_sw128(sw + C0_OFF_A + stage * C0_A_BYTES, 0, 1024);
#pragma unroll
                for (int nh = 0; nh < 2; ++nh, ++unit) {
                    const int acc = unit & 1;
                    mbar_wait(&tmem_empty[acc], ((unit >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint64_t db = make_smem_desc_sw128(sw + nh * (C0_W_BYTES / 2), 0, 1024);
#pragma unroll
                    for (int k = 0; k < C0_K / C0_UMMA_K; ++k)
                        tc_mma_f16(tmem_base + acc * 256, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, k != 0 ? 1u : 0u);
                    if (nh == 1) tc_commit(&empty_bar[stage]);       // both halves have read the A tile
                    tc_commit(&tmem_full[acc]);
                }
                if (++stage == C0_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 4 + C0_PROD_WARPS) {
        // ===================== A producers: thread == row of the tile =====================
        const int r = (warp - 4) * 32 + lane;
        griddep_wait();                                   // the audio may come from the previous kernel in the stream (ingest)
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
            const long long row = (long long)tile * C0_BLOCK_M + r;
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = 0.f;
            const bool ok = row < p.rows;
            if (ok) {
                const int b = (int)(row / p.L0), f = (int)(row - (long long)b * p.L0);
                const float* src = p.wav + (long long)b * p.S + (long long)f * p.stride;
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < p.k) x[j] = __ldg(src + j);
            }
            // statistics: a = [x_0 .. x_{k-1}, 1] (the 1 sits at index k); gram = s[16] | G[16][16], zero outside (k + 1)^2
            float mean = 0.f, e2 = 0.f;
            {
                float a[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) a[j] = j < p.k ? x[j] : (j == p.k ? 1.0f : 0.f);
                const int n = p.k + 1;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (i < n) {
                        const float4* g4 = reinterpret_cast<const float4*>(gram + 16 + i * 16);
                        float t = 0.f;
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 g = g4[j4];
                            t = fmaf(g.x, a[4 * j4 + 0], t); t = fmaf(g.y, a[4 * j4 + 1], t);
                            t = fmaf(g.z, a[4 * j4 + 2], t); t = fmaf(g.w, a[4 * j4 + 3], t);
                        }
                        e2 = fmaf(a[i], t, e2);
                        mean = fmaf(a[i], gram[i], mean);
                    }
                }
            }
            const float var = fmaxf(e2 - mean * mean, 0.0f);
            const float rstd = rsqrtf(var + p.eps);
            // A row (64 bf16, eight 16-byte chunks): [hi 0..15 | hi 0..15 | lo 0..15 | 1 1 0 ..] with zeros past the k taps
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bf16 h0 = __float2bfloat16_rn(x[2 * j]), h1 = __float2bfloat16_rn(x[2 * j + 1]);
                const float l0 = x[2 * j] - __bfloat162float(h0), l1 = x[2 * j + 1] - __bfloat162float(h1);
                hi[j] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
                lo[j] = pack_bf16x2(l0, l1);
            }
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* srow = smem + C0_OFF_A + stage * C0_A_BYTES + r * 128;
            const int sw = r & 7;
            const uint4 h_a = make_uint4(hi[0], hi[1], hi[2], hi[3]), h_b = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            const uint4 l_a = make_uint4(lo[0], lo[1], lo[2], lo[3]), l_b = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            const uint32_t one2 = ok ? 0x3F803F80u : 0u;             // bf16 (1.0, 1.0): bias hi + lo; rows past the end stay zero
            *reinterpret_cast<uint4*>(srow + ((0 ^ sw) << 4)) = h_a;
            *reinterpret_cast<uint4*>(srow + ((1 ^ sw) << 4)) = h_b;
            *reinterpret_cast<uint4*>(srow + ((2 ^ sw) << 4)) = h_a;
            *reinterpret_cast<uint4*>(srow + ((3 ^ sw) << 4)) = h_b;
            *reinterpret_cast<uint4*>(srow + ((4 ^ sw) << 4)) = l_a;
            *reinterpret_cast<uint4*>(srow + ((5 ^ sw) << 4)) = l_b;
            *reinterpret_cast<uint4*>(srow + ((6 ^ sw) << 4)) = make_uint4(one2, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(srow + ((7 ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            stat[(it & (C0_STAT_RING - 1)) * 128 + r] = make_float2(rstd, -mean * rstd);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&full_bar[stage]); mbar_arrive(&stat_bar[it & (C0_STAT_RING - 1)]); }
            if (++stage == C0_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 4 + C0_PROD_WARPS) {
        // ===================== epilogue: 16 warps, lane quarter q, column group cg (64 of the half's 256 channels) =====================
        const int w = warp - (4 + C0_PROD_WARPS);
        const int q = warp & 3, cg = w >> 2;
        const int r = q * 32 + lane;
        const int bar_id = 1 + cg;
        const bool issuer = r == 0;
        uint8_t* stage_tile = smem + C0_OFF_STORE + cg * 16384;
        const int swz = (r >> 1) & 3;
        int unit = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
            float2 st = make_float2(0.f, 0.f);
#pragma unroll 1
            for (int nh = 0; nh < 2; ++nh, ++unit) {
                const int acc = unit & 1;
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * 256 + cg * 64;
                const int col_base = nh * 256 + cg * 64;
                if (nh == 0) {
                    // slot (it & 7) is rewritten for tile it + 8, whose A stage frees only after the MMAs of tile it + 4, which in turn
                    // need this warp's release of tile it + 3: the slot cannot change (or its barrier wrap) while it is read here
                    mbar_wait(&stat_bar[it & (C0_STAT_RING - 1)], (it >> 3) & 1);
                    st = stat[(it & (C0_STAT_RING - 1)) * 128 + r];
                }
                mbar_wait(&tmem_full[acc], (unit >> 1) & 1);
                tc_fence_after();
                uint32_t a[2][16];
                tmem_ld_32x32b_x16(taddr, a[0]);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    tmem_ld_wait();
                    if (s + 1 < 4) tmem_ld_32x32b_x16(taddr + (s + 1) * 16, a[(s + 1) & 1]);
                    else {                                               // accumulator half fully read: the MMA thread may reuse it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    }
                    const float* g_s = par + col_base + s * 16;
                    const float* h_s = par + C0_N + col_base + s * 16;
                    uint32_t o[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 gg = *reinterpret_cast<const float4*>(g_s + 4 * j);
                        const float4 hh = *reinterpret_cast<const float4*>(h_s + 4 * j);
                        const float v0 = gelu_fast(fmaf(fmaf(__uint_as_float(a[s & 1][4 * j + 0]), st.x, st.y), gg.x, hh.x));
                        const float v1 = gelu_fast(fmaf(fmaf(__uint_as_float(a[s & 1][4 * j + 1]), st.x, st.y), gg.y, hh.y));
                        const float v2 = gelu_fast(fmaf(fmaf(__uint_as_float(a[s & 1][4 * j + 2]), st.x, st.y), gg.z, hh.z));
                        const float v3 = gelu_fast(fmaf(fmaf(__uint_as_float(a[s & 1][4 * j + 3]), st.x, st.y), gg.w, hh.w));
                        o[2 * j] = pack_bf16x2(v0, v1); o[2 * j + 1] = pack_bf16x2(v2, v3);
                    }
                    // 32-column chunk c = s >> 1 of this half goes to buffer c & 1; 16-byte piece (s & 1) * 2 + {0, 1} of the 64-byte row
                    uint8_t* srow = stage_tile + (s >> 1) * 8192 + r * 64;
                    *reinterpret_cast<uint4*>(srow + ((((s & 1) * 2 + 0) ^ swz) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(srow + ((((s & 1) * 2 + 1) ^ swz) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
                    if (s & 1) {
                        fence_proxy_async_smem();
                        if (issuer) tma_store_wait_read<0>();            // the previous store has read the OTHER buffer: the next chunk may overwrite it
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                        if (issuer) {
                            tma_store_2d(&tmap_out, stage_tile + (s >> 1) * 8192, col_base + (s >> 1) * 32, tile * C0_BLOCK_M);
                            tma_store_commit();
                        }
                    }
                }
            }
        }
        if (issuer) tma_store_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// conv0 weights [C, k] fp32 + bias [C] -> [C, 64] bf16: [w_hi (16) | w_lo (16) | w_hi (16) | b_hi b_lo 0 ..]
__global__ void conv0_pack_w_bias_kernel(const float* __restrict__ w, const float* __restrict__ bias, bf16* __restrict__ out, int C, int k) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int i = 0; i < 64; ++i) out[c * 64 + i] = __float2bfloat16_rn(0.f);
    for (int t = 0; t < k; ++t) {
        const float wv = w[c * k + t];
        const bf16 hi = __float2bfloat16_rn(wv);
        const bf16 lo = __float2bfloat16_rn(wv - __bfloat162float(hi));
        out[c * 64 + t] = hi; out[c * 64 + 16 + t] = lo; out[c * 64 + 32 + t] = hi;
    }
    const float bv = bias[c];
    const bf16 bh = __float2bfloat16_rn(bv);
    out[c * 64 + 48] = bh;
    out[c * 64 + 49] = __float2bfloat16_rn(bv - __bfloat162float(bh));
}

// s_i = sum_n W'_ni / C and G_ij = sum_n W'_ni W'_nj / C over W' = [w | b] (index k = the bias), accumulated in double in the fixed
// order n = 0 .. C-1 (one thread per entry): gram = s[16] | G[16][16], zero outside (k + 1) x (k + 1).
__global__ void conv0_gram_kernel(const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ gram, int C, int k) {
    const int e = threadIdx.x;               // 0 .. 271
    if (e >= 16 + 256) return;
    const int n1 = k + 1;
    double acc = 0.0;
    if (e < 16) {
        if (e < n1) for (int n = 0; n < C; ++n) acc += e < k ? (double)w[n * k + e] : (double)bias[n];
    } else {
        const int i = (e - 16) >> 4, j = (e - 16) & 15;
        if (i < n1 && j < n1)
            for (int n = 0; n < C; ++n) {
                const double a = i < k ? (double)w[n * k + i] : (double)bias[n];
                const double b = j < k ? (double)w[n * k + j] : (double)bias[n];
                acc += a * b;
            }
    }
    gram[e] = (float)(acc / C);
}

}  // namespace

int conv0_tc_pack(const float* w, const float* bias, void* w64, float* gram, int C, int k, cudaStream_t stream) {
    if (C != C0_N || k < 1 || k > 15) { set_error("conv0_tc_pack: needs C == 512 and 1 <= k <= 15 (got C=%d k=%d)", C, k); return -1; }
    conv0_pack_w_bias_kernel<<<(C + 127) / 128, 128, 0, stream>>>(w, bias, static_cast<bf16*>(w64), C, k);
    SLSB_CUDA_CHECK(cudaGetLastError());
    conv0_gram_kernel<<<1, 288, 0, stream>>>(w, bias, gram, C, k);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int conv0_tc(const float* wav, const void* w64, const float* gram, const float* ln_w, const float* ln_b, void* out, int B, int S, int L0, int k,
             int stride, float eps, int num_sms, cudaStream_t stream) {
    if (k < 1 || k > 15) { set_error("conv0_tc: 1 <= k <= 15 (got %d)", k); return -1; }
    if (B <= 0 || L0 <= 0) return 0;
    Conv0Dev dp{};
    dp.wav = wav; dp.gram = gram; dp.ln_w = ln_w; dp.ln_b = ln_b;
    dp.rows = (long long)B * L0; dp.S = S; dp.L0 = L0; dp.k = k; dp.stride = stride; dp.eps = eps;
    dp.m_tiles = (int)((dp.rows + C0_BLOCK_M - 1) / C0_BLOCK_M);
    CUtensorMap tw, to;
    {
        uint64_t dims[2] = {(uint64_t)C0_K, (uint64_t)C0_N};
        uint64_t strides[1] = {(uint64_t)C0_K * 2};
        uint32_t box[2] = {C0_K, 256};
        if (encode_tmap_bf16(&tw, w64, 2, dims, strides, box)) return -1;
    }
    {
        uint64_t dims[2] = {(uint64_t)C0_N, (uint64_t)dp.rows};
        uint64_t strides[1] = {(uint64_t)C0_N * 2};
        uint32_t box[2] = {32, C0_BLOCK_M};
        if (encode_tmap_bf16_sw64(&to, out, 2, dims, strides, box)) return -1;
    }
    static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
    if (first_use_on_device(&configured_on)) {
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(conv0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C0_SMEM));
    }
    const int grid = dp.m_tiles < num_sms ? dp.m_tiles : num_sms;
    SLSB_CUDA_CHECK(launch_pdl(conv0_tc_kernel, dim3(grid), dim3(C0_THREADS), C0_SMEM, stream, tw, to, dp));
    return 0;
}

}  // namespace slsb

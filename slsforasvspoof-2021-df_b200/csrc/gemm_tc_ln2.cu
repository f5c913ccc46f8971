// Feature-extractor layer as ONE tensor-core kernel, CTA-pair version:
//     out = GELU(LayerNorm_512(A * W^T + bias))   (bf16 out; wav2vec2.py:785-822, extractor_mode="layer_norm")
//
// gemm_tc_ln.cu keeps a whole 128 x 512 LayerNorm row block in one CTA's TMEM, which fills all 512 columns: the MMA of
// the next tile cannot start before the epilogue has drained the accumulator, so mainloop and epilogue alternate
// (measured: 11.7 us of epilogue + 20 us of mainloop per conv1 tile).  Here a CLUSTER OF TWO CTAs shares one row block:
//   * CTA r of the pair owns channels [256 r, 256 r + 256): a 128 x 256 accumulator, so TMEM holds TWO of them and the
//     epilogue of tile i overlaps the MMAs of tile i + 1 (same pipeline as gemm_tc.cu: 4-stage TMA ring, 48 KB / stage);
//   * the LayerNorm statistics of a row need all 512 channels: each CTA reduces its 256 columns to (sum, sum of squares),
//     writes the pair into the PEER's shared memory (st.shared::cluster) and arrives on the peer's mbarrier; after the
//     exchange both CTAs hold the full-row mean / rstd and normalise their own half;
//   * epilogue: 8 warps, thread == row, two column halves of 128; pass 1 statistics, pass 2 affine + GELU -> bf16 ->
//     SWIZZLE_128B staging tile -> TMA store.  The pre-norm activations never touch HBM.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace slsb {
namespace {

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16, NCH = 512, NHALF = 256;
constexpr int kStages = 4;
constexpr int kStageA = BLOCK_M * BLOCK_K * 2;          // 16 KB
constexpr int kStageB = NHALF * BLOCK_K * 2;            // 32 KB
constexpr int kStage = kStageA + kStageB;               // 48 KB
constexpr int kStoreOffset = kStages * kStage;          // 2 x [128 rows x 64 bf16] SWIZZLE_128B staging tiles (one per column half)
constexpr int kXchOffset = kStoreOffset + 2 * 16384;    // float2 xch[2 buffers][128 rows]: written by the PEER CTA
constexpr int kBarOffset = kXchOffset + 2 * 128 * 8;
constexpr int kSmemBytes = kBarOffset + 256;
constexpr int kThreads = 384;
static_assert(kSmemBytes <= 232448, "shared memory budget");

struct Ln2Dev {
    int M, K, batches, m_tiles;
    int conv_cin, conv_stride;
    const float* bias; const float* ln_w; const float* ln_b;
    float eps;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_peer(uint32_t smem_addr, uint32_t peer) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(peer));
    return r;
}
__device__ __forceinline__ void st_peer_f2(uint32_t cluster_addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_peer(uint32_t cluster_bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > SLSB_MBAR_TIMEOUT_CYCLES) {
            printf("slsb: cluster mbarrier timeout block=%d thread=%d parity=%u\n", (int)blockIdx.x, (int)threadIdx.x, parity);
            __trap();
        }
    }
}

template <int A_MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_gemm_ln2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_out, const Ln2Dev p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("slsb: dynamic smem base not 1024-aligned\n"); __trap(); }
    float2* part = reinterpret_cast<float2*>(smem + kStoreOffset);   // in-CTA combine scratch [2 halves][128 rows]: borrows the first
                                                                      // 2 KB of half 0's staging tile between two of its stores
    float2* xch = reinterpret_cast<float2*>(smem + kXchOffset);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;           // [2]
    uint64_t* xch_full = tmem_empty + 2;            // [2] the peer's partial statistics for buffer b have landed (4 warp arrivals)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xch_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();        // which 256-channel half of the row block this CTA owns
    const uint32_t peer = rank ^ 1u;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_tiles = p.batches * p.m_tiles;
    const int num_kb = p.K / BLOCK_K;

    griddep_launch();
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); tma_prefetch_desc(&tmap_out); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); mbar_init(&xch_full[s], 4); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    cluster_sync_all();                 // both CTAs' barriers are initialised before anybody arrives remotely
    griddep_wait();                     // prologue overlapped the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
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
                    tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, (int)rank * NHALF);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, NHALF);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * NHALF;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStage);
                    const uint64_t da = make_smem_desc_sw128(sa, 0, 1024);
                    const uint64_t db = make_smem_desc_sw128(sa + kStageA, 0, 1024);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        tc_mma_f16(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&empty_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tmem_full[acc]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps: thread == row, column half = 128 of this CTA's 256 channels =====================
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const int ch0 = (int)rank * NHALF + half * 128;                 // first absolute channel of this thread's columns
        const float* bias_g = p.bias + ch0;
        const float* g_g = p.ln_w + ch0;
        const float* h_g = p.ln_b + ch0;
        uint8_t* stage_tile = smem + kStoreOffset + half * 16384;
        uint8_t* srow = stage_tile + r * 128;
        const int bar_id = 2 + half;
        const uint32_t peer_xch = map_to_peer(smem_u32(xch), peer);
        const uint32_t peer_bar = map_to_peer(smem_u32(xch_full), peer);
        int it = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
            const int m_blk = tile % p.m_tiles, b = tile / p.m_tiles;
            const int acc = it & 1, buf = it & 1;
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * NHALF + half * 128;
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            tc_fence_after();
            // ---- pass 1: (sum, sum of squares) over this thread's 128 columns; TMEM loads double-buffered in registers
            float s = 0.f, ss = 0.f;
            {
                uint32_t a[2][32];
                tmem_ld_32x32b_x32(taddr, a[0]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c + 1 < 4) tmem_ld_32x32b_x32(taddr + (c + 1) * 32, a[(c + 1) & 1]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias_g + c * 32 + 4 * j));
                        const float v0 = __uint_as_float(a[c & 1][4 * j + 0]) + bb.x, v1 = __uint_as_float(a[c & 1][4 * j + 1]) + bb.y;
                        const float v2 = __uint_as_float(a[c & 1][4 * j + 2]) + bb.z, v3 = __uint_as_float(a[c & 1][4 * j + 3]) + bb.w;
                        s += (v0 + v1) + (v2 + v3);
                        ss = fmaf(v0, v0, ss); ss = fmaf(v1, v1, ss); ss = fmaf(v2, v2, ss); ss = fmaf(v3, v3, ss);
                    }
                    tmem_ld_wait();
                }
            }
            // ---- combine the two column halves inside the CTA, then exchange the 256-channel partial with the peer CTA
            if (half == 0 && r == 0) tma_store_wait_read<0>();        // half 0's last store no longer reads the tile the scratch borrows
            asm volatile("bar.sync 1, 256;" ::: "memory");            // (the 8 epilogue warps only)
            part[half * 128 + r] = make_float2(s, ss);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float2 other = part[(half ^ 1) * 128 + r];
            asm volatile("bar.sync 1, 256;" ::: "memory");            // scratch dead: pass 2 may overwrite the staging tile
            const float cs = s + other.x, css = ss + other.y;          // this CTA's 256 channels
            if (half == 0) {
                st_peer_f2(peer_xch + (uint32_t)(buf * 128 + r) * 8u, cs, css);
                __syncwarp();
                if (lane == 0) mbar_arrive_peer(peer_bar + (uint32_t)buf * 8u);
            }
            mbar_wait_cluster(&xch_full[buf], (it >> 1) & 1);
            const float2 px = xch[buf * 128 + r];
            const float mean = (cs + px.x) * (1.0f / NCH);
            const float var = fmaxf((css + px.y) * (1.0f / NCH) - mean * mean, 0.0f);
            const float rstd = rsqrtf(var + p.eps);
            // ---- pass 2: normalise + affine + GELU -> bf16 staging tile (64 columns at a time) -> TMA store
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                uint32_t a[2][32];
                tmem_ld_32x32b_x32(taddr + g * 64, a[0]);
                tmem_ld_32x32b_x32(taddr + g * 64 + 32, a[1]);
                if (r == 0) tma_store_wait_read<0>();
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                tmem_ld_wait();
                if (g == 1) {                                          // accumulator fully read -> hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c0 = g * 64 + h * 32;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias_g + c0 + 4 * j));
                        const float4 gg = __ldg(reinterpret_cast<const float4*>(g_g + c0 + 4 * j));
                        const float4 hh = __ldg(reinterpret_cast<const float4*>(h_g + c0 + 4 * j));
                        v[4 * j + 0] = gelu_fast(fmaf((__uint_as_float(a[h][4 * j + 0]) + bb.x - mean) * rstd, gg.x, hh.x));
                        v[4 * j + 1] = gelu_fast(fmaf((__uint_as_float(a[h][4 * j + 1]) + bb.y - mean) * rstd, gg.y, hh.y));
                        v[4 * j + 2] = gelu_fast(fmaf((__uint_as_float(a[h][4 * j + 2]) + bb.z - mean) * rstd, gg.z, hh.z));
                        v[4 * j + 3] = gelu_fast(fmaf((__uint_as_float(a[h][4 * j + 3]) + bb.w - mean) * rstd, gg.w, hh.w));
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
                    if constexpr (A_MODE == A_CONV) tma_store_3d(&tmap_out, stage_tile, ch0 + g * 64, m_blk * BLOCK_M, b);
                    else tma_store_2d(&tmap_out, stage_tile, ch0 + g * 64, m_blk * BLOCK_M);
                    tma_store_commit();
                }
            }
        }
        if (r == 0) tma_store_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer may still be writing into this CTA's exchange buffer until it is done too
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int A_MODE>
int launch_ln2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const Ln2Dev& dp, int num_sms, cudaStream_t stream) {
    static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
    static int max_pairs = 0;
    auto kern = tc_gemm_ln2_kernel<A_MODE>;
    if (first_use_on_device(&configured_on)) {
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        // persistent grid = as many 2-CTA clusters as can be co-resident (a GPC with an odd SM count leaves one SM unpaired)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms & ~1); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms / 2; }
        max_pairs = n < num_sms / 2 ? n : num_sms / 2;
    }
    const int tiles = dp.batches * dp.m_tiles;
    const int pairs = tiles < max_pairs ? tiles : max_pairs;
    SLSB_CUDA_CHECK(launch_pdl(kern, dim3(2 * pairs), dim3(kThreads), kSmemBytes, stream, ta, tb, to, dp));
    return 0;
}

}  // namespace

int tc_gemm_ln_gelu_pair(const TcLnGemmArgs& g, int num_sms, cudaStream_t stream) {
    if (g.N != NCH) { set_error("tc_gemm_ln_gelu_pair: N must be 512 (got %d)", g.N); return -1; }
    if (g.K % BLOCK_K != 0 || g.K <= 0) { set_error("tc_gemm_ln_gelu_pair: K=%d must be a positive multiple of 64", g.K); return -1; }
    if (g.M <= 0 || g.batches <= 0) return 0;
    Ln2Dev dp{};
    dp.M = g.M; dp.K = g.K; dp.batches = g.batches; dp.m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
    dp.conv_cin = g.conv_cin; dp.conv_stride = g.conv_stride;
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
        uint32_t box[2] = {BLOCK_K, NHALF};
        if (encode_tmap_bf16(&tb, g.W, 2, dims, strides, box)) return -1;
    }
    if (g.a_mode == A_PLAIN) {
        uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.M};
        uint64_t strides[1] = {(uint64_t)g.lda * 2};
        uint32_t box[2] = {BLOCK_K, BLOCK_M};
        if (encode_tmap_bf16(&ta, g.A, 2, dims, strides, box)) return -1;
        return launch_ln2<A_PLAIN>(ta, tb, to, dp, num_sms, stream);
    }
    const uint64_t C = g.conv_cin, s = g.conv_stride, Lin = g.conv_lin;
    uint64_t dims[4] = {C, s, (Lin + s - 1) / s, (uint64_t)g.batches};
    uint64_t strides[3] = {C * 2, s * C * 2, Lin * C * 2};
    uint32_t box[4] = {BLOCK_K, 1, BLOCK_M, 1};
    if (encode_tmap_bf16(&ta, g.A, 4, dims, strides, box)) return -1;
    return launch_ln2<A_CONV>(ta, tb, to, dp, num_sms, stream);
}

}  // namespace slsb

// Feature-extractor layer as ONE tensor-core kernel, CTA-pair version with SIXTEEN epilogue warps (conv1..6, the default):
//     out = GELU(LayerNorm_512(A * W^T + bias))   (bf16 out; wav2vec2.py:785-822, extractor_mode="layer_norm")
//
// Mainloop, cluster layout and statistics exchange are those of gemm_tc_ln2.cu (CTA r of a 2-CTA cluster owns channels
// [256 r, 256 r + 256) of a 128-row block, two 128 x 256 accumulators in TMEM, (sum, sum of squares) exchanged through distributed
// shared memory).  Here 16 epilogue warps (4 per scheduler) take 64 columns each instead of 8 warps with 128:
//   * pass 1: (sum, sum of squares) over 64 columns, 16-column TMEM loads double-buffered; the four column groups of a row combine
//     through 4 KB of scratch borrowed from column group 0's staging buffer, then the pair exchange as before;
//   * pass 2: normalise + affine + one-MUFU GELU -> bf16 -> 8 KB SWIZZLE_64B staging buffer per column group (32 columns at a time)
//     -> TMA store.
// Measured (SLSB_LN2_TRACE=1 python tools/conv_trace.py, conv1: K = 1536, 24 k-blocks per tile): the epilogue of a tile takes ~11 k
// cycles (pass 1 1.5 k, statistics exchange 2.7 k, pass 2 7 k) and hides completely behind the next tile's mainloop, which runs
// 783 clocks per k-block against 542 at the tensor peak: the 128 x 256 x 64 cta_group::1 step moves 48 KB INTO shared memory (TMA)
// and reads 48 KB out of it (MMA operands), 96 KB per k-block at 128 B/clk = 750 clocks - the kernel sits on the shared-memory
// bandwidth of this MMA shape (65 % tensor-pipe activity in ncu), not on the L2 feed (a multicast A tile changed nothing), not on
// the epilogue (8 -> 16 warps: 573 -> 561 us for conv1) and a four-CTA cta_group::2 variant (80 KB per k-block and SM) lost more
// to cluster placement and 256-row tiles than it gained.
// 640 threads, 96 registers.  SLSB_LN_GEMM_EPI8=1 selects the 8-warp kernel (A/B).
#include "common.cuh"
#include "kernels.h"
#include <cstdio>
#include <cstdlib>

namespace slsb {
namespace {

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16, NCH = 512, NHALF = 256;
constexpr int kStages = 4;
constexpr int kStageA = BLOCK_M * BLOCK_K * 2;          // 16 KB
constexpr int kStageB = NHALF * BLOCK_K * 2;            // 32 KB
constexpr int kStage = kStageA + kStageB;               // 48 KB
constexpr int kStoreOffset = kStages * kStage;          // 4 x [128 rows x 32 bf16] SWIZZLE_64B staging buffers (one per 64-channel column group)
constexpr int kXchOffset = kStoreOffset + 4 * 8192;     // float2 xch[2 buffers][128 rows]: written by the PEER CTA
constexpr int kBarOffset = kXchOffset + 2 * 128 * 8;
constexpr int kSmemBytes = kBarOffset + 256;
constexpr int kEpiWarps = 16;
constexpr int kThreads = 32 * (4 + kEpiWarps);
static_assert(kSmemBytes <= 232448, "shared memory budget");

struct Ln2xDev {
    int M, K, batches, m_tiles;
    int conv_cin, conv_stride;
    const float* bias; const float* ln_w; const float* ln_b;
    float eps;
    long long* trace;   // SLSB_LN2_TRACE=1 (tuning only): clock64 stamps of pair 0 / CTA 0, trace[it * 8 + event]
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
tc_gemm_ln2x_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_out, const Ln2xDev p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("slsb: dynamic smem base not 1024-aligned\n"); __trap(); }
    float2* part = reinterpret_cast<float2*>(smem + kStoreOffset);   // in-CTA combine scratch [4 groups][128 rows]: borrows the first
                                                                      // 4 KB of group 0's staging buffer between two of its stores
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
    // events: 0 accumulator free (MMA), 1 first k-block landed, 2 last MMA issued, 3 accumulator complete seen by the epilogue,
    // 4 pass 1 done, 5 peer statistics received, 6 epilogue done, 7 first load of the tile issued
#define LN2_TRACE(it_, ev_) do { if (p.trace != nullptr && pair == 0 && rank == 0 && (it_) < 64) p.trace[(it_) * 8 + (ev_)] = clock64(); } while (0)

    griddep_launch();
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); tma_prefetch_desc(&tmap_out); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpiWarps); mbar_init(&xch_full[s], 4); }
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
            int itp = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++itp) {
                const int m_blk = tile % p.m_tiles, b = tile / p.m_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (kb == 0) LN2_TRACE(itp, 7);
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
                LN2_TRACE(it, 0);
                const uint32_t d_tmem = tmem_base + acc * NHALF;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (kb == 0) LN2_TRACE(it, 1);
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
                LN2_TRACE(it, 2);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps: thread == row, column group cg = 64 of this CTA's 256 channels =====================
        const int q = warp & 3, cg = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const int ch0 = (int)rank * NHALF + cg * 64;                    // first absolute channel of this thread's columns
        const float* bias_g = p.bias + ch0;
        const float* g_g = p.ln_w + ch0;
        const float* h_g = p.ln_b + ch0;
        uint8_t* stage_buf = smem + kStoreOffset + cg * 8192;
        uint8_t* srow = stage_buf + r * 64;
        const int swz = (r >> 1) & 3;                                    // SWIZZLE_64B: 16-byte piece c of row r sits at c ^ ((r >> 1) & 3)
        const int bar_id = 2 + cg;
        const bool issuer = r == 0;
        const uint32_t peer_xch = map_to_peer(smem_u32(xch), peer);
        const uint32_t peer_bar = map_to_peer(smem_u32(xch_full), peer);
        int it = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
            const int m_blk = tile % p.m_tiles, b = tile / p.m_tiles;
            const int acc = it & 1, buf = it & 1;
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * NHALF + cg * 64;
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            tc_fence_after();
            if (warp == 4 && lane == 0) LN2_TRACE(it, 3);
            // ---- pass 1: (sum, sum of squares) over this thread's 64 columns; 16-column TMEM loads double-buffered in registers
            float s = 0.f, ss = 0.f;
            {
                uint32_t a[2][16];
                tmem_ld_32x32b_x16(taddr, a[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    tmem_ld_wait();
                    if (c + 1 < 4) tmem_ld_32x32b_x16(taddr + (c + 1) * 16, a[(c + 1) & 1]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias_g + c * 16 + 4 * j));
                        const float v0 = __uint_as_float(a[c & 1][4 * j + 0]) + bb.x, v1 = __uint_as_float(a[c & 1][4 * j + 1]) + bb.y;
                        const float v2 = __uint_as_float(a[c & 1][4 * j + 2]) + bb.z, v3 = __uint_as_float(a[c & 1][4 * j + 3]) + bb.w;
                        s += (v0 + v1) + (v2 + v3);
                        ss = fmaf(v0, v0, ss); ss = fmaf(v1, v1, ss); ss = fmaf(v2, v2, ss); ss = fmaf(v3, v3, ss);
                    }
                }
            }
            if (warp == 4 && lane == 0) LN2_TRACE(it, 4);
            // ---- combine the four column groups inside the CTA (fixed order 0, 1, 2, 3), then exchange the 256-channel partial with the peer
            if (cg == 0 && issuer) tma_store_wait_read<0>();          // group 0's last store no longer reads the buffer the scratch borrows
            asm volatile("bar.sync 1, 512;" ::: "memory");            // (the 16 epilogue warps only)
            part[cg * 128 + r] = make_float2(s, ss);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            const float2 p0 = part[r], p1 = part[128 + r], p2 = part[256 + r], p3 = part[384 + r];
            asm volatile("bar.sync 1, 512;" ::: "memory");            // scratch dead: pass 2 may overwrite the staging buffer
            const float cs = (p0.x + p1.x) + (p2.x + p3.x), css = (p0.y + p1.y) + (p2.y + p3.y);     // this CTA's 256 channels
            if (cg == 0) {
                st_peer_f2(peer_xch + (uint32_t)(buf * 128 + r) * 8u, cs, css);
                __syncwarp();
                if (lane == 0) mbar_arrive_peer(peer_bar + (uint32_t)buf * 8u);
            }
            mbar_wait_cluster(&xch_full[buf], (it >> 1) & 1);
            if (warp == 4 && lane == 0) LN2_TRACE(it, 5);
            const float2 px = xch[buf * 128 + r];
            const float mean = (cs + px.x) * (1.0f / NCH);
            const float var = fmaxf((css + px.y) * (1.0f / NCH) - mean * mean, 0.0f);
            const float rstd = rsqrtf(var + p.eps);
            const float nmr = -mean * rstd;
            // ---- pass 2: normalise + affine + GELU -> bf16 -> staging buffer (32 columns at a time) -> TMA store
            {
                uint32_t a[2][16];
                tmem_ld_32x32b_x16(taddr, a[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    tmem_ld_wait();
                    if (c + 1 < 4) tmem_ld_32x32b_x16(taddr + (c + 1) * 16, a[(c + 1) & 1]);
                    else {                                             // accumulator fully read -> hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    }
                    uint32_t o[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias_g + c * 16 + 4 * j));
                        const float4 gg = __ldg(reinterpret_cast<const float4*>(g_g + c * 16 + 4 * j));
                        const float4 hh = __ldg(reinterpret_cast<const float4*>(h_g + c * 16 + 4 * j));
                        const float v0 = gelu_fast(fmaf(fmaf(__uint_as_float(a[c & 1][4 * j + 0]) + bb.x, rstd, nmr), gg.x, hh.x));
                        const float v1 = gelu_fast(fmaf(fmaf(__uint_as_float(a[c & 1][4 * j + 1]) + bb.y, rstd, nmr), gg.y, hh.y));
                        const float v2 = gelu_fast(fmaf(fmaf(__uint_as_float(a[c & 1][4 * j + 2]) + bb.z, rstd, nmr), gg.z, hh.z));
                        const float v3 = gelu_fast(fmaf(fmaf(__uint_as_float(a[c & 1][4 * j + 3]) + bb.w, rstd, nmr), gg.w, hh.w));
                        o[2 * j] = pack_bf16x2(v0, v1); o[2 * j + 1] = pack_bf16x2(v2, v3);
                    }
                    if ((c & 1) == 0) {                                // first half of a 32-column chunk: the buffer must be free
                        if (issuer) tma_store_wait_read<0>();
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    }
                    *reinterpret_cast<uint4*>(srow + ((((c & 1) * 2 + 0) ^ swz) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(srow + ((((c & 1) * 2 + 1) ^ swz) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
                    if (c & 1) {
                        fence_proxy_async_smem();
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                        if (issuer) {
                            if constexpr (A_MODE == A_CONV) tma_store_3d(&tmap_out, stage_buf, ch0 + (c >> 1) * 32, m_blk * BLOCK_M, b);
                            else tma_store_2d(&tmap_out, stage_buf, ch0 + (c >> 1) * 32, m_blk * BLOCK_M);
                            tma_store_commit();
                        }
                    }
                }
            }
            if (warp == 4 && lane == 0) LN2_TRACE(it, 6);
        }
        if (issuer) tma_store_wait<0>();
    }
#undef LN2_TRACE
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer may still be writing into this CTA's exchange buffer until it is done too
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int A_MODE>
int launch_ln2x(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const Ln2xDev& dp, int num_sms, cudaStream_t stream) {
    static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
    static int max_pairs = 0;
    auto kern = tc_gemm_ln2x_kernel<A_MODE>;
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
    static const bool trace_on = getenv("SLSB_LN2_TRACE") && atoi(getenv("SLSB_LN2_TRACE")) != 0;
    if (trace_on) {   // tuning aid: timeline of pair 0, printed after a stream sync (never on in production)
        Ln2xDev q = dp;
        static long long tr_buf[64 * 8];
        long long* tr_dev = nullptr;
        SLSB_CUDA_CHECK(cudaMalloc(&tr_dev, sizeof(tr_buf)));
        SLSB_CUDA_CHECK(cudaMemsetAsync(tr_dev, 0, sizeof(tr_buf), stream));
        q.trace = tr_dev;
        SLSB_CUDA_CHECK(launch_pdl(kern, dim3(2 * pairs), dim3(kThreads), kSmemBytes, stream, ta, tb, to, q));
        SLSB_CUDA_CHECK(cudaMemcpyAsync(tr_buf, tr_dev, sizeof(tr_buf), cudaMemcpyDeviceToHost, stream));
        SLSB_CUDA_CHECK(cudaStreamSynchronize(stream));
        long long t0 = 0;
        for (int i = 0; i < 64 * 8; ++i) if (tr_buf[i] && (!t0 || tr_buf[i] < t0)) t0 = tr_buf[i];
        fprintf(stderr, "ln2x_trace M=%d K=%d batches=%d pairs=%d tiles=%d\n", dp.M, dp.K, dp.batches, pairs, tiles);
        for (int it = 0; it < 12; ++it) {
            if (!tr_buf[it * 8 + 0]) break;
            fprintf(stderr, "  it %2d: load_first %7lld | acc_free %7lld first_kb %7lld mma_done_issue %7lld | acc_ready %7lld pass1 %7lld stats %7lld epi_done %7lld\n", it,
                    tr_buf[it * 8 + 7] - t0, tr_buf[it * 8 + 0] - t0, tr_buf[it * 8 + 1] - t0, tr_buf[it * 8 + 2] - t0, tr_buf[it * 8 + 3] - t0,
                    tr_buf[it * 8 + 4] - t0, tr_buf[it * 8 + 5] - t0, tr_buf[it * 8 + 6] - t0);
        }
        cudaFree(tr_dev);
        return 0;
    }
    SLSB_CUDA_CHECK(launch_pdl(kern, dim3(2 * pairs), dim3(kThreads), kSmemBytes, stream, ta, tb, to, dp));
    return 0;
}

}  // namespace

int tc_gemm_ln_gelu_pair16(const TcLnGemmArgs& g, int num_sms, cudaStream_t stream) {
    if (g.N != NCH) { set_error("tc_gemm_ln_gelu_pair16: N must be 512 (got %d)", g.N); return -1; }
    if (g.K % BLOCK_K != 0 || g.K <= 0) { set_error("tc_gemm_ln_gelu_pair16: K=%d must be a positive multiple of 64", g.K); return -1; }
    if (g.M <= 0 || g.batches <= 0) return 0;
    Ln2xDev dp{};
    dp.M = g.M; dp.K = g.K; dp.batches = g.batches; dp.m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
    dp.conv_cin = g.conv_cin; dp.conv_stride = g.conv_stride;
    dp.bias = g.bias; dp.ln_w = g.ln_w; dp.ln_b = g.ln_b; dp.eps = g.eps;
    CUtensorMap ta, tb, to;
    if (g.a_mode == A_PLAIN) {
        uint64_t dims[2] = {(uint64_t)NCH, (uint64_t)g.M};
        uint64_t strides[1] = {(uint64_t)NCH * 2};
        uint32_t box[2] = {32, BLOCK_M};
        if (encode_tmap_bf16_sw64(&to, g.out, 2, dims, strides, box)) return -1;
    } else {
        uint64_t dims[3] = {(uint64_t)NCH, (uint64_t)g.M, (uint64_t)g.batches};
        uint64_t strides[2] = {(uint64_t)NCH * 2, (uint64_t)g.out_batch_stride * 2};
        uint32_t box[3] = {32, BLOCK_M, 1};
        if (encode_tmap_bf16_sw64(&to, g.out, 3, dims, strides, box)) return -1;
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
        return launch_ln2x<A_PLAIN>(ta, tb, to, dp, num_sms, stream);
    }
    const uint64_t C = g.conv_cin, s = g.conv_stride, Lin = g.conv_lin;
    uint64_t dims[4] = {C, s, (Lin + s - 1) / s, (uint64_t)g.batches};
    uint64_t strides[3] = {C * 2, s * C * 2, Lin * C * 2};
    uint32_t box[4] = {BLOCK_K, 1, BLOCK_M, 1};
    if (encode_tmap_bf16(&ta, g.A, 4, dims, strides, box)) return -1;
    return launch_ln2x<A_CONV>(ta, tb, to, dp, num_sms, stream);
}

}  // namespace slsb

// bf16 tensor-core GEMM for sm_100a:  C = act(A * W^T + bias) (+ residual)
//
//   * operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) into a multi-stage smem ring,
//   * tcgen05.mma (cta_group::1, kind::f16, 128 x BLOCK_N x 16) issued by ONE thread, fp32 accumulators in TMEM,
//   * two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1,
//   * persistent CTAs (grid = #SMs), static round-robin tile schedule,
//   * 8 epilogue warps read TMEM with tcgen05.ld and fuse bias / GELU / ReLU / fp32 residual.
//
// Three A-operand addressing modes share the kernel (the W operand is always a plain [N, K] K-major matrix,
// which is exactly nn.Linear's weight layout):
//   A_PLAIN : A is [M, K] row-major.                                  (q/k/v, out_proj, fc1, fc2, proj, SAE encoder)
//   A_CONV  : implicit GEMM for Conv1d(C->N, k taps, stride s) over channels-last activations [B, L_in, C]:
//             output row l reads the contiguous span x[b, s*l .. s*l+k-1, :]; a 4-D tensor map
//             (c, l_in % s, l_in / s, b) expresses every tap as a plain box -> no im2col buffer.
//             (replaces the cuDNN call behind wav2vec2.py:795, :824-841)
//   A_POS   : grouped positional conv (wav2vec2.py:862-875): for group g and tap t the A box is rows
//             [m0 + t, m0 + t + 128) x channels [64 g, 64 g + 64) of the zero-padded [B, T+128, 1024] stream.
//             (128 x 64 tiles: the N = 64 MMA is shared-memory-bound; kept for A/B and odd geometries.)
//   A_POS4  : the same convolution as a Toeplitz GEMM with full-width tiles: four consecutive output frames t = 4i + j share
//             one accumulator row, column (j, n) = 64 j + n, so N = 256; k-block u (0 .. K+2) multiplies input frame 4i + u
//             with weight tap u - j (zero outside [0, K): the 3-D weight map's out-of-range fill).  A rows = 2 utterances x 64
//             slots of i through a 4-D map (c, t % 4, t / 4, b) (T > 256: several 64-slot blocks per utterance pair); B = four
//             64-row boxes of the un-expanded weights.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>
#include <cstring>

namespace slsb {

namespace {

#ifndef SLSB_STAGES256
#define SLSB_STAGES256 4
#endif
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;       // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 128 + kNumEpiWarps * 32;   // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 spare, warps 4.. epilogue

template <int BLOCK_N> struct SmemPlan {
    static constexpr int kStageA = BLOCK_M * BLOCK_K * 2;
    static constexpr int kStageB = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStage = kStageA + kStageB;
    static constexpr int kStages = (BLOCK_N == 256) ? SLSB_STAGES256 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int kStoreOffset = kStages * kStage;           // 2 x [128 rows x 64 bf16] SWIZZLE_128B staging tiles for TMA stores
    static constexpr int kStoreBytes = (BLOCK_N >= 128) ? 2 * 16384 : 0;
    static constexpr int kBiasOffset = kStoreOffset + kStoreBytes;    // this tile's bias slice
    static constexpr int kBarOffset = kBiasOffset + BLOCK_N * 4;
    static constexpr int kBytes = kBarOffset + 256 /*barriers + tmem ptr*/;
};

struct DevParams {
    int M, N, K;
    int batches, m_tiles, n_tiles;
    int conv_cin, conv_stride;
    int pos_T, pos_B;  // A_POS4: frames per utterance / utterances
    int pos_sb;        // A_POS4: 256-frame slot blocks per utterance (m_blk = pair * pos_sb + slot block)
    void* out;
    long long ldc, out_batch_stride;
    const float* bias;
    const float* residual;
    long long ldr, res_batch_stride;
    int act, out_bf16;
    int kb_per_split;  // A_PLAIN split-K: batch index b selects k-blocks [b * kb_per_split, ...) and partial-output slab b (0 = off)
    int tma_store;     // bf16 output leaves through smem staging + cp.async.bulk.tensor stores
    int red_add;       // fp32 output aliases the fp32 residual: out += acc + bias through TMA reduce-add stores (no residual loads); 2 = 16-column ping-pong
    int res_tma;       // fp32 output = acc + bias + fp32 residual, residual tile fetched by TMA into the staging tile, summed in place, TMA-stored
    int tail_split;    // pair kernel: tiles of the partial last round are cut into this many column slices (1, 2 or 4)
    int debug_flags;   // bit0: epilogue does everything except the global stores / residual loads (mainloop ceiling measurements)
    long long* trace;  // SLSB_GEMM_TRACE=1 (pair kernel, tuning only): clock64 stamps of pair 0, trace[it * 8 + event]
};

// One warp's share of a tile: kCols accumulator columns of its 32 TMEM lanes (thread == output row).
// The fp32 residual of chunk c+1 is fetched into registers while chunk c is converted and stored (the strided
// row-per-thread global reads are latency-, not bandwidth-limited), and the first chunk's residual is requested
// before the accumulator is even complete.
template <int ACT, bool OUT_BF16, bool HAS_RES, int kCols>
__device__ __forceinline__ void epilogue_tile(const DevParams& p, uint32_t taddr, long long out_off, long long res_off, int col_base,
                                              bool row_ok, uint64_t* full_bar, uint32_t full_parity) {
    constexpr int kChunks = kCols / 32;
    if (p.debug_flags & 1) row_ok = false;
    float4 rnext[8];
    if constexpr (HAS_RES) {
        if (row_ok) {
            const float4* r4 = reinterpret_cast<const float4*>(p.residual + res_off + col_base);
#pragma unroll
            for (int j = 0; j < 8; ++j) rnext[j] = __ldg(r4 + j);
        }
    }
    mbar_wait(full_bar, full_parity);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(taddr + c * 32, acc);
        float4 rcur[8];
        if constexpr (HAS_RES) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rcur[j] = rnext[j];
            if (row_ok && c + 1 < kChunks) {
                const float4* r4 = reinterpret_cast<const float4*>(p.residual + res_off + col_base + (c + 1) * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) rnext[j] = __ldg(r4 + j);
            }
        }
        const int col0 = col_base + c * 32;
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + b.x;
            v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b.y;
            v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b.z;
            v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b.w;
        }
        if constexpr (ACT == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
        } else if constexpr (ACT == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (!row_ok) continue;
        if constexpr (HAS_RES) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[4 * j + 0] += rcur[j].x; v[4 * j + 1] += rcur[j].y; v[4 * j + 2] += rcur[j].z; v[4 * j + 3] += rcur[j].w;
            }
        }
        if constexpr (OUT_BF16) {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + out_off + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 w;
                w.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                w.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                w.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                w.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                o[j] = w;
            }
        } else {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    }
}

// bf16 output through shared memory + TMA store: each column half (4 warps = 128 rows) converts 64 columns at a time into a
// SWIZZLE_128B staging tile and one thread issues a cp.async.bulk.tensor store.  Rows past M / L_out are clipped by the tensor
// map, global writes are full 128-byte lines, and the L1/LSU never sees the 128 different rows of a tile.
template <int ACT, int kCols, bool kConvOut>
__device__ __forceinline__ void epilogue_tile_tma(const DevParams& p, const CUtensorMap* tmap_out, uint32_t taddr, uint8_t* stage_tile,
                                                  const float* bias_s, int half, int r, int col_base, int row0, int b,
                                                  uint64_t* full_bar, uint32_t full_parity, uint32_t empty_bar_addr,
                                                  int kGroups = kCols / 64) {
    // kGroups < kCols / 64: a column slice of a tile (pair kernel tail); 0 = this half has no columns, it only releases the accumulator
    const int bar_id = 1 + half;
    const bool issuer = r == 0;
    mbar_wait(full_bar, full_parity);
    tc_fence_after();
    if (kGroups == 0) {
        tc_fence_before();
        __syncwarp();
        if ((r & 31) == 0) mbar_arrive_cluster(empty_bar_addr);
        return;
    }
#pragma unroll 1
    for (int g = 0; g < kGroups; ++g) {
        uint32_t acc[2][32];
        tmem_ld_32x32b_x32(taddr + g * 64, acc[0]);
        tmem_ld_32x32b_x32(taddr + g * 64 + 32, acc[1]);
        if (issuer) tma_store_wait_read<0>();                        // the previous store no longer reads the staging tile
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        tmem_ld_wait();
        if (g == kGroups - 1) {                                      // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if ((r & 31) == 0) mbar_arrive_cluster(empty_bar_addr);
        }
        uint8_t* srow = stage_tile + r * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float* bs = bias_s + half * kCols + g * 64 + h * 32;
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 bb = *reinterpret_cast<const float4*>(bs + 4 * j);
                v[4 * j + 0] = __uint_as_float(acc[h][4 * j + 0]) + bb.x;
                v[4 * j + 1] = __uint_as_float(acc[h][4 * j + 1]) + bb.y;
                v[4 * j + 2] = __uint_as_float(acc[h][4 * j + 2]) + bb.z;
                v[4 * j + 3] = __uint_as_float(acc[h][4 * j + 3]) + bb.w;
            }
            if constexpr (ACT == ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
            } else if constexpr (ACT == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 w;
                w.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                w.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                w.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                w.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                const int c16 = h * 4 + j;                           // 16-byte chunk index inside the 128-byte row
                *reinterpret_cast<uint4*>(srow + ((c16 ^ (r & 7)) << 4)) = w;
            }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (issuer) {
            if constexpr (kConvOut) tma_store_3d(tmap_out, stage_tile, col_base + g * 64, row0, b);
            else tma_store_2d(tmap_out, stage_tile, col_base + g * 64, row0);
            tma_store_commit();
        }
    }
}

// A_POS4 epilogue: accumulator row r = (utterance bb, slot i), column (j, n): out[b, 4i + j, 64 g + n] = x + gelu(acc + bias)
// (wav2vec2.py:915-917).  128 columns per warp = shifts j = 2 half, 2 half + 1.
__device__ __forceinline__ void epilogue_tile_pos4(const DevParams& p, uint32_t taddr, int half, int r, int g, int m_blk,
                                                   uint64_t* full_bar, uint32_t full_parity) {
    const int pair = m_blk / p.pos_sb, sb = m_blk - pair * p.pos_sb;
    const int bb = r >> 6, i = sb * 64 + (r & 63);
    const int b = 2 * pair + bb;
    mbar_wait(full_bar, full_parity);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        const int j = half * 2 + (c >> 1), n0 = (c & 1) * 32;
        const int t = 4 * i + j;
        const bool ok = b < p.pos_B && t < p.pos_T && !(p.debug_flags & 1);
        const long long off = ((long long)b * p.pos_T + t) * p.ldc + g * 64 + n0;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(taddr + c * 32, acc);
        float4 rv[8];
        if (ok) {
            const float4* r4 = reinterpret_cast<const float4*>(p.residual + off);
#pragma unroll
            for (int q = 0; q < 8; ++q) rv[q] = __ldg(r4 + q);
        }
        tmem_ld_wait();
        if (!ok) continue;
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + g * 64 + n0);
        float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 bb4 = __ldg(b4 + q);
            float4 o;
            o.x = rv[q].x + gelu_fast(__uint_as_float(acc[4 * q + 0]) + bb4.x);
            o.y = rv[q].y + gelu_fast(__uint_as_float(acc[4 * q + 1]) + bb4.y);
            o.z = rv[q].z + gelu_fast(__uint_as_float(acc[4 * q + 2]) + bb4.z);
            o.w = rv[q].w + gelu_fast(__uint_as_float(acc[4 * q + 3]) + bb4.w);
            o4[q] = o;
        }
    }
}

// In-place residual stream: out[tile] += acc + bias with cp.reduce.async.bulk.tensor (fp32 add performed at the L2): store-only
// epilogue, 32 fp32 columns (one 16 KB SWIZZLE_128B staging tile) at a time.  Used when the caller passes residual == out
// (x += branch(x), wav2vec2.py:1053 / :1058 with the stream updated in place).
template <int kCols>
__device__ __forceinline__ void epilogue_tile_red_tma(const CUtensorMap* tmap_out, uint32_t taddr, uint8_t* stage_tile, const float* bias_s,
                                                      int half, int r, int col_base, int row0, uint64_t* full_bar, uint32_t full_parity,
                                                      uint32_t empty_bar_addr, int dbg) {
    constexpr int kChunks = kCols / 32;
    const int bar_id = 1 + half;
    const bool issuer = r == 0;
    mbar_wait(full_bar, full_parity);
    tc_fence_after();
    uint8_t* srow = stage_tile + r * 128;
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(taddr + c * 32, acc);
        if (issuer) tma_store_wait_read<0>();                        // the previous reduce-store no longer reads the staging tile
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // (also publishes this tile's bias slice on the first pass)
        tmem_ld_wait();
        if (c == kChunks - 1) {                                      // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if ((r & 31) == 0) mbar_arrive_cluster(empty_bar_addr);
        }
        const float* bs = bias_s + half * kCols + c * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 bb = *reinterpret_cast<const float4*>(bs + 4 * j);
            *reinterpret_cast<float4*>(srow + ((j ^ (r & 7)) << 4)) =
                make_float4(__uint_as_float(acc[4 * j + 0]) + bb.x, __uint_as_float(acc[4 * j + 1]) + bb.y,
                            __uint_as_float(acc[4 * j + 2]) + bb.z, __uint_as_float(acc[4 * j + 3]) + bb.w);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (issuer && !(dbg & 4)) {
            tma_reduce_add_2d(tmap_out, stage_tile, col_base + c * 32, row0);
            tma_store_commit();
        }
    }
}

// The same reduce-add epilogue with the 16 KB staging tile used as TWO 8 KB buffers of 16 fp32 columns (64-byte rows, SWIZZLE_64B)
// in ping-pong: while the reduce-store of chunk c - 1 is still reading its buffer, the warps already convert chunk c into the
// other one; the issuer confirms "store c - 1 has read its buffer" right before the barrier that publishes chunk c, so at most
// one store is in flight per half and every buffer is provably free when it is written again.  One named barrier per chunk
// instead of two, and no thread ever waits for a store that was issued in the same breath (the 32-column version above waits
// for the full smem read of the store it has just issued before it may touch the tile again: K = 1024 tiles - out_proj - were
// epilogue-bound at ~2x their mainloop).  SLSB_RED_ADD_V1=1 selects the 32-column version (A/B).
template <int kCols>
__device__ __forceinline__ void epilogue_tile_red_tma16(const CUtensorMap* tmap_out16, uint32_t taddr, uint8_t* stage_tile, const float* bias_s,
                                                        int half, int r, int col_base, int row0, uint64_t* full_bar, uint32_t full_parity,
                                                        uint32_t empty_bar_addr, int dbg) {
    constexpr int kChunks = kCols / 16;
    const int bar_id = 1 + half;
    const bool issuer = r == 0;
    const int sw = (r >> 1) & 3;
    if (issuer) tma_store_wait_read<0>();                            // stores of the previous tile (other epilogue flavours included)
    mbar_wait(full_bar, full_parity);
    tc_fence_after();
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // this tile's bias slice is visible; both buffers are free
    uint32_t acc[2][16];
    tmem_ld_32x32b_x16(taddr, acc[0]);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        tmem_ld_wait();
        if (c + 1 < kChunks) tmem_ld_32x32b_x16(taddr + (c + 1) * 16, acc[(c + 1) & 1]);
        uint8_t* srow = stage_tile + (c & 1) * 8192 + r * 64;
        const float* bs = bias_s + half * kCols + c * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 bb = *reinterpret_cast<const float4*>(bs + 4 * j);
            *reinterpret_cast<float4*>(srow + ((j ^ sw) << 4)) =
                make_float4(__uint_as_float(acc[c & 1][4 * j + 0]) + bb.x, __uint_as_float(acc[c & 1][4 * j + 1]) + bb.y,
                            __uint_as_float(acc[c & 1][4 * j + 2]) + bb.z, __uint_as_float(acc[c & 1][4 * j + 3]) + bb.w);
        }
        fence_proxy_async_smem();
        if (c == kChunks - 1) {                                      // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if ((r & 31) == 0) mbar_arrive_cluster(empty_bar_addr);
        }
        if (issuer && c > 0) tma_store_wait_read<0>();               // store c - 1 has read its buffer: chunk c + 1 may overwrite it
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (issuer && !(dbg & 4)) {
            tma_reduce_add_2d(tmap_out16, stage_tile + (c & 1) * 8192, col_base + c * 16, row0);
            tma_store_commit();
        }
    }
}

// fp32 residual stream through the GEMM epilogue without touching the LSU's global path: per 32-column chunk the issuer
// TMA-loads the residual tile [128 rows x 32 fp32] into the half's staging tile, every thread adds its accumulator row + bias
// IN PLACE (own row, SWIZZLE_128B chunk positions), and the same tile is TMA-stored as the new residual stream.
// (out = x + branch(x), wav2vec2.py:1053 / :1058, with x kept in fp32.)
template <int kCols>
__device__ __forceinline__ void epilogue_tile_res_tma(const CUtensorMap* tmap_res, const CUtensorMap* tmap_out, uint32_t taddr, uint8_t* stage_tile,
                                                      const float* bias_s, int half, int r, int col_base, int row0,
                                                      uint64_t* full_bar, uint32_t full_parity, uint32_t empty_bar_addr,
                                                      uint64_t* res_bar, uint32_t& res_phase, int dbg = 0) {
    // dbg (SLSB_DEBUG_FLAGS, measurements only): bit1 = no residual TMA loads (residual reads as whatever the tile holds),
    // bit2 = no TMA stores
    constexpr int kChunks = kCols / 32;
    const int bar_id = 1 + half;
    const bool issuer = r == 0;
    const bool do_load = !(dbg & 2), do_store = !(dbg & 4);
    if (issuer && do_load) {                                         // residual chunk 0 travels while the MMAs of this tile still run
        tma_store_wait_read<0>();
        mbar_expect_tx(res_bar, 128 * 128);
        tma_load_2d(stage_tile, tmap_res, res_bar, col_base, row0);
    }
    mbar_wait(full_bar, full_parity);
    tc_fence_after();
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // this tile's bias slice (written by the caller) is visible
    uint8_t* srow = stage_tile + r * 128;
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(taddr + c * 32, acc);
        if (do_load) {
            mbar_wait(res_bar, res_phase);
            res_phase ^= 1u;
        }
        tmem_ld_wait();
        if (c == kChunks - 1) {                                      // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if ((r & 31) == 0) mbar_arrive_cluster(empty_bar_addr);
        }
        const float* bs = bias_s + half * kCols + c * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4* slot = reinterpret_cast<float4*>(srow + ((j ^ (r & 7)) << 4));
            const float4 rv = *slot;
            const float4 bb = *reinterpret_cast<const float4*>(bs + 4 * j);
            float4 o;
            o.x = __uint_as_float(acc[4 * j + 0]) + bb.x + rv.x; o.y = __uint_as_float(acc[4 * j + 1]) + bb.y + rv.y;
            o.z = __uint_as_float(acc[4 * j + 2]) + bb.z + rv.z; o.w = __uint_as_float(acc[4 * j + 3]) + bb.w + rv.w;
            *slot = o;
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (issuer) {
            if (do_store) {
                tma_store_2d(tmap_out, stage_tile, col_base + c * 32, row0);
                tma_store_commit();
            }
            if (c + 1 < kChunks && do_load) {
                tma_store_wait_read<0>();                             // the store has read the tile: it can take the next residual chunk
                mbar_expect_tx(res_bar, 128 * 128);
                tma_load_2d(stage_tile, tmap_res, res_bar, col_base + (c + 1) * 32, row0);
            }
        }
    }
}

template <int BLOCK_N, int A_MODE>
__global__ void __launch_bounds__(kNumThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res, const DevParams p) {
    using Plan = SmemPlan<BLOCK_N>;
    constexpr int kStages = Plan::kStages;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;   // two accumulator stages (power of two >= 32 for BLOCK_N in {64,128,256})

    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("slsb: dynamic smem base not 1024-aligned\n"); __trap(); }
    float* bias_s = reinterpret_cast<float*>(smem + Plan::kBiasOffset);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Plan::kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* res_bar = tmem_empty + 2;            // [2] residual chunk of column half h has landed in its staging tile
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.batches * p.m_tiles * p.n_tiles;
    const int num_kb = p.K / BLOCK_K;

    griddep_launch();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (p.tma_store || p.res_tma || p.red_add) tma_prefetch_desc(&tmap_out);
        if (p.res_tma) tma_prefetch_desc(&tmap_res);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kNumEpiWarps); mbar_init(&res_bar[s], 1); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    griddep_wait();                     // the prologue above overlapped the previous kernel's tail; its outputs are visible from here

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int t = tile;
                const int n_blk = t % p.n_tiles; t /= p.n_tiles;      // n fastest: the CTAs running together share one band of A
                const int m_blk = t % p.m_tiles;
                const int b = t / p.m_tiles;
                int kb_lo = 0, kb_hi = num_kb;
                if constexpr (A_MODE == A_PLAIN) {
                    if (p.kb_per_split > 0) { kb_lo = b * p.kb_per_split; kb_hi = min(num_kb, kb_lo + p.kb_per_split); }
                }
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Plan::kStage;
                    uint8_t* sb = sa + Plan::kStageA;
                    mbar_expect_tx(&full_bar[stage], Plan::kStage);
                    if constexpr (A_MODE == A_PLAIN) {
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
                    } else if constexpr (A_MODE == A_CONV) {
                        const int k0 = kb * BLOCK_K;
                        const int tap = k0 / p.conv_cin, c = k0 - tap * p.conv_cin;
                        tma_load_4d(sa, &tmap_a, &full_bar[stage], c, tap % p.conv_stride, m_blk * BLOCK_M + tap / p.conv_stride, b);
                    } else if constexpr (A_MODE == A_POS) {
                        tma_load_3d(sa, &tmap_a, &full_bar[stage], n_blk * BLOCK_K, m_blk * BLOCK_M + kb, b);
                    } else {       // A_POS4: input frame 4i + kb of utterances 2 m_blk, 2 m_blk + 1, channels of group n_blk
                        const int pair = m_blk / p.pos_sb, sb = m_blk - pair * p.pos_sb;     // 64 slots (256 output frames) per tile
                        tma_load_4d(sa, &tmap_a, &full_bar[stage], n_blk * 64, kb & 3, sb * 64 + (kb >> 2), 2 * pair);
                    }
                    if constexpr (A_MODE == A_POS4) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) tma_load_3d(sb + j * 8192, &tmap_b, &full_bar[stage], 0, kb - j, n_blk * 64);
                    } else {
                        tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                int nkb = num_kb;
                if constexpr (A_MODE == A_PLAIN) {
                    if (p.kb_per_split > 0) {
                        const int kb_lo = (tile / (p.m_tiles * p.n_tiles)) * p.kb_per_split;
                        nkb = min(num_kb, kb_lo + p.kb_per_split) - kb_lo;
                    }
                }
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Plan::kStage);
                    const uint32_t sb = sa + Plan::kStageA;
                    const uint64_t da = make_smem_desc_sw128(sa, 0, 1024);
                    const uint64_t db = make_smem_desc_sw128(sb, 0, 1024);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance both descriptors by k * 16 elements * 2 B = 32 B inside the 128-B swizzle row
                        tc_mma_f16(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(&empty_bar[stage]);          // smem slot free once these MMAs have read it
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tmem_full[acc]);                // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int q = warp & 3;                    // TMEM lane quadrant this warp may access
        const int half = (warp - 4) >> 2;          // column half
        constexpr int kColsPerWarp = BLOCK_N / 2;
        const int r = q * 32 + lane;
        uint32_t res_phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            int t = tile;
            const int n_blk = t % p.n_tiles; t /= p.n_tiles;
            const int m_blk = t % p.m_tiles;
            const int b = t / p.m_tiles;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int row = m_blk * BLOCK_M + r;
            const bool row_ok = row < p.M;
            const long long out_off = (long long)b * p.out_batch_stride + (long long)row * p.ldc;
            const long long res_off = (long long)b * p.res_batch_stride + (long long)row * p.ldr;
            const uint32_t taddr0 = tmem_base + (uint32_t(q * 32) << 16) + acc * BLOCK_N + half * kColsPerWarp;
            const int col_base = n_blk * BLOCK_N + half * kColsPerWarp;
            if constexpr (A_MODE == A_POS4) {
                epilogue_tile_pos4(p, taddr0, half, r, n_blk, m_blk, &tmem_full[acc], acc_phase);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                continue;
            }
            if constexpr (BLOCK_N == 256 && A_MODE == A_PLAIN) {
                if (p.red_add) {
                    bias_s[half * kColsPerWarp + r] = __ldg(p.bias + col_base + r);       // visible after the first named barrier
                    uint8_t* stage_tile = smem + Plan::kStoreOffset + half * 16384;
                    epilogue_tile_red_tma<kColsPerWarp>(&tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, m_blk * BLOCK_M,
                                                        &tmem_full[acc], acc_phase, smem_u32(&tmem_empty[acc]), p.debug_flags);
                    continue;
                }
                if (p.res_tma) {
                    bias_s[half * kColsPerWarp + r] = __ldg(p.bias + col_base + r);       // visible after the first named barrier
                    uint8_t* stage_tile = smem + Plan::kStoreOffset + half * 16384;
                    epilogue_tile_res_tma<kColsPerWarp>(&tmap_res, &tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, m_blk * BLOCK_M,
                                                        &tmem_full[acc], acc_phase, smem_u32(&tmem_empty[acc]), &res_bar[half], res_phase);
                    continue;
                }
            }
            if constexpr (BLOCK_N >= 128) {
                if (p.tma_store) {
                    // this half's 128 bias values -> smem (visible after the first named barrier of the tile)
                    bias_s[half * kColsPerWarp + r] = __ldg(p.bias + col_base + r);
                    uint8_t* stage_tile = smem + Plan::kStoreOffset + half * 16384;
                    constexpr bool kConvOut = (A_MODE == A_CONV);
                    if (p.act == ACT_GELU)
                        epilogue_tile_tma<ACT_GELU, kColsPerWarp, kConvOut>(p, &tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, m_blk * BLOCK_M, b, &tmem_full[acc], acc_phase, smem_u32(&tmem_empty[acc]));
                    else if (p.act == ACT_RELU)
                        epilogue_tile_tma<ACT_RELU, kColsPerWarp, kConvOut>(p, &tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, m_blk * BLOCK_M, b, &tmem_full[acc], acc_phase, smem_u32(&tmem_empty[acc]));
                    else
                        epilogue_tile_tma<ACT_NONE, kColsPerWarp, kConvOut>(p, &tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, m_blk * BLOCK_M, b, &tmem_full[acc], acc_phase, smem_u32(&tmem_empty[acc]));
                    continue;
                }
            }
            const int sel = p.act * 4 + p.out_bf16 * 2 + (p.residual != nullptr ? 1 : 0);
#define SLSB_EPI(A_, O_, R_) epilogue_tile<A_, O_, R_, kColsPerWarp>(p, taddr0, out_off, res_off, col_base, row_ok, &tmem_full[acc], acc_phase)
            switch (sel) {
                case ACT_NONE * 4 + 2 + 0: SLSB_EPI(ACT_NONE, true, false); break;
                case ACT_GELU * 4 + 2 + 0: SLSB_EPI(ACT_GELU, true, false); break;
                case ACT_NONE * 4 + 0 + 1: SLSB_EPI(ACT_NONE, false, true); break;
                case ACT_NONE * 4 + 0 + 0: SLSB_EPI(ACT_NONE, false, false); break;
                case ACT_RELU * 4 + 0 + 0: SLSB_EPI(ACT_RELU, false, false); break;
                case ACT_GELU * 4 + 0 + 1: SLSB_EPI(ACT_GELU, false, true); break;
                case ACT_GELU * 4 + 0 + 0: SLSB_EPI(ACT_GELU, false, false); break;
                case ACT_RELU * 4 + 2 + 0: SLSB_EPI(ACT_RELU, true, false); break;
                default: mbar_wait(&tmem_full[acc], acc_phase); break;   // launcher rejects other combinations
            }
#undef SLSB_EPI
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if ((p.tma_store || p.res_tma || p.red_add) && r == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// =================================================================================================
// CTA-pair version for the big plain GEMMs (A_PLAIN, 256-wide tiles): tcgen05.mma.cta_group::2
//
// A cluster of two CTAs (one TPC) computes a 256 x 256 tile: CTA r holds rows [128 r, 128 r + 128) of A and of the accumulator
// and rows [128 r, 128 r + 128) of the 256 W rows (the N dimension): the pair's tensor cores read BOTH CTAs' shared memory, so
// each SM stages 32 KB instead of 48 KB per 64-deep k-block (2/3 of the L2 -> SM operand traffic per FLOP) and 6 pipeline stages
// fit where the single-CTA kernel has 4.  Only the leader (rank 0) issues MMAs; both CTAs' TMA loads complete on the leader's
// `full` barrier; tcgen05.commit multicasts the `empty` / `tmem_full` arrivals to both CTAs; the epilogue warps of both CTAs
// hand the accumulator back through the leader's `tmem_empty` barrier.  Epilogues are the ones of the single-CTA kernel.
// =================================================================================================
struct Plan2 {
    static constexpr int kStageA = BLOCK_M * BLOCK_K * 2;            // this CTA's 128 rows of A
    static constexpr int kStageB = 128 * BLOCK_K * 2;                // this CTA's 128 of the tile's 256 W rows
    static constexpr int kStage = kStageA + kStageB;                 // 32 KB
    static constexpr int kStages = 6;
    static constexpr int kStoreOffset = kStages * kStage;
    static constexpr int kStoreBytes = 2 * 16384;
    static constexpr int kBiasOffset = kStoreOffset + kStoreBytes;
    static constexpr int kBarOffset = kBiasOffset + 256 * 4;
    static constexpr int kBytes = kBarOffset + 256;
};

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_cluster(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// TMA load whose completion is signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// MMA completion -> arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// Static schedule of the pair kernel.  Tiles are ordered n-fastest with the ragged last row block (M % 256 rows, cheaper) at the
// end.  The floor(tiles / P) full rounds hand tile i * P + p to pair p in even rounds and to pair P - 1 - p in odd rounds (snake
// order).  The tiles left for a partial last round are cut into `split` column slices of 256 / split columns (one MMA of that N,
// same K order, so every output element is computed exactly as in a whole tile) and dealt round-robin, so the last round costs
// ceil(rem * split / P) / split of a round instead of a whole one (qkv, M = 12864: 9 -> 8.5 rounds; fc1: 12 -> 11.25).
struct PairItem { int m2, n0, width; };
__host__ __device__ __forceinline__ bool pair_item(int it, int pair, int num_pairs, int num_tiles, int n_tiles, int split, PairItem& w) {
    const int full_rounds = num_tiles / num_pairs;
    int tile, sub = 0;
    w.width = 256;
    if (it < full_rounds) {
        tile = it * num_pairs + ((it & 1) ? (num_pairs - 1 - pair) : pair);
    } else {
        const int j = (it - full_rounds) * num_pairs + pair;
        if (j >= (num_tiles - full_rounds * num_pairs) * split) return false;
        tile = full_rounds * num_pairs + j / split;
        sub = j % split;
        w.width = 256 / split;
    }
    w.m2 = tile / n_tiles;
    w.n0 = (tile % n_tiles) * 256 + sub * w.width;
    return true;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res, const DevParams p) {
    constexpr int kStages = Plan2::kStages;
    constexpr int BLOCK_N = 256;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("slsb: dynamic smem base not 1024-aligned\n"); __trap(); }
    float* bias_s = reinterpret_cast<float*>(smem + Plan2::kBiasOffset);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Plan2::kBarOffset);      // used in the leader CTA only
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;                                            // used in the leader CTA only (16 warp arrivals)
    uint64_t* res_bar = tmem_empty + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int m2_tiles = (p.M + 255) / 256;
    const int num_tiles = m2_tiles * p.n_tiles;
    const int num_kb = p.K / BLOCK_K;
    const int split = p.tail_split;
    // events: 0 accumulator free (MMA), 1 first k-block landed, 2 last MMA issued, 3 epilogue enters, 4 accumulator complete seen by
    // the epilogue, 5 epilogue done, 6 first load of the tile issued, 7 last load issued
#define GEMM_TRACE(it_, ev_) do { if (p.trace != nullptr && pair == 0 && rank == 0 && (it_) < 64) p.trace[(it_) * 8 + (ev_)] = clock64(); } while (0)

    griddep_launch();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (p.tma_store || p.res_tma || p.red_add) tma_prefetch_desc(&tmap_out);
        if (p.res_tma) tma_prefetch_desc(&tmap_res);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 2 * kNumEpiWarps); mbar_init(&res_bar[s], 1); }
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_barrier();                  // both CTAs' barriers and TMEM are ready before any cross-CTA traffic
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    griddep_wait();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; completion on the leader's barrier) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            PairItem w;
            for (int it = 0; pair_item(it, pair, num_pairs, num_tiles, p.n_tiles, split, w); ++it) {
                const int row0 = w.m2 * 256 + (int)rank * 128;
                const int wrow0 = w.n0 + (int)rank * (w.width >> 1);       // this CTA's half of the slice's W rows (box: 128 rows)
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (kb == 0) GEMM_TRACE(it, 6);
                    if (kb == num_kb - 1) GEMM_TRACE(it, 7);
                    uint8_t* sa = smem + stage * Plan2::kStage;
                    uint8_t* sb = sa + Plan2::kStageA;
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Plan2::kStage);
                    const uint32_t bar = map_cluster(smem_u32(&full_bar[stage]), 0);
                    tma_load_2d_pair(sa, &tmap_a, bar, kb * BLOCK_K, row0);
                    tma_load_2d_pair(sb, &tmap_b, bar, kb * BLOCK_K, wrow0);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread of the leader CTA) =====================
        if (lane == 0 && rank == 0) {
            int stage = 0; uint32_t phase = 0;
            PairItem w;
            for (int it = 0; pair_item(it, pair, num_pairs, num_tiles, p.n_tiles, split, w); ++it) {
                const uint32_t idesc = w.width == 256 ? make_idesc_bf16(256, 256) : w.width == 128 ? make_idesc_bf16(256, 128) : make_idesc_bf16(256, 64);
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                GEMM_TRACE(it, 0);
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (kb == 0) GEMM_TRACE(it, 1);
                    const uint32_t sa = smem_u32(smem + stage * Plan2::kStage);
                    const uint64_t da = make_smem_desc_sw128(sa, 0, 1024);
                    const uint64_t db = make_smem_desc_sw128(sa + Plan2::kStageA, 0, 1024);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        tc_mma_f16_pair(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit_pair(&empty_bar[stage]);     // both CTAs may refill this stage
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit_pair(&tmem_full[acc]);           // both CTAs' epilogues may read their 128 rows
                GEMM_TRACE(it, 2);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps (both CTAs, own 128 rows) =====================
        const int q = warp & 3, half = (warp - 4) >> 2;
        constexpr int kColsPerWarp = BLOCK_N / 2;
        const int r = q * 32 + lane;
        uint32_t res_phase = 0;
        PairItem w;
        for (int it = 0; pair_item(it, pair, num_pairs, num_tiles, p.n_tiles, split, w); ++it) {
            const int row0 = w.m2 * 256 + (int)rank * 128;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            // this half-warpgroup's columns of the slice: 2 / 1 groups of 64 columns; a 64-column slice belongs to half 0 alone
            const int groups = w.width == 256 ? 2 : (w.width == 128 || half == 0) ? 1 : 0;
            const int col_off = w.width == 64 ? 0 : half * (w.width >> 1);
            const uint32_t taddr0 = tmem_base + (uint32_t(q * 32) << 16) + acc * BLOCK_N + col_off;
            const int col_base = w.n0 + col_off;
            const uint32_t free_bar = map_cluster(smem_u32(&tmem_empty[acc]), 0);
            if (r < groups * 64) bias_s[half * kColsPerWarp + r] = __ldg(p.bias + col_base + r);   // visible after the first named barrier of the tile
            uint8_t* stage_tile = smem + Plan2::kStoreOffset + half * 16384;
            if (p.trace != nullptr && warp == 4 && lane == 0) {
                GEMM_TRACE(it, 3);
                mbar_wait(&tmem_full[acc], acc_phase);
                GEMM_TRACE(it, 4);
            }
            if (p.red_add == 2) {
                epilogue_tile_red_tma16<kColsPerWarp>(&tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, row0, &tmem_full[acc], acc_phase, free_bar,
                                                      p.debug_flags);
            } else if (p.red_add) {
                epilogue_tile_red_tma<kColsPerWarp>(&tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, row0, &tmem_full[acc], acc_phase, free_bar,
                                                    p.debug_flags);
            } else if (p.res_tma) {
                epilogue_tile_res_tma<kColsPerWarp>(&tmap_res, &tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, row0,
                                                    &tmem_full[acc], acc_phase, free_bar, &res_bar[half], res_phase, p.debug_flags);
            } else if (p.act == ACT_GELU) {
                epilogue_tile_tma<ACT_GELU, kColsPerWarp, false>(p, &tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, row0, 0, &tmem_full[acc], acc_phase, free_bar, groups);
            } else if (p.act == ACT_RELU) {
                epilogue_tile_tma<ACT_RELU, kColsPerWarp, false>(p, &tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, row0, 0, &tmem_full[acc], acc_phase, free_bar, groups);
            } else {
                epilogue_tile_tma<ACT_NONE, kColsPerWarp, false>(p, &tmap_out, taddr0, stage_tile, bias_s, half, r, col_base, row0, 0, &tmem_full[acc], acc_phase, free_bar, groups);
            }
            if (warp == 4 && lane == 0) GEMM_TRACE(it, 5);
        }
        if (r == 0) tma_store_wait<0>();
    }
#undef GEMM_TRACE
    tc_fence_before();
    __syncthreads();
    cluster_barrier();                  // nobody frees TMEM / exits while the peer's tensor core may still touch this CTA
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// column slices per tile of the partial last round: the s in {1, 2, 4} with the smallest cost ceil(rem * s / P) / s (ties -> smaller s)
int pair_tail_split(int tiles, int pairs) {
    const int rem = tiles % pairs;
    int best_num = 1, best_den = 1, split = 1;
    if (rem == 0) return 1;
    for (int s = 2; s <= 4; s *= 2) {
        const int num = (rem * s + pairs - 1) / pairs;
        if (num * best_den < best_num * s) { best_num = num; best_den = s; split = s; }
    }
    return split;
}

int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr, const DevParams& dp, int num_sms,
                cudaStream_t stream) {
    static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
    static int max_pairs = 0;
    if (first_use_on_device(&configured_on)) {
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan2::kBytes));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms & ~1); cfg.blockDim = dim3(kNumThreads); cfg.dynamicSmemBytes = Plan2::kBytes;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, tc_gemm_pair_kernel, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms / 2; }
        max_pairs = n < num_sms / 2 ? n : num_sms / 2;
    }
    const int tiles = ((dp.M + 255) / 256) * dp.n_tiles;
    const int pairs = tiles < max_pairs ? tiles : max_pairs;
    DevParams q = dp;
    q.tail_split = 1;
    static const bool tail_on = !(getenv("SLSB_NO_TAIL_SPLIT") && atoi(getenv("SLSB_NO_TAIL_SPLIT")) != 0);
    if (tail_on && dp.tma_store && !dp.red_add && !dp.res_tma) q.tail_split = pair_tail_split(tiles, pairs);   // bf16 TMA-store epilogue handles column slices
    static const bool trace_on = getenv("SLSB_GEMM_TRACE") && atoi(getenv("SLSB_GEMM_TRACE")) != 0;
    if (trace_on) {   // tuning aid: timeline of pair 0, printed after a stream sync (never on in production)
        long long* tr_dev = nullptr;                     // plain device memory: a managed buffer would page-fault inside the kernel
        static long long tr_buf[64 * 8];
        SLSB_CUDA_CHECK(cudaMalloc(&tr_dev, sizeof(tr_buf)));
        SLSB_CUDA_CHECK(cudaMemsetAsync(tr_dev, 0, sizeof(tr_buf), stream));
        q.trace = tr_dev;
        SLSB_CUDA_CHECK(launch_pdl(tc_gemm_pair_kernel, dim3(2 * pairs), dim3(kNumThreads), Plan2::kBytes, stream, ta, tb, to, tr, q));
        SLSB_CUDA_CHECK(cudaMemcpyAsync(tr_buf, tr_dev, sizeof(tr_buf), cudaMemcpyDeviceToHost, stream));
        SLSB_CUDA_CHECK(cudaStreamSynchronize(stream));
        long long t0 = 0;
        for (int i = 0; i < 64 * 8; ++i) if (tr_buf[i] && (!t0 || tr_buf[i] < t0)) t0 = tr_buf[i];
        fprintf(stderr, "gemm_trace M=%d N=%d K=%d pairs=%d tiles=%d split=%d red_add=%d tma_store=%d act=%d\n", dp.M, dp.N, dp.K, pairs, tiles, q.tail_split,
                dp.red_add, dp.tma_store, dp.act);
        for (int it = 0; it < 64; ++it) {
            if (!tr_buf[it * 8 + 0]) break;
            fprintf(stderr, "  it %2d: acc_free %7lld first_kb %7lld mma_done_issue %7lld | epi_enter %7lld acc_ready %7lld epi_done %7lld | load_first %7lld load_last %7lld\n", it,
                    tr_buf[it * 8 + 0] - t0, tr_buf[it * 8 + 1] - t0, tr_buf[it * 8 + 2] - t0, tr_buf[it * 8 + 3] - t0, tr_buf[it * 8 + 4] - t0,
                    tr_buf[it * 8 + 5] - t0, tr_buf[it * 8 + 6] - t0, tr_buf[it * 8 + 7] - t0);
        }
        cudaFree(tr_dev);
        return 0;
    }
    SLSB_CUDA_CHECK(launch_pdl(tc_gemm_pair_kernel, dim3(2 * pairs), dim3(kNumThreads), Plan2::kBytes, stream, ta, tb, to, tr, q));
    return 0;
}

bool epilogue_supported(int act, int out_bf16, bool has_res) {
    const int sel = act * 4 + out_bf16 * 2 + (has_res ? 1 : 0);
    switch (sel) {
        case ACT_NONE * 4 + 2 + 0: case ACT_GELU * 4 + 2 + 0: case ACT_NONE * 4 + 0 + 1: case ACT_NONE * 4 + 0 + 0:
        case ACT_RELU * 4 + 0 + 0: case ACT_GELU * 4 + 0 + 1: case ACT_GELU * 4 + 0 + 0: case ACT_RELU * 4 + 2 + 0:
            return true;
        default: return false;
    }
}

template <int BLOCK_N, int A_MODE>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr, const DevParams& dp, int num_sms, cudaStream_t stream) {
    using Plan = SmemPlan<BLOCK_N>;
    static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
    auto kern = tc_gemm_kernel<BLOCK_N, A_MODE>;
    if (first_use_on_device(&configured_on)) {
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::kBytes));
    }
    const int tiles = dp.batches * dp.m_tiles * dp.n_tiles;
    const int grid = tiles < num_sms ? tiles : num_sms;
    SLSB_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(kNumThreads), Plan::kBytes, stream, ta, tb, to, tr, dp));
    return 0;
}

}  // namespace

// Host-side replay of the pair kernel's static schedule (same pair_item() the device roles call): rows of
// (pair, round, first row, first column, columns).  Test hook: every output column block must be produced exactly once.
int pair_schedule(int M, int N, int num_pairs, int* items, int max_items, int* split_out) {
    if (M < 1 || N < 256 || N % 256 != 0 || num_pairs < 1) { set_error("pair_schedule: M >= 1, N %% 256 == 0, num_pairs >= 1"); return -1; }
    const int n_tiles = N / 256, tiles = ((M + 255) / 256) * n_tiles;
    const int pairs = tiles < num_pairs ? tiles : num_pairs;
    const int split = pair_tail_split(tiles, pairs);
    if (split_out) *split_out = split;
    int n = 0;
    for (int p = 0; p < pairs; ++p) {
        PairItem w;
        for (int it = 0; pair_item(it, p, pairs, tiles, n_tiles, split, w); ++it) {
            if (n < max_items) { int* r = items + 5 * n; r[0] = p; r[1] = it; r[2] = w.m2 * 256; r[3] = w.n0; r[4] = w.width; }
            ++n;
        }
    }
    return n;
}

int tc_gemm(const TcGemmArgs& g, int num_sms, cudaStream_t stream) {
    if (g.K % BLOCK_K != 0 || g.K <= 0) { set_error("tc_gemm: K=%d must be a positive multiple of %d", g.K, BLOCK_K); return -1; }
    if (!epilogue_supported(g.act, g.out_bf16, g.residual != nullptr)) {
        set_error("tc_gemm: unsupported epilogue act=%d out_bf16=%d res=%d", g.act, g.out_bf16, g.residual != nullptr);
        return -1;
    }
    int block_n;
    if (g.a_mode == A_POS4) {
        if (g.act != ACT_GELU || g.out_bf16 || g.residual == nullptr || g.pos_tp % 4 != 0 || g.N % 64 != 0 || g.K % 64 != 0) {
            set_error("tc_gemm: A_POS4 needs fp32 out + residual + GELU, padded length %% 4 == 0 (got %d), N %% 64 == 0", g.pos_tp);
            return -1;
        }
        block_n = 256;
    } else if (g.a_mode == A_POS) block_n = 64;
    else if (g.N % 256 == 0) block_n = 256;
    else if (g.N % 128 == 0) block_n = 128;
    else if (g.N % 64 == 0) block_n = 64;
    else { set_error("tc_gemm: N=%d must be a multiple of 64", g.N); return -1; }
    if (g.M <= 0 || g.batches <= 0) return 0;

    DevParams dp{};
    dp.M = g.M; dp.N = g.N; dp.K = g.K;
    dp.batches = g.batches;
    dp.m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
    dp.n_tiles = g.N / block_n;
    dp.conv_cin = g.conv_cin; dp.conv_stride = g.conv_stride;
    dp.out = g.out; dp.ldc = g.ldc; dp.out_batch_stride = g.out_batch_stride;
    dp.bias = g.bias; dp.residual = g.residual; dp.ldr = g.ldr; dp.res_batch_stride = g.res_batch_stride;
    dp.act = g.act; dp.out_bf16 = g.out_bf16;
    if (g.a_mode == A_POS4) {       // tiles: (utterance pair, group); k-blocks: taps + 3 input frames
        dp.pos_T = g.M; dp.pos_B = g.batches;
        dp.pos_sb = (g.M + 255) / 256;                       // utterances longer than 256 frames take several slot blocks
        dp.batches = 1; dp.m_tiles = ((g.batches + 1) / 2) * dp.pos_sb; dp.n_tiles = g.N / 64;
        dp.K = g.K + 3 * BLOCK_K;
    }
    if (g.k_splits > 1) {
        if (g.a_mode != A_PLAIN || g.batches != 1) { set_error("tc_gemm: split-K needs A_PLAIN and batches == 1"); return -1; }
        const int total_kb = g.K / BLOCK_K;
        dp.kb_per_split = (total_kb + g.k_splits - 1) / g.k_splits;
        dp.batches = (total_kb + dp.kb_per_split - 1) / dp.kb_per_split;      // every split owns >= 1 k-block
        if (dp.batches != g.k_splits) { set_error("tc_gemm: k_splits=%d leaves empty splits for K=%d (use %d)", g.k_splits, g.K, dp.batches); return -1; }
    }
    { const char* dbg = getenv("SLSB_DEBUG_FLAGS"); dp.debug_flags = dbg ? atoi(dbg) : 0; }

    CUtensorMap ta, tb, to, tr;
    memset(&to, 0, sizeof(to));
    memset(&tr, 0, sizeof(tr));
    // fp32 output with an fp32 residual of full-width tiles: residual in / sum out through TMA (epilogue_tile_res_tma)
    dp.res_tma = (!g.out_bf16 && g.residual != nullptr && g.act == ACT_NONE && block_n == 256 && g.a_mode == A_PLAIN && g.k_splits <= 1 &&
                  g.ldc % 4 == 0 && g.ldr % 4 == 0 && !getenv("SLSB_NO_RES_TMA")) ? 1 : 0;
    // residual == out: the stream is advanced in place by reduce-add stores, no residual tile is loaded at all
    dp.red_add = (dp.res_tma && static_cast<const void*>(g.residual) == g.out && g.ldr == g.ldc && !getenv("SLSB_NO_RED_ADD")) ? 1 : 0;
    if (dp.red_add) dp.res_tma = 0;
    if (dp.red_add) {
        uint64_t dims[2] = {(uint64_t)g.N, (uint64_t)g.M};
        uint32_t box[2] = {32, BLOCK_M};
        uint64_t so[1] = {(uint64_t)g.ldc * 4};
        if (encode_tmap_f32(&to, g.out, 2, dims, so, box)) return -1;
    }
    if (dp.res_tma) {
        uint64_t dims[2] = {(uint64_t)g.N, (uint64_t)g.M};
        uint32_t box[2] = {32, BLOCK_M};
        uint64_t so[1] = {(uint64_t)g.ldc * 4}, sr[1] = {(uint64_t)g.ldr * 4};
        if (encode_tmap_f32(&to, g.out, 2, dims, so, box)) return -1;
        if (encode_tmap_f32(&tr, g.residual, 2, dims, sr, box)) return -1;
    }
    // bf16 outputs of full-width tiles leave through TMA stores (BLOCK_N = 256 -> two 128-column halves of two 64-column groups)
    dp.tma_store = (g.out_bf16 && g.residual == nullptr && block_n == 256 && g.a_mode != A_POS && g.a_mode != A_POS4 && g.k_splits <= 1 &&
                    !getenv("SLSB_NO_TMA_STORE")) ? 1 : 0;
    if (dp.tma_store) {
        if (g.a_mode == A_CONV) {
            uint64_t dims[3] = {(uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.batches};
            uint64_t strides[2] = {(uint64_t)g.ldc * 2, (uint64_t)g.out_batch_stride * 2};
            uint32_t box[3] = {64, BLOCK_M, 1};
            if (encode_tmap_bf16(&to, g.out, 3, dims, strides, box)) return -1;
        } else {
            uint64_t dims[2] = {(uint64_t)g.N, (uint64_t)g.M};
            uint64_t strides[1] = {(uint64_t)g.ldc * 2};
            uint32_t box[2] = {64, BLOCK_M};
            if (encode_tmap_bf16(&to, g.out, 2, dims, strides, box)) return -1;
        }
    }
    if (g.a_mode == A_POS4) {   // W [N][taps][64] viewed as (c, tap, n): one box = the 64 out-channels of a group at one tap
        const uint64_t taps = (uint64_t)g.K / 64;
        uint64_t dims[3] = {64, taps, (uint64_t)g.N};
        uint64_t strides[2] = {128, taps * 128};
        uint32_t box[3] = {64, 1, 64};
        if (encode_tmap_bf16(&tb, g.W, 3, dims, strides, box)) return -1;
    } else {   // W: [N, K] K-major
        uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.N};
        uint64_t strides[1] = {(uint64_t)g.ldw * 2};
        uint32_t box[2] = {BLOCK_K, (uint32_t)block_n};
        if (encode_tmap_bf16(&tb, g.W, 2, dims, strides, box)) return -1;
    }
    if (g.a_mode == A_PLAIN) {
        uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.M};
        uint64_t strides[1] = {(uint64_t)g.lda * 2};
        uint32_t box[2] = {BLOCK_K, BLOCK_M};
        if (encode_tmap_bf16(&ta, g.A, 2, dims, strides, box)) return -1;
    } else if (g.a_mode == A_CONV) {
        // activations [B, L_in, C] viewed as (c, l_in % s, l_in / s, b)
        const uint64_t C = g.conv_cin, s = g.conv_stride, Lin = g.conv_lin;
        uint64_t dims[4] = {C, s, (Lin + s - 1) / s, (uint64_t)g.batches};
        uint64_t strides[3] = {C * 2, s * C * 2, Lin * C * 2};
        uint32_t box[4] = {BLOCK_K, 1, BLOCK_M, 1};
        if (encode_tmap_bf16(&ta, g.A, 4, dims, strides, box)) return -1;
    } else if (g.a_mode == A_POS4) {
        // zero-padded stream [B, Tp, D] viewed as (c, t % 4, t / 4, b); box = 64 channels x 64 slots x 2 utterances
        const uint64_t Dd = g.pos_dim, Tp = g.pos_tp;
        uint64_t dims[4] = {Dd, 4, Tp / 4, (uint64_t)g.batches};
        uint64_t strides[3] = {Dd * 2, 4 * Dd * 2, Tp * Dd * 2};
        uint32_t box[4] = {BLOCK_K, 1, 64, 2};
        if (encode_tmap_bf16(&ta, g.A, 4, dims, strides, box)) return -1;
    } else {
        // zero-padded stream [B, Tp, D]; box = 64 channels x 128 frames
        uint64_t dims[3] = {(uint64_t)g.pos_dim, (uint64_t)g.pos_tp, (uint64_t)g.batches};
        uint64_t strides[2] = {(uint64_t)g.pos_dim * 2, (uint64_t)g.pos_tp * g.pos_dim * 2};
        uint32_t box[3] = {BLOCK_K, BLOCK_M, 1};
        if (encode_tmap_bf16(&ta, g.A, 3, dims, strides, box)) return -1;
    }
    if (g.a_mode == A_PLAIN) {
        static int use_pair = -1;
        if (use_pair < 0) { const char* v = getenv("SLSB_GEMM_PAIR"); use_pair = v ? atoi(v) : 1; }
        if (block_n == 256 && use_pair && (dp.tma_store || dp.res_tma || dp.red_add) && dp.kb_per_split == 0 && g.M >= 256) {
            // CTA-pair kernel: the W map's box is this CTA's 128 of the tile's 256 rows
            uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.N};
            uint64_t strides[1] = {(uint64_t)g.ldw * 2};
            uint32_t box[2] = {BLOCK_K, 128};
            if (encode_tmap_bf16(&tb, g.W, 2, dims, strides, box)) return -1;
            static const bool red_v1 = getenv("SLSB_RED_ADD_V1") && atoi(getenv("SLSB_RED_ADD_V1")) != 0;
            if (dp.red_add && !red_v1) {      // 16-column ping-pong reduce-add epilogue: 64-byte rows, SWIZZLE_64B
                uint64_t odims[2] = {(uint64_t)g.N, (uint64_t)g.M};
                uint32_t obox[2] = {16, BLOCK_M};
                uint64_t so[1] = {(uint64_t)g.ldc * 4};
                if (encode_tmap_f32_sw64(&to, g.out, 2, odims, so, obox)) return -1;
                dp.red_add = 2;
            }
            return launch_pair(ta, tb, to, tr, dp, num_sms, stream);
        }
        if (block_n == 256) return launch<256, A_PLAIN>(ta, tb, to, tr, dp, num_sms, stream);
        if (block_n == 128) return launch<128, A_PLAIN>(ta, tb, to, tr, dp, num_sms, stream);
        return launch<64, A_PLAIN>(ta, tb, to, tr, dp, num_sms, stream);
    } else if (g.a_mode == A_CONV) {
        if (block_n == 256) return launch<256, A_CONV>(ta, tb, to, tr, dp, num_sms, stream);
        if (block_n == 128) return launch<128, A_CONV>(ta, tb, to, tr, dp, num_sms, stream);
        return launch<64, A_CONV>(ta, tb, to, tr, dp, num_sms, stream);
    }
    if (g.a_mode == A_POS4) return launch<256, A_POS4>(ta, tb, to, tr, dp, num_sms, stream);
    return launch<64, A_POS>(ta, tb, to, tr, dp, num_sms, stream);
}

}  // namespace slsb

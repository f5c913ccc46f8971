// Multi-head self-attention over the fused qkv activations (fairseq MultiheadAttention, eval; ctor wav2vec2.py:1009-1014,
// call :1046-1052).  q arrives pre-scaled by d^-1/2 (folded into W_q/b_q at pack time); keys >= len_b get -inf.
//
//   attention_tc   : bf16, tcgen05.  One CTA per (utterance, head, 128-query tile); S = Q K^T accumulates in TMEM
//                    (128 lanes x Tp columns), four warps (thread == query row) run the fp32 softmax straight out of TMEM,
//                    write P (bf16) into a SWIZZLE_128B K-major smem tile that aliases the dead Q/K tiles, and a second
//                    tcgen05.mma (V as MN-major B operand, exactly the [key][d] layout TMA delivers) produces O in the
//                    TMEM columns S vacated.  256 TMEM columns + 96 KB smem per CTA -> two CTAs per SM overlap
//                    softmax with the other CTA's TMA/MMA.  T <= 256.  (attn_tc_v1_kernel, kept for A/B.)
//   attention_tc   : the production kernel further down (persistent, warp-specialised, P in TMEM), T <= 512.
//   attention_simt : fp32 math, fp32 or bf16 I/O, any T (key tiles of 64, online softmax).  fp32 verification mode and
//                    the T > 512 path.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace slsb {
namespace {

constexpr int HD = 64;   // head dim

// =================================================================================================
// SIMT
// =================================================================================================
constexpr int SQ_ROWS = 32;      // query rows per block (4 per warp)
constexpr int SK_TILE = 64;

template <typename T>
__global__ void __launch_bounds__(256) attn_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tn, int H, const int* __restrict__ lens) {
    extern __shared__ float simt_smem[];
    float (*Qs)[HD] = reinterpret_cast<float (*)[HD]>(simt_smem);
    float (*Ks)[HD + 1] = reinterpret_cast<float (*)[HD + 1]>(simt_smem + SQ_ROWS * HD);
    float (*Vs)[HD] = reinterpret_cast<float (*)[HD]>(simt_smem + SQ_ROWS * HD + SK_TILE * (HD + 1));
    float (*Ps)[SK_TILE][4] = reinterpret_cast<float (*)[SK_TILE][4]>(simt_smem + SQ_ROWS * HD + SK_TILE * (HD + 1) + SK_TILE * HD);

    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * SQ_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D3 = 3 * H * HD;
    const int len = lens ? min(lens[b], Tn) : Tn;
    const T* base = qkv + (long long)b * Tn * D3;

    for (int i = threadIdx.x; i < SQ_ROWS * HD; i += 256) {
        const int r = i / HD, d = i % HD;
        Qs[r][d] = (q0 + r < Tn) ? to_f32<T>(base[(long long)(q0 + r) * D3 + h * HD + d]) : 0.f;
    }
    float m[4], l[4], o[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) { m[r] = -INFINITY; l[r] = 0.f; o[r][0] = o[r][1] = 0.f; }

    for (int k0 = 0; k0 < len; k0 += SK_TILE) {
        __syncthreads();
        for (int i = threadIdx.x; i < SK_TILE * HD; i += 256) {
            const int j = i / HD, d = i % HD;
            const bool ok = k0 + j < len;
            Ks[j][d] = ok ? to_f32<T>(base[(long long)(k0 + j) * D3 + H * HD + h * HD + d]) : 0.f;
            Vs[j][d] = ok ? to_f32<T>(base[(long long)(k0 + j) * D3 + 2 * H * HD + h * HD + d]) : 0.f;
        }
        __syncthreads();
        float s[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r][0] = s[r][1] = 0.f;
#pragma unroll 8
        for (int d = 0; d < HD; ++d) {
            const float ka = Ks[lane][d], kb = Ks[lane + 32][d];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float q = Qs[warp * 4 + r][d];
                s[r][0] = fmaf(q, ka, s[r][0]);
                s[r][1] = fmaf(q, kb, s[r][1]);
            }
        }
        const bool ok0 = k0 + lane < len, ok1 = k0 + lane + 32 < len;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float s0 = ok0 ? s[r][0] : -INFINITY, s1 = ok1 ? s[r][1] : -INFINITY;
            const float mt = warp_max(fmaxf(s0, s1));
            const float mn = fmaxf(m[r], mt);            // finite: every tile has >= 1 valid key
            const float corr = __expf(m[r] - mn);
            const float p0 = ok0 ? __expf(s0 - mn) : 0.f, p1 = ok1 ? __expf(s1 - mn) : 0.f;
            l[r] = l[r] * corr + warp_sum(p0 + p1);
            o[r][0] *= corr; o[r][1] *= corr;
            m[r] = mn;
            Ps[warp][lane][r] = p0;
            Ps[warp][lane + 32][r] = p1;
        }
        __syncwarp();
#pragma unroll 4
        for (int j = 0; j < SK_TILE; ++j) {
            const float4 p = *reinterpret_cast<const float4*>(&Ps[warp][j][0]);
            const float va = Vs[j][lane], vb = Vs[j][lane + 32];
            o[0][0] = fmaf(p.x, va, o[0][0]); o[0][1] = fmaf(p.x, vb, o[0][1]);
            o[1][0] = fmaf(p.y, va, o[1][0]); o[1][1] = fmaf(p.y, vb, o[1][1]);
            o[2][0] = fmaf(p.z, va, o[2][0]); o[2][1] = fmaf(p.z, vb, o[2][1]);
            o[3][0] = fmaf(p.w, va, o[3][0]); o[3][1] = fmaf(p.w, vb, o[3][1]);
        }
        __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int t = q0 + warp * 4 + r;
        if (t >= Tn) continue;
        const float inv = 1.0f / l[r];
        T* op = out + ((long long)b * Tn + t) * (H * HD) + h * HD;
        op[lane] = from_f32<T>(o[r][0] * inv);
        op[lane + 32] = from_f32<T>(o[r][1] * inv);
    }
}

// =================================================================================================
// tcgen05
// =================================================================================================
constexpr int AQ = 128;                       // query rows per CTA
constexpr int kSmemQ = AQ * 128;              // 16 KB
constexpr int kSmemK = 256 * 128;             // 32 KB (Tp <= 256 keys)
constexpr int kSmemPExtra = 16 * 1024;        // P (64 KB) aliases Q + K + this
constexpr int kSmemV = 256 * 128;             // 32 KB
constexpr int kAttnSmem = kSmemQ + kSmemK + kSmemPExtra + kSmemV + 64 + 4 * 128 * 4;

constexpr int kAttnThreads = 288;     // warps 0-7: softmax / epilogue (two per TMEM lane quadrant), warp 8: TMA + MMA issue

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_tc_v1_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
               bf16* __restrict__ out, int Tn, int Tp, int H, const int* __restrict__ lens) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("slsb: dynamic smem base not 1024-aligned\n"); __trap(); }
    uint8_t* sQ = smem;
    uint8_t* sK = smem + kSmemQ;
    uint8_t* sP = smem;                                    // alias: valid once S = QK^T has completed
    uint8_t* sV = smem + kSmemQ + kSmemK + kSmemPExtra;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kSmemV);   // [0] loads, [1] S ready, [2] P ready, [3] O ready
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
    float* xch = reinterpret_cast<float*>(bars + 6);             // [2 halves][128 rows] row max, then [2][128] row sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AQ;
    const int len = lens ? min(lens[b], Tn) : Tn;
    const int D = H * HD;

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 256); mbar_init(&bars[3], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<256>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 8) {
        if (lane == 0) {
            mbar_expect_tx(&bars[0], (uint32_t)(kSmemQ + 2 * Tp * 128));
            tma_load_2d(sQ, &tm_q, &bars[0], h * HD, b * Tn + q0);
            tma_load_2d(sK, &tm_kv, &bars[0], D + h * HD, b * Tn);
            tma_load_2d(sV, &tm_kv, &bars[0], 2 * D + h * HD, b * Tn);
            mbar_wait(&bars[0], 0);
            tc_fence_after();
            {   // S[128 x Tp] = Q K^T
                const uint32_t idesc = make_idesc_bf16(AQ, Tp);
                const uint64_t da = make_smem_desc_sw128(smem_u32(sQ), 0, 1024);
                const uint64_t db = make_smem_desc_sw128(smem_u32(sK), 0, 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) tc_mma_f16(tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, k != 0);
                tc_commit(&bars[1]);
            }
            mbar_wait(&bars[2], 0);      // P written by the softmax warps
            tc_fence_after();
            {   // O[128 x 64] = P V ; V is [key][d] = MN-major B
                const uint32_t idesc = make_idesc_bf16(AQ, HD, 0, 1);
                const int ksteps = Tp / 16;
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint64_t da = make_smem_desc_sw128(smem_u32(sP) + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024);
                    const uint64_t db = make_smem_desc_sw128(smem_u32(sV) + kk * 2048, 32768, 1024);
                    tc_mma_f16(tmem, da, db, idesc, kk != 0);
                }
                tc_commit(&bars[3]);
            }
        }
    } else {
        // softmax + epilogue: thread == query row == TMEM lane; the two warps of a quadrant split the key columns
        const int q = warp & 3, half = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t trow = tmem + (uint32_t(q * 32) << 16);
        const int nchunk = Tp / 16;
        const int c_lo = half == 0 ? 0 : (nchunk + 1) / 2;
        const int c_hi = half == 0 ? (nchunk + 1) / 2 : nchunk;
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        // pass 1: row max over this warp's chunks (TMEM loads double-buffered in registers)
        float mx = -INFINITY;
        {
            uint32_t a[2][16];
            if (c_lo < c_hi) { tmem_ld_32x32b_x16(trow + c_lo * 16, a[0]); tmem_ld_wait(); }
#pragma unroll 1
            for (int c = c_lo; c < c_hi; c += 2) {
                if (c + 1 < c_hi) tmem_ld_32x32b_x16(trow + (c + 1) * 16, a[1]);
                if (c * 16 + 16 <= len) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(a[0][j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (c * 16 + j < len) mx = fmaxf(mx, __uint_as_float(a[0][j]));
                }
                tmem_ld_wait();
                if (c + 1 >= c_hi) break;
                if (c + 2 < c_hi) tmem_ld_32x32b_x16(trow + (c + 2) * 16, a[0]);
                if (c * 16 + 32 <= len) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(a[1][j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (c * 16 + 16 + j < len) mx = fmaxf(mx, __uint_as_float(a[1][j]));
                }
                tmem_ld_wait();
            }
        }
        xch[half * 128 + r] = mx;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]);            // at least one key is valid, so the row max is finite
        const float mxl = mx * 1.4426950408889634f;
        float sum = 0.f;
        auto emit = [&](const uint32_t (&a)[16], int c) {
            float p[16];
            const bool full = c * 16 + 16 <= len;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float e = ex2_approx(fmaf(__uint_as_float(a[j]), 1.4426950408889634f, -mxl));
                p[j] = (full || c * 16 + j < len) ? e : 0.f;
                sum += p[j];
            }
            uint8_t* blk = sP + (c >> 2) * 16384 + r * 128;
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                uint4 w;
                w.x = pack_bf16x2(p[8 * v + 0], p[8 * v + 1]); w.y = pack_bf16x2(p[8 * v + 2], p[8 * v + 3]);
                w.z = pack_bf16x2(p[8 * v + 4], p[8 * v + 5]); w.w = pack_bf16x2(p[8 * v + 6], p[8 * v + 7]);
                const int c8 = (c & 3) * 2 + v;
                *reinterpret_cast<uint4*>(blk + ((c8 ^ (r & 7)) << 4)) = w;
            }
        };
        {
            uint32_t a[2][16];
            if (c_lo < c_hi) { tmem_ld_32x32b_x16(trow + c_lo * 16, a[0]); tmem_ld_wait(); }
#pragma unroll 1
            for (int c = c_lo; c < c_hi; c += 2) {
                if (c + 1 < c_hi) tmem_ld_32x32b_x16(trow + (c + 1) * 16, a[1]);
                emit(a[0], c);
                tmem_ld_wait();
                if (c + 1 >= c_hi) break;
                if (c + 2 < c_hi) tmem_ld_32x32b_x16(trow + (c + 2) * 16, a[0]);
                emit(a[1], c + 1);
                tmem_ld_wait();
            }
        }
        xch[256 + half * 128 + r] = sum;
        fence_proxy_async_smem();     // generic-proxy smem writes of P -> visible to the tensor core (async proxy)
        tc_fence_before();
        mbar_arrive(&bars[2]);
        mbar_wait(&bars[3], 0);
        tc_fence_after();
        asm volatile("bar.sync 1, 256;" ::: "memory");        // partner's row sum is visible
        const float inv = 1.0f / (sum + xch[256 + (half ^ 1) * 128 + r]);
        const int t = q0 + r;
        {
            uint32_t a[32];
            tmem_ld_32x32b_x32(trow + half * 32, a);          // each warp of the pair writes 32 of the 64 output columns
            tmem_ld_wait();
            if (t < Tn) {
                uint4* op = reinterpret_cast<uint4*>(out + ((long long)b * Tn + t) * D + h * HD + half * 32);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    uint4 w;
                    w.x = pack_bf16x2(__uint_as_float(a[8 * v + 0]) * inv, __uint_as_float(a[8 * v + 1]) * inv);
                    w.y = pack_bf16x2(__uint_as_float(a[8 * v + 2]) * inv, __uint_as_float(a[8 * v + 3]) * inv);
                    w.z = pack_bf16x2(__uint_as_float(a[8 * v + 4]) * inv, __uint_as_float(a[8 * v + 5]) * inv);
                    w.w = pack_bf16x2(__uint_as_float(a[8 * v + 6]) * inv, __uint_as_float(a[8 * v + 7]) * inv);
                    op[v] = w;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<256>(tmem);
    }
}


// =================================================================================================
// tcgen05, persistent + warp-specialised (the production kernel)
//
//   item  = one (utterance, head): K and V are loaded ONCE per item and shared by its query tiles;
//   unit  = one 128-query tile of an item; units alternate between two softmax warpgroups (WG0 = warps 0-3, WG1 = warps 4-7),
//           each owning one 256-column half of TMEM;
//   warp 8 = TMA producer (2-stage ring of {Q tiles, K, V}, 96 KB / stage);
//   warps 9, 10 = one single-thread MMA issuer per TMEM half, so neither warpgroup ever waits behind the other one's
//            barrier or MMA issue.
//   S = Q K^T lands in TMEM (fp32, Tp columns); the thread that owns a query row (== TMEM lane) takes the row max and the
//   exponentials straight from TMEM in compact loops and writes P back INTO TMEM as packed bf16 over score columns it has
//   already consumed (tcgen05.st), so the second MMA (O = P V) takes its A operand from tensor memory and P never touches
//   shared memory; V is the MN-major B operand in exactly the [key][d] layout TMA delivers.  O lands in columns 192..255
//   of the same half, is pulled into registers, the half is handed back to the MMA warp at once, and the normalised rows
//   leave through a SWIZZLE_128B smem tile + 3-D TMA store (rows >= T are clipped by the tensor map).
// =================================================================================================
// narrow (T <= 256): 2 stages of {2 Q tiles, 256 keys of K, 256 of V}; wide (T <= 512): 1 stage of {4 Q tiles, 512 keys, 512 keys}
template <bool kWide> struct AttnGeo {
    static constexpr int STAGES = kWide ? 1 : 2;
    static constexpr int QCAP = kWide ? 4 : 2;            // 128-row query tiles per item
    static constexpr int KVCAP = kWide ? 512 : 256;       // keys per item
    static constexpr int SQ = QCAP * AQ * 128;
    static constexpr int SK = KVCAP * 128;
    static constexpr int SV = KVCAP * 128;
    static constexpr int STAGE = SQ + SK + SV;            // 96 KB / 192 KB
    static constexpr int OUT_OFFSET = STAGES * STAGE;     // 2 x [128 rows x 64 bf16] output staging tiles (one per warpgroup)
    static constexpr int BAR_OFFSET = OUT_OFFSET + 2 * 16384;
    static constexpr int SMEM = BAR_OFFSET + 256;
};
constexpr int P_THREADS = 384;                // 3 warpgroups: 2 x 4 softmax warps, then {TMA producer, 2 MMA issuers, 1 spare warp}
// Register re-allocation between the warpgroups (setmaxnreg): the kernel is compiled for 168 registers (65536 / 384); the service
// warpgroup gives registers back (56 are plenty for its descriptor arithmetic) and the two softmax warpgroups grow to 224, which
// keeps both TMEM load buffers, the exponentials and the 64 output values of a row in registers.  With a flat 168 the prefetched
// TMEM chunk was spilled to local memory right after every tcgen05.ld (8 STL.64 + 16 LDL per 16 scores: a third of the loop).
constexpr int P_REGS_SERVICE = 56, P_REGS_SOFTMAX = 224;
static_assert(128 * (168 - P_REGS_SERVICE) >= 2 * 128 * (P_REGS_SOFTMAX - 168), "register budget of the three warpgroups");
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
// three-input maximum (FMNMX3): half the instructions of the row-max pass
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
constexpr int P_OCOL = 192;                   // O accumulator columns inside a TMEM half
constexpr int P_KB_MAX = 192;                 // widest key block when an item needs several (S must stay clear of the live O columns)

// 32 score columns in registers -> four independent running maxima; columns >= lim are padding / other utterances / stale
__device__ __forceinline__ void max32(const uint32_t (&a)[32], int col0, int lim, float (&mx)[4]) {
    if (col0 + 32 <= lim) {
#pragma unroll
        for (int j = 0; j < 32; j += 8)
#pragma unroll
            for (int q = 0; q < 4; ++q) mx[q] = max3(mx[q], __uint_as_float(a[j + q]), __uint_as_float(a[j + 4 + q]));
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (col0 + j < lim) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(a[j]));
    }
}

// Key blocks: an item whose padded key count Tp fits a TMEM half (Tp <= 256) is ONE block: S -> row max -> P -> O, as described
// above.  A longer item (wide mode, Tp <= 512) is cut into nkb blocks of kbw <= 192 keys and scheduled in two rounds over the
// same TMEM columns: round 1 recomputes S block by block only for the row maxima, round 2 recomputes it again, exponentiates
// against the GLOBAL row max and accumulates O += P_kb V_kb.  QK^T is 4 MMAs per block (the tensor pipe idles ~85 % of this
// kernel), so recomputing it is cheaper than rescaling O, and the result is what a single block would give: no online re-base,
// no dependence on the blocking.
template <bool kWide>
__global__ void __launch_bounds__(P_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_out,
               int Tn, int Tp, int H, int n_items, int n_qt, int kvbox, int nkb_arg, int kbw_arg, const int* __restrict__ lens, int stagger,
               long long* __restrict__ trace) {
    using G = AttnGeo<kWide>;
    const int nkb = kWide ? nkb_arg : 1;          // narrow items are one block: compile-time, so that path keeps its codegen
    const int kbw = kWide ? kbw_arg : Tp;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("slsb: dynamic smem base not 1024-aligned\n"); __trap(); }
    // optional timeline of CTA 0 (slsb_op_attention_trace): trace[unit * 16 + event] = clock64()
#define ATT_TRACE(unit, ev) do { if (trace != nullptr && blockIdx.x == 0 && (unit) < 64) trace[(unit) * 16 + (ev)] = clock64(); } while (0)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::BAR_OFFSET);
    uint64_t* full = bars;                    // [STAGES] loads landed
    uint64_t* stage_free = bars + 2;          // [STAGES] every MMA that reads the stage has completed
    uint64_t* s_ready = bars + 4;             // [2] S block complete in TMEM half w (one completion per phase)
    uint64_t* p_ready = bars + 6;             // [2] S block consumed / P written back to TMEM half w (4 warp arrivals per phase)
    uint64_t* o_ready = bars + 8;             // [2] O = P V complete (once per unit)
    uint64_t* tmem_free = bars + 10;          // [2] O has been read out of half w (4 warp arrivals)
    uint64_t* pv_done = bars + 12;            // [2] P_kb V_kb complete: its P columns may be overwritten by the next S block
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = H * HD;
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items blockIdx.x, +grid, ...
    const int my_units = my_items * n_qt;
    const int nph = nkb == 1 ? 1 : 2 * nkb;   // phases per unit (see above)

    griddep_launch();
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        tma_prefetch_desc(&tm_out);
        for (int i = 0; i < G::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&stage_free[i], n_qt); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_ready[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_ready[i], 1); mbar_init(&tmem_free[i], 4); mbar_init(&pv_done[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == 9) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    griddep_wait();                     // the prologue overlapped the QKV GEMM's tail; qkv is visible from here

    if (warp >= 8) {
        reg_dealloc<P_REGS_SERVICE>();
    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int n_kvbox = (Tp + kvbox - 1) / kvbox;
            const uint32_t tx = (uint32_t)(n_qt * AQ * 128 + 2 * n_kvbox * kvbox * 128);
            for (int n = 0; n < my_items; ++n) {
                const int item = blockIdx.x + n * gridDim.x;
                const int b = item / H, h = item - b * H;
                const int st = n % G::STAGES;
                mbar_wait(&stage_free[st], ((n / G::STAGES) & 1) ^ 1);
                ATT_TRACE(n * n_qt, 9);
                uint8_t* base = smem + st * G::STAGE;
                mbar_expect_tx(&full[st], tx);
                for (int qt = 0; qt < n_qt; ++qt) tma_load_2d(base + qt * (AQ * 128), &tm_q, &full[st], h * HD, b * Tn + qt * AQ);
                for (int i = 0; i < n_kvbox; ++i) {
                    tma_load_2d(base + G::SQ + i * kvbox * 128, &tm_kv, &full[st], D + h * HD, b * Tn + i * kvbox);
                    tma_load_2d(base + G::SQ + G::SK + i * kvbox * 128, &tm_kv, &full[st], 2 * D + h * HD, b * Tn + i * kvbox);
                }
            }
        }
    } else if (warp == 9 || warp == 10) {
        // ===================== MMA issuers: warp 9 serves TMEM half 0, warp 10 half 1 (one thread each) =====================
        // Two independent issuers: issuing the 13 TS-MMAs of one half (~1k cycles of issue back-pressure) never delays the other
        // half's S, and neither warpgroup waits behind the other one's barrier.  `stagger` (SLSB_ATTN_STAGGER, default 0) can
        // start half 1 late to run the warpgroups in anti-phase; measured neutral (the softmax warps are issue-bound, not
        // MUFU-bound: profiles/r01_attention_notes.md).
        if (lane == 0) {
            const int w = warp - 9;
            const uint32_t idesc_o = make_idesc_bf16(AQ, HD, 0, 1);
            const uint32_t th = tmem + w * 256;
            uint32_t ph = 0, pvn = 0;
            for (int u = w, k = 0; u < my_units; u += 2, ++k) {
                const int n = u / n_qt, qt = u - n * n_qt, st = n % G::STAGES;
                mbar_wait(&full[st], (n / G::STAGES) & 1);
                ATT_TRACE(u, 10);
                if (k == 0 && w == 1 && stagger > 0) { const long long t0 = clock64(); while (clock64() - t0 < stagger) { } }
                mbar_wait(&tmem_free[w], (k & 1) ^ 1);
                tc_fence_after();
                ATT_TRACE(u, 0);
                const uint32_t sq = smem_u32(smem + st * G::STAGE + qt * (AQ * 128));
                const uint32_t sk = smem_u32(smem + st * G::STAGE + G::SQ);
                const uint32_t sv = smem_u32(smem + st * G::STAGE + G::SQ + G::SK);
                const uint64_t da = make_smem_desc_sw128(sq, 0, 1024);
                for (int p = 0; p < nph; ++p) {
                    const int kb = p < nkb ? p : p - nkb;
                    const bool do_pv = nkb == 1 || p >= nkb;
                    const int kb0 = kb * kbw, wk = min(kbw, Tp - kb0);
                    // S block = Q_qt K[kb0 : kb0 + wk]^T -> TMEM half w, columns [0, wk)
                    const uint32_t idesc_s = make_idesc_bf16(AQ, wk);
                    const uint64_t dk = make_smem_desc_sw128(sk + kb0 * 128, 0, 1024);
#pragma unroll
                    for (int kk = 0; kk < HD / 16; ++kk) tc_mma_f16(th, da + uint64_t(2 * kk), dk + uint64_t(2 * kk), idesc_s, kk != 0);
                    tc_commit(&s_ready[w]);
                    if (p == 0) ATT_TRACE(u, 1);
                    mbar_wait(&p_ready[w], ph & 1); ++ph;
                    tc_fence_after();
                    if (!do_pv) continue;
                    if (p + 1 == nph) ATT_TRACE(u, 2);
                    // O (+)= P_kb V_kb ; A = packed bf16 P in TMEM, B = V [key][d] (MN-major)
                    uint64_t db = make_smem_desc_sw128(sv + kb0 * 128, 32768, 1024);
                    for (int kk = 0; kk < wk / 16; ++kk) {
                        tc_mma_f16_ts(th + P_OCOL, th + kk * 8, db, idesc_o, (kb | kk) != 0);
                        db += 2048 >> 4;                 // next 16 keys of V
                    }
                    if (p + 1 < nph) {                   // the next S block overwrites the P columns this product reads
                        tc_commit(&pv_done[w]);
                        mbar_wait(&pv_done[w], pvn & 1); ++pvn;
                        tc_fence_after();
                    }
                }
                tc_commit(&o_ready[w]);
                tc_commit(&stage_free[st]);          // one arrival per unit of the item (barrier count n_qt): all its MMAs are done
                ATT_TRACE(u, 3);
            }
        }
    }
    } else {
        reg_alloc<P_REGS_SOFTMAX>();
        // ===================== softmax + epilogue warpgroups: thread == query row == TMEM lane =====================
        const int w = warp >> 2, q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t trow = tmem + (uint32_t(q * 32) << 16) + w * 256;
        uint8_t* stage_tile = smem + G::OUT_OFFSET + w * 16384;
        const int bar_id = 1 + w;
        const bool issuer = r == 0;
        constexpr float kLog2e = 1.4426950408889634f;
        uint32_t ph = 0;
        for (int u = w, k = 0; u < my_units; u += 2, ++k) {
            const int n = u / n_qt, qt = u - n * n_qt;
            const int item = blockIdx.x + n * gridDim.x;
            const int b = item / H, h = item - b * H;
            const int len = lens ? min(lens[b], Tn) : Tn;
            const int rows_valid = min(AQ, Tn - qt * AQ);
            const bool active = q * 32 < rows_valid;           // warp-uniform: quadrants with no valid query row skip the math
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            float mxl = 0.f;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            for (int p = 0; p < nph; ++p) {
                const int kb = p < nkb ? p : p - nkb;
                const bool do_max = nkb == 1 || p < nkb, do_exp = nkb == 1 || p >= nkb;
                const int kb0 = kb * kbw, wk = min(kbw, Tp - kb0);
                const int lim = min(len - kb0, wk);            // valid columns of this block (may be <= 0: all padding)
                const int npiece = (wk + 31) / 32, nchunk = wk / 16;
                mbar_wait(&s_ready[w], ph & 1); ++ph;
                tc_fence_after();
                if (issuer && p == 0) ATT_TRACE(u, 4);
                if (active) {
                    if (do_max) {
                        // ---- row max, 32-column TMEM loads double-buffered in registers (a piece may run past the block into stale
                        // columns of the half: masked by lim)
                        // three pieces in flight per wait (96 registers: the softmax warpgroups own 224 each): one piece per wait
                        // made this pass a chain of TMEM load latencies (~210 cycles per piece, 1.5k cycles per unit)
                        uint32_t a[3][32];
#pragma unroll 1
                        for (int c = 0; c < npiece; c += 3) {
#pragma unroll
                            for (int g = 0; g < 3; ++g)
                                if (c + g < npiece) tmem_ld_32x32b_x32(trow + (c + g) * 32, a[g]);
                            tmem_ld_wait();
#pragma unroll
                            for (int g = 0; g < 3; ++g)
                                if (c + g < npiece) max32(a[g], (c + g) * 32, lim, mx4);
                        }
                        if (issuer && p == 0) ATT_TRACE(u, 5);
                    }
                    if (do_exp) {
                        // ---- p = exp(s - max) -> packed bf16 written back over the columns of S this thread has already consumed
                        // (P chunk c covers 32-bit columns [8c, 8c+8); the S columns still to be read start at 16(c+1))
                        if (p == 0 || p == nkb) mxl = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * kLog2e;   // key 0 is valid: finite
                        auto exp_chunk = [&](const uint32_t (&cur)[16], int c) {
                            float pe[16];
                            if (c * 16 + 16 <= lim) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) pe[j] = ex2_approx(fmaf(__uint_as_float(cur[j]), kLog2e, -mxl));
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) pe[j] = c * 16 + j < lim ? ex2_approx(fmaf(__uint_as_float(cur[j]), kLog2e, -mxl)) : 0.f;
                            }
                            s0 += (pe[0] + pe[4]) + (pe[8] + pe[12]); s1 += (pe[1] + pe[5]) + (pe[9] + pe[13]);
                            s2 += (pe[2] + pe[6]) + (pe[10] + pe[14]); s3 += (pe[3] + pe[7]) + (pe[11] + pe[15]);
                            uint32_t pk[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(pe[2 * j], pe[2 * j + 1]);
                            tmem_st_32x32b_x8(trow + c * 8, pk);
                        };
                        uint32_t a0[16], a1[16];
                        tmem_ld_32x32b_x16(trow, a0);
                        tmem_ld_wait();
#pragma unroll 1
                        for (int c = 0; c < nchunk; c += 2) {
                            if (c + 1 < nchunk) tmem_ld_32x32b_x16(trow + (c + 1) * 16, a1);
                            exp_chunk(a0, c);
                            tmem_ld_wait();
                            if (c + 1 >= nchunk) break;
                            if (c + 2 < nchunk) tmem_ld_32x32b_x16(trow + (c + 2) * 16, a0);
                            exp_chunk(a1, c + 1);
                            tmem_ld_wait();
                        }
                        tmem_st_wait();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_ready[w]);
            }
            const float sum = (s0 + s1) + (s2 + s3);
            if (issuer) ATT_TRACE(u, 6);
            // ---- epilogue: O / sum -> bf16 -> swizzled smem tile -> TMA store
            mbar_wait(&o_ready[w], k & 1);
            tc_fence_after();
            if (issuer) ATT_TRACE(u, 7);
            uint32_t o[64];
            if (active) {
                tmem_ld_32x32b_x32(trow + P_OCOL, o);
                tmem_ld_32x32b_x32(trow + P_OCOL + 32, o + 32);
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_free[w]);          // the MMA warp may start S(u + 2) while we store
            if (issuer) { ATT_TRACE(u, 8); tma_store_wait_read<0>(); }       // the previous store no longer reads the staging tile
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (active) {
                const float inv = 1.0f / sum;
                uint8_t* srow = stage_tile + r * 128;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    uint4 wv;
                    wv.x = pack_bf16x2(__uint_as_float(o[8 * c8 + 0]) * inv, __uint_as_float(o[8 * c8 + 1]) * inv);
                    wv.y = pack_bf16x2(__uint_as_float(o[8 * c8 + 2]) * inv, __uint_as_float(o[8 * c8 + 3]) * inv);
                    wv.z = pack_bf16x2(__uint_as_float(o[8 * c8 + 4]) * inv, __uint_as_float(o[8 * c8 + 5]) * inv);
                    wv.w = pack_bf16x2(__uint_as_float(o[8 * c8 + 6]) * inv, __uint_as_float(o[8 * c8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(srow + ((c8 ^ (r & 7)) << 4)) = wv;
                }
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (issuer) {
                tma_store_3d(&tm_out, stage_tile, h * HD, qt * AQ, b);     // rows t >= Tn fall outside the (d, t, b) tensor: clipped
                tma_store_commit();
            }
        }
        if (issuer) tma_store_wait<0>();
    }
#undef ATT_TRACE
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

}  // namespace

int attention_simt(const void* qkv, void* out, int io_bf16, int B, int T, int H, const int* lens, cudaStream_t stream) {
    dim3 grid((T + SQ_ROWS - 1) / SQ_ROWS, H, B);
    constexpr int kSmem = (SQ_ROWS * HD + SK_TILE * (HD + 1) + SK_TILE * HD + 8 * SK_TILE * 4) * (int)sizeof(float);
    static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
    if (first_use_on_device(&configured_on)) {
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(attn_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(attn_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    }
    if (io_bf16) attn_simt_kernel<bf16><<<grid, 256, kSmem, stream>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), T, H, lens);
    else attn_simt_kernel<float><<<grid, 256, kSmem, stream>>>(static_cast<const float*>(qkv), static_cast<float*>(out), T, H, lens);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int attention_tc_v1(const void* qkv, void* out, int B, int T, int H, const int* lens, int num_sms, cudaStream_t stream) {
    (void)num_sms;
    const int Tp = (T + 15) / 16 * 16;
    if (Tp > 256) { set_error("attention_tc_v1: T=%d > 256 (use attention_simt)", T); return -1; }
    const int D = H * HD;
    CUtensorMap tq, tkv;
    uint64_t dims[2] = {(uint64_t)(3 * D), (uint64_t)B * T};
    uint64_t strides[1] = {(uint64_t)(3 * D) * 2};
    uint32_t boxq[2] = {HD, AQ}, boxkv[2] = {HD, (uint32_t)Tp};
    if (encode_tmap_bf16(&tq, qkv, 2, dims, strides, boxq)) return -1;
    if (encode_tmap_bf16(&tkv, qkv, 2, dims, strides, boxkv)) return -1;
    static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
    if (first_use_on_device(&configured_on)) {
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    }
    dim3 grid((T + AQ - 1) / AQ, H, B);
    attn_tc_v1_kernel<<<grid, kAttnThreads, kAttnSmem, stream>>>(tq, tkv, static_cast<bf16*>(out), T, Tp, H, lens);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace slsb

namespace slsb {
int attention_tc(const void* qkv, void* out, int B, int T, int H, const int* lens, int num_sms, cudaStream_t stream, long long* trace) {
    if (T > 512) { set_error("attention_tc: T=%d > 512 (use attention_simt)", T); return -1; }
    const bool wide = T > 256;
    // padded key count: a multiple of 16 (MMA k-steps of P V); wide items load K / V as two boxes of Tp / 2 keys (TMA boxes hold
    // at most 256 rows), so there Tp is a multiple of 32
    const int Tp = wide ? (T + 31) / 32 * 32 : (T + 15) / 16 * 16;
    const int kvbox = wide ? Tp / 2 : Tp;
    const int nkb = wide ? (Tp + P_KB_MAX - 1) / P_KB_MAX : 1;
    const int kbw = wide ? ((Tp + nkb - 1) / nkb + 15) / 16 * 16 : Tp;
    const int D = H * HD;
    CUtensorMap tq, tkv, to;
    uint64_t dims[2] = {(uint64_t)(3 * D), (uint64_t)B * T};
    uint64_t strides[1] = {(uint64_t)(3 * D) * 2};
    uint32_t boxq[2] = {HD, AQ}, boxkv[2] = {HD, (uint32_t)kvbox};
    if (encode_tmap_bf16(&tq, qkv, 2, dims, strides, boxq)) return -1;
    if (encode_tmap_bf16(&tkv, qkv, 2, dims, strides, boxkv)) return -1;
    // output viewed as (d, t, b) so that a 128-row store box is clipped at the end of ITS utterance
    uint64_t odims[3] = {(uint64_t)D, (uint64_t)T, (uint64_t)B};
    uint64_t ostrides[2] = {(uint64_t)D * 2, (uint64_t)T * D * 2};
    uint32_t obox[3] = {HD, AQ, 1};
    if (encode_tmap_bf16(&to, out, 3, odims, ostrides, obox)) return -1;
    static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
    if (first_use_on_device(&configured_on)) {
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnGeo<false>::SMEM));
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnGeo<true>::SMEM));
    }
    const int n_items = B * H, n_qt = (T + AQ - 1) / AQ;
    const int grid = n_items < num_sms ? n_items : num_sms;
    static int stagger = -1;
    if (stagger < 0) { const char* sv = getenv("SLSB_ATTN_STAGGER"); stagger = sv ? atoi(sv) : 0; }
    if (wide)
        SLSB_CUDA_CHECK(launch_pdl(attn_tc_kernel<true>, dim3(grid), dim3(P_THREADS), AttnGeo<true>::SMEM, stream, tq, tkv, to, T, Tp, H, n_items, n_qt,
                                   kvbox, nkb, kbw, lens, stagger, trace));
    else
        SLSB_CUDA_CHECK(launch_pdl(attn_tc_kernel<false>, dim3(grid), dim3(P_THREADS), AttnGeo<false>::SMEM, stream, tq, tkv, to, T, Tp, H, n_items, n_qt,
                                   kvbox, nkb, kbw, lens, stagger, trace));
    return 0;
}
}  // namespace slsb

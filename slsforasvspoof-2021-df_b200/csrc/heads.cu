// Classifier-head kernels (HBM-bound, integer/selection work stays on CUDA cores):
//   * exact per-row top-k by 4-pass radix select (threshold + canonical lowest-index tie cut), never sorting;
//   * streaming mean-pool of the kept activations in a fixed order (bit-stable, no atomics);
//   * window top-k (model_window_topk.py:118-203) as window sums -> select -> votes -> select;
//   * LayerNorm + MLP + log-softmax classifier (model.py:183-189);
//   * SLS layer weighting / fused weighted sum + BN + SELU + 3x3 max-pool / tail (model_backup.py:186-202 + upstream).
#include "common.cuh"
#include "kernels.h"
#include <math.h>

namespace slsb {
namespace {

__device__ __forceinline__ uint32_t order_key(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// inclusive suffix sum over 256 per-thread values (thread t gets sum_{d >= t} v[d]); scratch: 8 ints
__device__ __forceinline__ int block_suffix_sum_256(int v, int* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_down_sync(0xffffffffu, s, o);
        if (lane + o < 32) s += n;
    }
    if (lane == 0) scratch[warp] = s;
    __syncthreads();
    int add = 0;
    for (int w = warp + 1; w < 8; ++w) add += scratch[w];
    __syncthreads();
    return s + add;
}
// exclusive prefix sum over 256 per-thread values
__device__ __forceinline__ int block_prefix_excl_256(int v, int* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += n;
    }
    if (lane == 31) scratch[warp] = s;
    __syncthreads();
    int add = 0;
    for (int w = 0; w < warp; ++w) add += scratch[w];
    __syncthreads();
    return s + add - v;
}

__device__ __forceinline__ bool kept(float sel, int idx, float thr, int cut) { return sel > thr || (sel == thr && idx < cut); }

__global__ void densify_kernel(const float* __restrict__ acts, const float* __restrict__ sel, const float* __restrict__ thr,
                               const int* __restrict__ cut, float* __restrict__ out, long long rows, int D) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= rows * D) return;
    const long long row = i4 / D;
    const int c = (int)(i4 - row * D);
    const float t = thr[row]; const int ct = cut[row];
    const float4 a = *reinterpret_cast<const float4*>(acts + i4);
    const float4 s = *reinterpret_cast<const float4*>(sel + i4);
    float4 o;
    o.x = kept(s.x, c, t, ct) ? a.x : 0.f; o.y = kept(s.y, c + 1, t, ct) ? a.y : 0.f;
    o.z = kept(s.z, c + 2, t, ct) ? a.z : 0.f; o.w = kept(s.w, c + 3, t, ct) ? a.w : 0.f;
    *reinterpret_cast<float4*>(out + i4) = o;
}

// compact form of the kept activations: one warp per row, ascending feature index, slots past the row's count get idx -1
__global__ void __launch_bounds__(256) compact_kept_kernel(const float* __restrict__ acts, const float* __restrict__ sel, const float* __restrict__ thr,
                                                           const int* __restrict__ cut, int* __restrict__ idx_out, float* __restrict__ val_out,
                                                           int* __restrict__ count_out, long long rows, int D, int k) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float t = thr[row]; const int ct = cut[row];
    int n = 0;
    for (int c0 = 0; c0 < D; c0 += 32) {
        const int c = c0 + lane;
        const bool keep = kept(sel[row * D + c], c, t, ct);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int pos = n + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < k) { idx_out[row * k + pos] = c; val_out[row * k + pos] = acts[row * D + c]; }
        n += __popc(m);
    }
    if (n > k) n = k;
    for (int j = n + lane; j < k; j += 32) { idx_out[row * k + j] = -1; val_out[row * k + j] = 0.f; }
    if (lane == 0 && count_out) count_out[row] = n;
}

__global__ void mean_pool_kept_kernel(const float* __restrict__ acts, const float* __restrict__ sel, const float* __restrict__ thr,
                                      const int* __restrict__ cut, float* __restrict__ pooled, int T, int D, const int* __restrict__ lens) {
    const int b = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= D) return;
    const int len = lens ? min(lens[b], T) : T;
    float s = 0.f;
    for (int t = 0; t < len; ++t) {
        const long long row = (long long)b * T + t;
        const float v = sel[row * D + f];
        if (kept(v, f, thr[row], cut[row])) s += acts[row * D + f];
    }
    pooled[(long long)b * D + f] = s / (float)len;
}

// =================================================================================================
// Fused selection + pooling (the scoring path of H-SAE and H-WIN): the fp32 activations are read ONCE per selection, the
// selection runs on register-resident rows, and neither window sums, votes nor dense codes reach HBM.
//
//   canonical pooling order (every path uses it, so retained and non-retained forwards give identical bits): an utterance is
//   cut into chunks of kSelRows frames; a chunk's kept activations are summed per feature in frame order (registers), the
//   chunk partials [B][n_chunks][D] are then summed in chunk order and divided by the frame count (pool_finish_kernel).
//   Chunk boundaries depend on T alone: a clip's pooled vector does not depend on the batch around it.
//
//   block = 256 threads, thread t owns the contiguous feature slice [t * EPT, (t + 1) * EPT) of every row it touches, so the
//   per-window keep masks (one 32-bit word per thread and window) never leave the thread that produced them.
// =================================================================================================
constexpr int kSelRows = 8;

struct SelSmem {
    int hist[256];
    int scratch[2][8];
    uint32_t prefix;
    int krem;
    int cut;
};

// exact k-th largest key of the block's row (keys in registers) + tie cut, 4 radix passes of 8 bits; all 256 threads call it.
// 4 barriers per pass: the histogram is cleared right after the barrier that published the previous digit, and the suffix-sum
// scratch is double-buffered by pass.
template <int EPT>
__device__ __forceinline__ void block_select(const uint32_t (&key)[EPT], int k, SelSmem& sm, uint32_t& thr_key, int& cut) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t prefix = 0, mask = 0;
    int krem = k;
    constexpr uint32_t kZeroKey = 0x80000000u;                // order_key(+0.0f)
    int zc = 0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) zc += (key[i] == kZeroKey);
    const int zc_warp = __reduce_add_sync(0xffffffffu, zc);
#pragma unroll 1
    for (int shift = 24, pass = 0; shift >= 0; shift -= 8, ++pass) {
        sm.hist[tid] = 0;
        __syncthreads();
        // Post-ReLU rows are about half exact zeros: as plain shared-memory atomics they all hit ONE bin and serialise 32 ways
        // per warp instruction (the first pass alone cost ~4 us per row).  The zeros are counted once per row (warp reduction,
        // above the loop) and enter the histogram with one atomic per warp; the other elements spread over the exponent bins.
        // (A __match_any_sync aggregation of every digit was measured slower than the plain atomics: 0.39 vs 0.26 ms per batch.)
        const bool zeros_in = (kZeroKey & mask) == prefix;
#pragma unroll
        for (int i = 0; i < EPT; ++i)
            if (key[i] != kZeroKey && (key[i] & mask) == prefix) atomicAdd(&sm.hist[(key[i] >> shift) & 255], 1);
        if (zeros_in && lane == 0 && zc_warp) atomicAdd(&sm.hist[(kZeroKey >> shift) & 255], zc_warp);
        __syncthreads();
        const int h = sm.hist[tid];
        int sfx = h;                                          // inclusive suffix sum over the 256 digits
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_down_sync(0xffffffffu, sfx, o);
            if (lane + o < 32) sfx += n;
        }
        if (lane == 0) sm.scratch[pass & 1][warp] = sfx;
        __syncthreads();
        for (int w = warp + 1; w < 8; ++w) sfx += sm.scratch[pass & 1][w];
        if (sfx >= krem && sfx - h < krem) {                  // exactly one digit holds the k-th largest key
            sm.prefix = prefix | (uint32_t(tid) << shift);
            sm.krem = krem - (sfx - h);
        }
        __syncthreads();
        prefix = sm.prefix; krem = sm.krem;
        mask |= 255u << shift;
    }
    // krem (>= 1) of the entries equal to the threshold are kept, lowest indices first
    int eq = 0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) eq += (key[i] == prefix);
    int inc = eq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) sm.scratch[0][warp] = inc;
    __syncthreads();
    int before = inc - eq;
    for (int w = 0; w < warp; ++w) before += sm.scratch[0][w];
    if (before < krem && before + eq >= krem) {
        int need = krem - before, c = 0;
#pragma unroll
        for (int i = 0; i < EPT; ++i)
            if (key[i] == prefix && need > 0) { --need; c = tid * EPT + i + 1; }
        sm.cut = c;
    }
    __syncthreads();
    thr_key = prefix;
    cut = sm.cut;
}

template <int EPT>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[EPT]) {
#pragma unroll
    for (int i = 0; i < EPT; i += 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p + threadIdx.x * EPT + i));
        v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
    }
}

// stand-alone selection (slsb_op_topk, other callers): one block (256 threads) per row
template <int EPT>
__global__ void __launch_bounds__(256) topk_threshold_kernel(const float* __restrict__ x, int D, int k, float* __restrict__ thr, int* __restrict__ tie_cut) {
    __shared__ SelSmem sm;
    const long long row = blockIdx.x;
    float v[EPT];
    load_row<EPT>(x + row * D, v);
    uint32_t key[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) key[i] = order_key(v[i]);
    uint32_t tk; int cut;
    block_select<EPT>(key, k, sm, tk, cut);
    if (threadIdx.x == 0) { thr[row] = key_to_float(tk); tie_cut[row] = cut; }
}

// H-SAE: grid (n_chunks, B).  Per row: select on the activations, thr / tie_cut written for the on-request consumers (dense /
// compact codes), kept activations of frames < len added to the chunk partial.  partial == nullptr: selection only.
template <int EPT>
__global__ void __launch_bounds__(256, EPT <= 16 ? 3 : 1) topk_pool_kernel(const float* __restrict__ acts, int T, int k, const int* __restrict__ lens,
                                                        float* __restrict__ thr, int* __restrict__ tie_cut, float* __restrict__ partial) {
    __shared__ SelSmem sm;
    constexpr int D = EPT * 256;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int len = lens ? min(lens[b], T) : T;
    const int t0 = chunk * kSelRows, t1 = min(t0 + kSelRows, T);
    float pool[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) pool[i] = 0.f;
    float nxt[EPT];
    load_row<EPT>(acts + ((long long)b * T + t0) * D, nxt);
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
        const long long row = (long long)b * T + t;
        uint32_t key[EPT];                                              // the row lives in registers as its order keys (a bijection)
#pragma unroll
        for (int i = 0; i < EPT; ++i) key[i] = order_key(nxt[i]);
        if (t + 1 < t1) load_row<EPT>(acts + (row + 1) * D, nxt);       // next row in flight behind this row's selection
        uint32_t tk; int cut;
        block_select<EPT>(key, k, sm, tk, cut);
        const float th = key_to_float(tk);
        if (threadIdx.x == 0) { thr[row] = th; tie_cut[row] = cut; }
        if (partial != nullptr && t < len) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const float v = key_to_float(key[i]);
                if (kept(v, threadIdx.x * EPT + i, th, cut)) pool[i] += v;
            }
        }
    }
    if (partial != nullptr) {
        float* dst = partial + ((long long)b * gridDim.x + chunk) * D + threadIdx.x * EPT;
#pragma unroll
        for (int i = 0; i < EPT; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(pool[i], pool[i + 1], pool[i + 2], pool[i + 3]);
    }
}

// pooled[b][f] = (sum over chunks, in chunk order, of partial[b][c][f]) / len_b
__global__ void pool_finish_kernel(const float* __restrict__ partial, int n_chunks, int T, int D, const int* __restrict__ lens, float* __restrict__ pooled) {
    const int b = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= D) return;
    const int len = lens ? min(lens[b], T) : T;
    float s = 0.f;
    for (int c = 0; c < n_chunks; ++c) s += partial[((long long)b * n_chunks + c) * D + f];
    pooled[(long long)b * D + f] = s / (float)len;
}

// H-WIN step 1 (model_window_topk.py:153-165): grid (nw, B).  Window sums in frame order (registers), per-window top-k, and the
// keep decision of this thread's EPT features as one 32-bit word: wmask[(b * nw + w) * 256 + t].
template <int EPT>
__global__ void __launch_bounds__(256) window_select_kernel(const float* __restrict__ acts, int T, int k, int window, int stride, int nw,
                                                            uint32_t* __restrict__ wmask) {
    static_assert(EPT <= 32, "one mask word per thread");
    __shared__ SelSmem sm;
    constexpr int D = EPT * 256;
    const int b = blockIdx.y, w = blockIdx.x;
    float s[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) s[i] = 0.f;
    const float* base = acts + ((long long)b * T + (long long)w * stride) * D;
#pragma unroll 2
    for (int j = 0; j < window; ++j) {
        float v[EPT];
        load_row<EPT>(base + (long long)j * D, v);
#pragma unroll
        for (int i = 0; i < EPT; ++i) s[i] += v[i];
    }
    uint32_t key[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) key[i] = order_key(s[i]);
    uint32_t tk; int cut;
    block_select<EPT>(key, k, sm, tk, cut);
    const float th = key_to_float(tk);
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) m |= kept(s[i], threadIdx.x * EPT + i, th, cut) ? (1u << i) : 0u;
    wmask[((long long)b * nw + w) * 256 + threadIdx.x] = m;
}

// H-WIN step 2 (model_window_topk.py:167-197): grid (n_chunks, B).  Per frame: votes = activation x (number of covering windows
// that selected the feature), accumulated in window order as the reference does; per-frame top-k on the votes; kept ACTIVATIONS
// pooled.  votes_out (optional, on request only) materialises the votes for the dense / compact code consumers.
template <int EPT>
__global__ void __launch_bounds__(256, EPT <= 16 ? 3 : 1) window_vote_pool_kernel(const float* __restrict__ acts, const uint32_t* __restrict__ wmask, int T, int k,
                                                               int window, int stride, int nw, float* __restrict__ thr, int* __restrict__ tie_cut,
                                                               float* __restrict__ votes_out, float* __restrict__ partial) {
    __shared__ SelSmem sm;
    constexpr int D = EPT * 256;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int t0 = chunk * kSelRows, t1 = min(t0 + kSelRows, T);
    float pool[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) pool[i] = 0.f;
    float nxt[EPT];
    load_row<EPT>(acts + ((long long)b * T + t0) * D, nxt);
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
        const long long row = (long long)b * T + t;
        float a[EPT], v[EPT];
#pragma unroll
        for (int i = 0; i < EPT; ++i) { a[i] = nxt[i]; v[i] = 0.f; }
        if (t + 1 < t1) load_row<EPT>(acts + (row + 1) * D, nxt);
        const int w_hi = min(t / stride, nw - 1);
        const int w_lo = (t - window + 1 > 0) ? (t - window + stride) / stride : 0;
        for (int w = w_lo; w <= w_hi; ++w) {                   // ascending window order (:175-185)
            const uint32_t m = __ldg(wmask + ((long long)b * nw + w) * 256 + threadIdx.x);
#pragma unroll
            for (int i = 0; i < EPT; ++i)
                if ((m >> i) & 1u) v[i] += a[i];
        }
        uint32_t key[EPT];
#pragma unroll
        for (int i = 0; i < EPT; ++i) key[i] = order_key(v[i]);
        if (votes_out != nullptr) {
            float* dst = votes_out + row * D + threadIdx.x * EPT;
#pragma unroll
            for (int i = 0; i < EPT; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        uint32_t tk; int cut;
        block_select<EPT>(key, k, sm, tk, cut);                        // the votes live on as their order keys (a bijection)
        const float th = key_to_float(tk);
        if (threadIdx.x == 0) { thr[row] = th; tie_cut[row] = cut; }
#pragma unroll
        for (int i = 0; i < EPT; ++i)
            if (kept(key_to_float(key[i]), threadIdx.x * EPT + i, th, cut)) pool[i] += a[i];
    }
    if (partial != nullptr) {
        float* dst = partial + ((long long)b * gridDim.x + chunk) * D + threadIdx.x * EPT;
#pragma unroll
        for (int i = 0; i < EPT; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(pool[i], pool[i + 1], pool[i + 2], pool[i + 3]);
    }
}

__global__ void mean_pool_kernel(const float* __restrict__ x, float* __restrict__ pooled, int T, int D, const int* __restrict__ lens) {
    const int b = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= D) return;
    const int len = lens ? min(lens[b], T) : T;
    float s = 0.f;
    for (int t = 0; t < len; ++t) s += x[((long long)b * T + t) * D + f];
    pooled[(long long)b * D + f] = s / (float)len;
}

__global__ void window_sums_kernel(const float* __restrict__ acts, float* __restrict__ sums, int T, int D, int window, int stride, int nw) {
    const int b = blockIdx.z, w = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= D) return;
    float s = 0.f;
    for (int j = 0; j < window; ++j) s += acts[((long long)b * T + w * stride + j) * D + f];
    sums[((long long)b * nw + w) * D + f] = s;
}

__global__ void window_votes_kernel(const float* __restrict__ acts, const float* __restrict__ sums, const float* __restrict__ thr_w,
                                    const int* __restrict__ cut_w, float* __restrict__ votes, int T, int D, int window, int stride, int nw) {
    const int b = blockIdx.z, t = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= D) return;
    const float a = acts[((long long)b * T + t) * D + f];
    float v = 0.f;
    const int w_hi = min(t / stride, nw - 1);
    const int w_lo = (t - window + 1 > 0) ? (t - window + stride) / stride : 0;
    for (int w = w_lo; w <= w_hi; ++w) {         // ascending window order, as the reference accumulates (:175-185)
        const long long wr = (long long)b * nw + w;
        if (kept(sums[wr * D + f], f, thr_w[wr], cut_w[wr])) v += a;
    }
    votes[((long long)b * T + t) * D + f] = v;
}

// ---- classifier: LN(D) -> Linear(D, Hd) -> ReLU -> Linear(Hd, 2) -> log_softmax
// stage 1: grid (Hd / 32, B): every block normalises its utterance's pooled row (cheap, redundant) and computes 32 hidden
// units, one warp per unit with coalesced weight reads; stage 2: one warp per utterance.
constexpr int CLS_UNITS = 32;
__global__ void __launch_bounds__(256) classifier_hidden_kernel(const float* __restrict__ pooled, int D, int Hd, const float* __restrict__ ln_w,
                                                                const float* __restrict__ ln_b, const float* __restrict__ w1,
                                                                const float* __restrict__ b1, float* __restrict__ hidden) {
    extern __shared__ float cls_smem[];   // D
    float* xs = cls_smem;
    __shared__ float red[8];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* x = pooled + (long long)b * D;
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) { const float v = x[i]; xs[i] = v; s += v; }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float mean = 0.f;
    for (int w = 0; w < 8; ++w) mean += red[w];
    mean /= (float)D;
    __syncthreads();
    float q = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) { const float d = xs[i] - mean; q = fmaf(d, d, q); }
    q = warp_sum(q);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    float var = 0.f;
    for (int w = 0; w < 8; ++w) var += red[w];
    const float rstd = 1.0f / sqrtf(var / (float)D + 1e-5f);
    for (int i = threadIdx.x; i < D; i += 256) xs[i] = (xs[i] - mean) * rstd * ln_w[i] + ln_b[i];
    __syncthreads();
    for (int u = blockIdx.x * CLS_UNITS + warp; u < min(Hd, (int)(blockIdx.x + 1) * CLS_UNITS); u += 8) {
        const float* wr = w1 + (long long)u * D;
        float a = 0.f;
        for (int i = lane * 4; i < D; i += 128) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + i));
            const float4 xv = *reinterpret_cast<const float4*>(xs + i);
            a = fmaf(wv.x, xv.x, a); a = fmaf(wv.y, xv.y, a); a = fmaf(wv.z, xv.z, a); a = fmaf(wv.w, xv.w, a);
        }
        a = warp_sum(a);
        if (lane == 0) hidden[(long long)b * Hd + u] = fmaxf(a + b1[u], 0.f);
    }
}
__global__ void classifier_out_kernel(const float* __restrict__ hidden, int Hd, const float* __restrict__ w2, const float* __restrict__ b2,
                                      float* __restrict__ logprob) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const float* hs = hidden + (long long)b * Hd;
    float l0 = 0.f, l1 = 0.f;
    for (int i = lane; i < Hd; i += 32) { l0 = fmaf(w2[i], hs[i], l0); l1 = fmaf(w2[Hd + i], hs[i], l1); }
    l0 = warp_sum(l0) + b2[0]; l1 = warp_sum(l1) + b2[1];
    if (lane == 0) {
        const float m = fmaxf(l0, l1);
        const float lse = m + logf(expf(l0 - m) + expf(l1 - m));
        logprob[b * 2 + 0] = l0 - lse; logprob[b * 2 + 1] = l1 - lse;
    }
}

// ---- SLS
struct LayerPtrs { const void* p[32]; };

// grid (n_layers, B), 256 threads x 4 channels: layer mean over frames, dot with fc0, sigmoid
__global__ void __launch_bounds__(256) sls_weights_kernel(LayerPtrs L, int T, int D, const float* __restrict__ fc0_w, const float* __restrict__ fc0_b,
                                                          float* __restrict__ layer_w, int n_layers, const int* __restrict__ lens) {
    __shared__ float red[8];
    const int l = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int len = lens ? min(lens[b], T) : T;
    const float* x = static_cast<const float*>(L.p[l]) + (long long)b * T * D;
    float dot = 0.f;
    for (int c = threadIdx.x * 4; c < D; c += 1024) {
        float4 s = make_float4(0, 0, 0, 0);
        for (int t = 0; t < len; ++t) {
            const float4 v = *reinterpret_cast<const float4*>(x + (long long)t * D + c);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        const float inv = 1.0f / (float)len;
        const float4 w = *reinterpret_cast<const float4*>(fc0_w + c);
        dot = fmaf(s.x * inv, w.x, dot); dot = fmaf(s.y * inv, w.y, dot); dot = fmaf(s.z * inv, w.z, dot); dot = fmaf(s.w * inv, w.w, dot);
    }
    dot = warp_sum(dot);
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        float z = fc0_b[0];
        for (int w = 0; w < 8; ++w) z += red[w];
        layer_w[b * n_layers + l] = 1.0f / (1.0f + expf(-z));
    }
}

__device__ __forceinline__ float selu(float x) {
    const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
    return scale * (x > 0.f ? x : alpha * (expf(x) - 1.0f));
}

// layer weights from the per-frame fc0 dots the LayerNorm kernels emitted while they advanced the residual stream
// (fc0(mean_t x) == mean_t fc0(x)): grid (n_layers, B), one warp, fixed-order reduction
__global__ void __launch_bounds__(32) sls_weights_from_dots_kernel(const float* __restrict__ dots, long long M, int T, const float* __restrict__ fc0_b,
                                                                   float* __restrict__ layer_w, int n_layers) {
    const int l = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
    const float* d = dots + (long long)l * M + (long long)b * T;
    float s = 0.f;
    for (int t = lane; t < T; t += 32) s += d[t];
    s = warp_sum(s);
    if (lane == 0) layer_w[b * n_layers + l] = 1.0f / (1.0f + expf(-(s / (float)T + fc0_b[0])));
}

// grid (T/3, B), D/4 threads x 4 channels: weighted layer sum for 3 frames (12 independent 16-byte loads in flight per
// layer group), BN (eval affine) + SELU, then 3x3 max pool -> out[b][i*(D/3)+j].  Every layer output is read exactly once.
template <typename TI, typename TO, int VEC>
__global__ void __launch_bounds__(256) sls_fuse_pool_kernel(LayerPtrs L, int n_layers, const float* __restrict__ layer_w, int T, int D,
                                                            const float* __restrict__ bn, float bn_eps, TO* __restrict__ out, int ldo) {
    // grid (T/3, B); every thread owns VEC consecutive channels of the block's 3 frames (8- or 16-byte loads):
    // weighted layer sum, BN (eval affine) + SELU into smem, then the 3x3 max pool -> out[b][i*(D/3)+j].  Each layer is read once.
    static_assert(VEC * sizeof(TI) == 16 || VEC * sizeof(TI) == 8, "8- or 16-byte loads");
    extern __shared__ float fp_smem[];   // [3][D]
    __shared__ float lw[32];
    const int i = blockIdx.x, b = blockIdx.y, c = threadIdx.x * VEC;
    if (threadIdx.x < n_layers) lw[threadIdx.x] = layer_w[b * n_layers + threadIdx.x];
    __syncthreads();
    const float g = bn[0] / sqrtf(bn[3] + bn_eps), beta = bn[1], rm = bn[2];
    if (c < D) {
        const long long off = ((long long)b * T + 3 * i) * D + c;
        float s[3][VEC];
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
            for (int e = 0; e < VEC; ++e) s[di][e] = 0.f;
#pragma unroll 4
        for (int l = 0; l < n_layers; ++l) {
            const float w = lw[l];
            const TI* base = static_cast<const TI*>(L.p[l]) + off;
#pragma unroll
            for (int di = 0; di < 3; ++di) {
                uint32_t t[VEC * sizeof(TI) / 4];
                if constexpr (VEC * sizeof(TI) == 16) {
                    const uint4 q = __ldcs(reinterpret_cast<const uint4*>(base + (long long)di * D));
                    t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
                } else {
                    const uint2 q = __ldcs(reinterpret_cast<const uint2*>(base + (long long)di * D));
                    t[0] = q.x; t[1] = q.y;
                }
                if constexpr (sizeof(TI) == 2) {
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(t);
#pragma unroll
                    for (int e = 0; e < VEC / 2; ++e) {
                        s[di][2 * e] = fmaf(__low2float(h[e]), w, s[di][2 * e]);
                        s[di][2 * e + 1] = fmaf(__high2float(h[e]), w, s[di][2 * e + 1]);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) s[di][e] = fmaf(__uint_as_float(t[e]), w, s[di][e]);
                }
            }
        }
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
            for (int e = 0; e < VEC; ++e) fp_smem[di * D + c + e] = selu((s[di][e] - rm) * g + beta);
    }
    __syncthreads();
    const int J = D / 3;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        float m = -INFINITY;
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
            for (int dj = 0; dj < 3; ++dj) m = fmaxf(m, fp_smem[di * D + 3 * j + dj]);
        out[(long long)b * ldo + i * J + j] = from_f32<TO>(m);
    }
    if (i == (int)gridDim.x - 1)          // zero the K padding of this utterance's row (fc1 reads whole 64-column k-blocks)
        for (int j = (int)gridDim.x * J + threadIdx.x; j < ldo; j += blockDim.x) out[(long long)b * ldo + j] = from_f32<TO>(0.f);
}

// ---- bulk-copy ring version (bf16 layers, D = 1024; default, SLSB_POOL_TMA=0 selects the plain-load kernel above) ----
// validated on B200 in round 2: tests/test_parity_gpu.py (sls, fp32 + bf16) green, 0.135 ms vs 0.165 ms per 64-clip batch
// The kernel above is long-scoreboard bound at 50 % occupancy (ncu: 3.8-4.2 TB/s): its bytes in flight live in registers.
// Here a producer warp streams the 25 per-layer chunks of the block's 3 frames (3 x 2 KB, contiguous) through an 8-stage shared
// memory ring with cp.async.bulk + mbarriers; 4 consumer warps accumulate w_l * x_l from shared memory (8 channels x 3 frames per
// thread).  Same arithmetic order as sls_fuse_pool_kernel<bf16, TO, 8> => identical bits.
constexpr int kPtStages = 8, kPtD = 1024, kPtChunk = 3 * kPtD * 2;
constexpr int kPtSmem = kPtStages * kPtChunk + 3 * kPtD * 4 + 32 * 4 + 2 * kPtStages * 8;

__device__ __forceinline__ void pool_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <typename TO>
__global__ void __launch_bounds__(160) sls_fuse_pool_tma_kernel(LayerPtrs L, int n_layers, const float* __restrict__ layer_w, int T,
                                                                const float* __restrict__ bn, float bn_eps, TO* __restrict__ out, int ldo) {
    constexpr int D = kPtD;
    extern __shared__ __align__(128) uint8_t pt_smem[];
    float* fp = reinterpret_cast<float*>(pt_smem + kPtStages * kPtChunk);          // [3][D] pooled-input tile
    float* lw = fp + 3 * D;                                                         // [32] layer weights of this utterance
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(lw + 32);
    uint64_t* empty_bar = full_bar + kPtStages;
    const int i = blockIdx.x, b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPtStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 4); }
        mbar_fence_init();
    }
    if (threadIdx.x < n_layers) lw[threadIdx.x] = layer_w[b * n_layers + threadIdx.x];
    __syncthreads();
    const long long off = ((long long)b * T + 3 * i) * D;
    if (warp == 4) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int l = 0; l < n_layers; ++l) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], kPtChunk);
                pool_bulk_load(pt_smem + stage * kPtChunk, static_cast<const bf16*>(L.p[l]) + off, kPtChunk, &full_bar[stage]);
                if (++stage == kPtStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int c = threadIdx.x * 8;
        float s[3][8];
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
            for (int e = 0; e < 8; ++e) s[di][e] = 0.f;
        int stage = 0; uint32_t phase = 0;
        for (int l = 0; l < n_layers; ++l) {
            const float w = lw[l];
            mbar_wait(&full_bar[stage], phase);
            const uint8_t* chunk = pt_smem + stage * kPtChunk;
#pragma unroll
            for (int di = 0; di < 3; ++di) {
                const uint4 q = *reinterpret_cast<const uint4*>(chunk + (di * D + c) * 2);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    s[di][2 * e] = fmaf(__low2float(h[e]), w, s[di][2 * e]);
                    s[di][2 * e + 1] = fmaf(__high2float(h[e]), w, s[di][2 * e + 1]);
                }
            }
            // release the chunk only after its three loads have returned (see mbar_arrive_after): the value depends on all of them
            const float dep = (s[0][0] + s[1][0]) + s[2][0];
            __syncwarp();
            if (lane == 0) mbar_arrive_after(&empty_bar[stage], dep);
            if (++stage == kPtStages) { stage = 0; phase ^= 1; }
        }
        const float g = bn[0] / sqrtf(bn[3] + bn_eps), beta = bn[1], rm = bn[2];
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
            for (int e = 0; e < 8; ++e) fp[di * D + c + e] = selu((s[di][e] - rm) * g + beta);
    }
    __syncthreads();
    const int J = D / 3;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        float m = -INFINITY;
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
            for (int dj = 0; dj < 3; ++dj) m = fmaxf(m, fp[di * D + 3 * j + dj]);
        out[(long long)b * ldo + i * J + j] = from_f32<TO>(m);
    }
    if (i == (int)gridDim.x - 1)
        for (int j = (int)gridDim.x * J + threadIdx.x; j < ldo; j += blockDim.x) out[(long long)b * ldo + j] = from_f32<TO>(0.f);
}

// partial sums [B][KS][N] -> h = selu(sum + bias) -> fc3 -> selu -> log_softmax ; one block of 1024 threads per utterance
// (one hidden unit per thread: the KS partials of a unit are read by consecutive threads -> coalesced rows of the partial matrix)
__global__ void __launch_bounds__(1024) sls_tail_kernel(const float* __restrict__ partial, int KS, int Hd, const float* __restrict__ b1,
                                                        const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ logprob) {
    __shared__ float red0[32], red1[32];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float l0 = 0.f, l1 = 0.f;
    for (int n = threadIdx.x; n < Hd; n += blockDim.x) {
        float s = 0.f;
        for (int ks = 0; ks < KS; ++ks) s += partial[((long long)b * KS + ks) * Hd + n];      // fixed order: bit-stable
        const float h = selu(s + b1[n]);
        l0 = fmaf(w3[n], h, l0); l1 = fmaf(w3[Hd + n], h, l1);
    }
    l0 = warp_sum(l0); l1 = warp_sum(l1);
    if (lane == 0) { red0[warp] = l0; red1[warp] = l1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = b3[0], c = b3[1];
        for (int w = 0; w < nwarps; ++w) { a += red0[w]; c += red1[w]; }
        a = selu(a); c = selu(c);
        const float m = fmaxf(a, c);
        const float lse = m + logf(expf(a - m) + expf(c - m));
        logprob[b * 2] = a - lse; logprob[b * 2 + 1] = c - lse;
    }
}

__global__ void scores_kernel(const float* __restrict__ logprob, float* __restrict__ scores, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) scores[b] = expf(logprob[b * 2 + 1]);
}

__global__ void __launch_bounds__(256) sqdiff_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ partial) {
    __shared__ float red[8];
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) { const float d = a[i] - b[i]; s = fmaf(d, d, s); }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; partial[blockIdx.x] = t; }
}
__global__ void sqdiff_final_kernel(const float* __restrict__ partial, int np, long long n, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { double t = 0.0; for (int i = 0; i < np; ++i) t += partial[i]; out[0] = (float)(t / (double)n); }
}

}  // namespace

int topk_threshold(const float* x, long long rows, int D, int k, float* thr, int* tie_cut, cudaStream_t stream) {
    if (rows <= 0) return 0;
    if (k < 1 || k > D) { set_error("topk: k=%d out of range for D=%d", k, D); return -1; }
    switch (D) {
        case 4096: topk_threshold_kernel<16><<<(unsigned)rows, 256, 0, stream>>>(x, D, k, thr, tie_cut); break;
        case 2048: topk_threshold_kernel<8><<<(unsigned)rows, 256, 0, stream>>>(x, D, k, thr, tie_cut); break;
        case 1024: topk_threshold_kernel<4><<<(unsigned)rows, 256, 0, stream>>>(x, D, k, thr, tie_cut); break;
        case 8192: topk_threshold_kernel<32><<<(unsigned)rows, 256, 0, stream>>>(x, D, k, thr, tie_cut); break;
        default: set_error("topk: dict size %d unsupported (1024/2048/4096/8192)", D); return -1;
    }
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int topk_densify(const float* acts, const float* thr, const int* tie_cut, float* encoded, long long rows, int D, cudaStream_t stream) {
    return votes_densify(acts, acts, thr, tie_cut, encoded, rows, D, stream);
}
int votes_densify(const float* acts, const float* votes, const float* thr, const int* tie_cut, float* encoded, long long rows, int D, cudaStream_t stream) {
    if (rows <= 0) return 0;
    const long long n4 = rows * D / 4;
    densify_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(acts, votes, thr, tie_cut, encoded, rows, D);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int votes_compact(const float* acts, const float* votes, const float* thr, const int* tie_cut, int* idx_out, float* val_out, int* count_out,
                  long long rows, int D, int k, cudaStream_t stream) {
    if (rows <= 0) return 0;
    if (D % 32 != 0) { set_error("votes_compact: D=%d must be a multiple of 32", D); return -1; }
    compact_kept_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(acts, votes, thr, tie_cut, idx_out, val_out, count_out, rows, D, k);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int topk_mean_pool(const float* acts, const float* thr, const int* tie_cut, float* pooled, int B, int T, int D, const int* lens, cudaStream_t stream) {
    return votes_mean_pool(acts, acts, thr, tie_cut, pooled, B, T, D, lens, stream);
}
int votes_mean_pool(const float* acts, const float* votes, const float* thr, const int* tie_cut, float* pooled, int B, int T, int D, const int* lens, cudaStream_t stream) {
    dim3 grid((D + 127) / 128, B);
    mean_pool_kept_kernel<<<grid, 128, 0, stream>>>(acts, votes, thr, tie_cut, pooled, T, D, lens);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int sel_chunks(int T) { return (T + kSelRows - 1) / kSelRows; }

#define SLSB_EPT_DISPATCH(D_, CALL)                                                                    \
    switch (D_) {                                                                                      \
        case 1024: { constexpr int EPT = 4; CALL; break; }                                             \
        case 2048: { constexpr int EPT = 8; CALL; break; }                                             \
        case 4096: { constexpr int EPT = 16; CALL; break; }                                            \
        case 8192: { constexpr int EPT = 32; CALL; break; }                                            \
        default: set_error("top-k: dict size %d unsupported (1024/2048/4096/8192)", D_); return -1;    \
    }

int topk_select_pool(const float* acts, int B, int T, int D, int k, const int* lens, float* thr, int* tie_cut, float* partial, float* pooled,
                     cudaStream_t stream) {
    if (B <= 0 || T <= 0) return 0;
    if (k < 1 || k > D) { set_error("topk: need 1 <= k <= D (k=%d, D=%d)", k, D); return -1; }
    if ((partial == nullptr) != (pooled == nullptr)) { set_error("topk_select_pool: partial and pooled go together"); return -1; }
    const int nc = sel_chunks(T);
    dim3 grid(nc, B);
    SLSB_EPT_DISPATCH(D, (topk_pool_kernel<EPT><<<grid, 256, 0, stream>>>(acts, T, k, lens, thr, tie_cut, partial)));
    if (pooled) pool_finish_kernel<<<dim3((D + 255) / 256, B), 256, 0, stream>>>(partial, nc, T, D, lens, pooled);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int window_select_pool(const float* acts, int B, int T, int D, int k, int window, int stride, int nw, uint32_t* wmask, float* thr, int* tie_cut,
                       float* votes_or_null, float* partial, float* pooled, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return 0;
    if (k < 1 || k > D) { set_error("topk: need 1 <= k <= D (k=%d, D=%d)", k, D); return -1; }
    if ((partial == nullptr) != (pooled == nullptr)) { set_error("window_select_pool: partial and pooled go together"); return -1; }
    const int nc = sel_chunks(T);
    SLSB_EPT_DISPATCH(D, (window_select_kernel<EPT><<<dim3(nw, B), 256, 0, stream>>>(acts, T, k, window, stride, nw, wmask)));
    SLSB_EPT_DISPATCH(D, (window_vote_pool_kernel<EPT><<<dim3(nc, B), 256, 0, stream>>>(acts, wmask, T, k, window, stride, nw, thr, tie_cut,
                                                                                         votes_or_null, partial)));
    if (pooled) pool_finish_kernel<<<dim3((D + 255) / 256, B), 256, 0, stream>>>(partial, nc, T, D, nullptr, pooled);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int mean_pool_frames(const float* x, float* pooled, int B, int T, int D, const int* lens, cudaStream_t stream) {
    dim3 grid((D + 127) / 128, B);
    mean_pool_kernel<<<grid, 128, 0, stream>>>(x, pooled, T, D, lens);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int window_sums(const float* acts, float* sums, int B, int T, int D, int window, int stride, int nw, cudaStream_t stream) {
    dim3 grid((D + 255) / 256, nw, B);
    window_sums_kernel<<<grid, 256, 0, stream>>>(acts, sums, T, D, window, stride, nw);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int window_votes(const float* acts, const float* sums, const float* thr_w, const int* cut_w, float* votes,
                 int B, int T, int D, int window, int stride, int nw, cudaStream_t stream) {
    dim3 grid((D + 255) / 256, T, B);
    window_votes_kernel<<<grid, 256, 0, stream>>>(acts, sums, thr_w, cut_w, votes, T, D, window, stride, nw);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int classifier_head(const float* pooled, int B, int D, int Hd, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* hidden, float* logprob, cudaStream_t stream) {
    if (D % 128 != 0) { set_error("classifier: D=%d must be a multiple of 128", D); return -1; }
    dim3 grid((Hd + CLS_UNITS - 1) / CLS_UNITS, B);
    classifier_hidden_kernel<<<grid, 256, D * sizeof(float), stream>>>(pooled, D, Hd, ln_w, ln_b, w1, b1, hidden);
    classifier_out_kernel<<<B, 32, 0, stream>>>(hidden, Hd, w2, b2, logprob);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int sls_layer_weights(const float* const* layers, int n_layers, int B, int T, int D, const float* fc0_w, const float* fc0_b,
                      float* layer_w, const int* lens, cudaStream_t stream) {
    if (n_layers > 32) { set_error("sls: at most 32 layers"); return -1; }
    LayerPtrs L{};
    for (int i = 0; i < n_layers; ++i) L.p[i] = layers[i];
    dim3 grid(n_layers, B);
    sls_weights_kernel<<<grid, 256, 0, stream>>>(L, T, D, fc0_w, fc0_b, layer_w, n_layers, lens);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int sls_layer_weights_from_dots(const float* dots, int n_layers, int B, int T, const float* fc0_b, float* layer_w, cudaStream_t stream) {
    dim3 grid(n_layers, B);
    sls_weights_from_dots_kernel<<<grid, 32, 0, stream>>>(dots, (long long)B * T, T, fc0_b, layer_w, n_layers);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int sls_fuse_pool(const void* const* layers, int layers_bf16, int n_layers, const float* layer_w, int B, int T, int D, const float* bn, float bn_eps,
                  void* out, int out_bf16, int ldo, cudaStream_t stream) {
    if (n_layers > 32 || D > 1024) { set_error("sls_fuse_pool: n_layers<=32, D<=1024"); return -1; }
    LayerPtrs L{};
    for (int i = 0; i < n_layers; ++i) L.p[i] = layers[i];
    if (D % 8 != 0) { set_error("sls_fuse_pool: D must be a multiple of 8"); return -1; }
    dim3 grid(T / 3, B);
    const size_t sm = 3 * D * sizeof(float);
    // bf16 layers: 4 channels per thread (8-byte loads, D / 4 threads) measured faster in the step than 8 per thread (0.14 vs 0.17 ms)
    static const int vec8 = getenv("SLSB_POOL_VEC") ? atoi(getenv("SLSB_POOL_VEC")) == 8 : 0;
    static const int pool_tma = getenv("SLSB_POOL_TMA") ? atoi(getenv("SLSB_POOL_TMA")) : 1;   // default since round 2: 0.135 vs 0.165 ms (0.72 vs 0.59 of the copy bandwidth), parity green
    if (layers_bf16 && pool_tma && D == kPtD) {
        static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
        if (first_use_on_device(&configured_on)) {
            SLSB_CUDA_CHECK(cudaFuncSetAttribute(sls_fuse_pool_tma_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPtSmem));
            SLSB_CUDA_CHECK(cudaFuncSetAttribute(sls_fuse_pool_tma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPtSmem));
        }
        if (out_bf16) sls_fuse_pool_tma_kernel<bf16><<<grid, 160, kPtSmem, stream>>>(L, n_layers, layer_w, T, bn, bn_eps, static_cast<bf16*>(out), ldo);
        else sls_fuse_pool_tma_kernel<float><<<grid, 160, kPtSmem, stream>>>(L, n_layers, layer_w, T, bn, bn_eps, static_cast<float*>(out), ldo);
    } else if (layers_bf16 && vec8) {
        const int nt = D / 8 > 32 ? D / 8 : 32;
        if (out_bf16) sls_fuse_pool_kernel<bf16, bf16, 8><<<grid, nt, sm, stream>>>(L, n_layers, layer_w, T, D, bn, bn_eps, static_cast<bf16*>(out), ldo);
        else sls_fuse_pool_kernel<bf16, float, 8><<<grid, nt, sm, stream>>>(L, n_layers, layer_w, T, D, bn, bn_eps, static_cast<float*>(out), ldo);
    } else if (layers_bf16) {
        if (out_bf16) sls_fuse_pool_kernel<bf16, bf16, 4><<<grid, 256, sm, stream>>>(L, n_layers, layer_w, T, D, bn, bn_eps, static_cast<bf16*>(out), ldo);
        else sls_fuse_pool_kernel<bf16, float, 4><<<grid, 256, sm, stream>>>(L, n_layers, layer_w, T, D, bn, bn_eps, static_cast<float*>(out), ldo);
    } else if (out_bf16) sls_fuse_pool_kernel<float, bf16, 4><<<grid, 256, sm, stream>>>(L, n_layers, layer_w, T, D, bn, bn_eps, static_cast<bf16*>(out), ldo);
    else sls_fuse_pool_kernel<float, float, 4><<<grid, 256, sm, stream>>>(L, n_layers, layer_w, T, D, bn, bn_eps, static_cast<float*>(out), ldo);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int sls_tail(const float* partial, int KS, int B, int Hd, const float* b1, const float* w3, const float* b3, float* logprob, cudaStream_t stream) {
    sls_tail_kernel<<<B, 1024, 0, stream>>>(partial, KS, Hd, b1, w3, b3, logprob);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int scores_from_logprob(const float* logprob, float* scores, int B, cudaStream_t stream) {
    scores_kernel<<<(B + 127) / 128, 128, 0, stream>>>(logprob, scores, B);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
int mse_loss(const float* a, const float* b, long long n, float* out_scalar, float* scratch, cudaStream_t stream) {
    const int np = 1024;
    sqdiff_partial_kernel<<<np, 256, 0, stream>>>(a, b, n, scratch);
    sqdiff_final_kernel<<<1, 32, 0, stream>>>(scratch, np, n, out_scalar);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace slsb

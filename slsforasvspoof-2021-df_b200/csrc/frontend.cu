// HBM-bound front-end kernels: conv0 + LayerNorm + GELU (fused), row LayerNorm (+GELU, + centred copy), frame padding
// for the positional conv, dtype conversion, synthetic-clip generation and the conv length formula.
// All are coalesced / vectorised, one warp per row with shuffle reductions, fp32 statistics.
#include "common.cuh"
#include "kernels.h"
#include <math.h>

namespace slsb {
namespace {

// ------------------------------------------------------------------------------------------------
// conv0: Conv1d(1 -> 512, k = 10, stride 5, bias) + Fp32LayerNorm(512) + GELU   (wav2vec2.py:795-813, first block)
// One warp per output frame; each lane owns 16 channels (c = 64 j + 2 lane + e) whose 10 taps live in registers,
// so a frame costs 10 broadcast loads + 160 FFMA + two shuffle reductions and one coalesced 1 KB / 2 KB store.
// ------------------------------------------------------------------------------------------------
constexpr int C0 = 512, K0 = 10, FRAMES_PER_WARP = 16;

template <typename TO, bool EXACT>
__global__ void __launch_bounds__(256) conv0_kernel(const float* __restrict__ wav, int S, int L0, int stride,
                                                    const float* __restrict__ w, const float* __restrict__ bias,
                                                    const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                    TO* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int f0 = (blockIdx.x * 8 + warp) * FRAMES_PER_WARP;
    if (f0 >= L0) return;
    float wr[8][2][K0], br[8][2], gw[8][2], gb[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = 64 * j + 2 * lane + e;
#pragma unroll
            for (int t = 0; t < K0; ++t) wr[j][e][t] = __ldg(w + c * K0 + t);
            br[j][e] = __ldg(bias + c); gw[j][e] = __ldg(ln_w + c); gb[j][e] = __ldg(ln_b + c);
        }
    const float* xb = wav + (long long)b * S;
    const int f1 = min(f0 + FRAMES_PER_WARP, L0);
    for (int f = f0; f < f1; ++f) {
        float x[K0];
#pragma unroll
        for (int t = 0; t < K0; ++t) x[t] = __ldg(xb + f * stride + t);
        float v[8][2];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float a = 0.f;
#pragma unroll
                for (int t = 0; t < K0; ++t) a = fmaf(wr[j][e][t], x[t], a);
                a += br[j][e];
                v[j][e] = a; s += a;
            }
        const float mean = warp_sum(s) * (1.0f / C0);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) { const float d = v[j][e] - mean; q = fmaf(d, d, q); }
        const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C0) + 1e-5f);
        TO* o = out + ((long long)b * L0 + f) * C0 + 2 * lane;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float y0 = gelu<EXACT>((v[j][0] - mean) * rstd * gw[j][0] + gb[j][0]);
            const float y1 = gelu<EXACT>((v[j][1] - mean) * rstd * gw[j][1] + gb[j][1]);
            if constexpr (sizeof(TO) == 2) *reinterpret_cast<uint32_t*>(o + 64 * j) = pack_bf16x2(y0, y1);
            else *reinterpret_cast<float2*>(o + 64 * j) = make_float2(y0, y1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// row LayerNorm, one warp per row, NV = C / 128 vectors of 4 per lane
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load4<bf16>(const bf16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x), b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, const float (&v)[4]) {
    uint2 t; t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
}

template <typename TI, typename TO, typename TO2, int NV, bool GELU, bool EXACT, bool DOT>
__global__ void __launch_bounds__(256, 4) ln_kernel(const TI* __restrict__ in, const bf16* __restrict__ add, float* __restrict__ sum_out,
                                                 TO* __restrict__ out, TO2* __restrict__ out2,
                                                 const float* __restrict__ sub, const float* __restrict__ w, const float* __restrict__ b,
                                                 const float* __restrict__ dot_w, float* __restrict__ dot_out, bf16* __restrict__ copy_out,
                                                 long long rows, float eps) {
    constexpr int C = NV * 128;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    griddep_launch();
    griddep_wait();
    if (row >= rows) return;
    const TI* x = in + row * C;
    float v[NV][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) load4<TI>(x + (i * 32 + lane) * 4, v[i]);      // all loads of the row in flight before the first use
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (add) {      // residual stream + bf16 branch output (out_proj / fc2), summed in fp32 and optionally written back
            float y[4];
            load4<bf16>(add + row * C + (i * 32 + lane) * 4, y);
            v[i][0] += y[0]; v[i][1] += y[1]; v[i][2] += y[2]; v[i][3] += y[3];
            if (sum_out) store4<float>(sum_out + row * C + (i * 32 + lane) * 4, v[i]);
        }
        if (copy_out) store4<bf16>(copy_out + row * C + (i * 32 + lane) * 4, v[i]);     // bf16 snapshot of the (pre-norm) stream row
        if constexpr (DOT) {   // SLS layer weighting: fc0 . x_row of the (pre-norm) residual stream
            const float4 dw = __ldg(reinterpret_cast<const float4*>(dot_w + (i * 32 + lane) * 4));
            dot = fmaf(v[i][0], dw.x, dot); dot = fmaf(v[i][1], dw.y, dot); dot = fmaf(v[i][2], dw.z, dot); dot = fmaf(v[i][3], dw.w, dot);
        }
        s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
    if constexpr (DOT) {
        dot = warp_sum(dot);                       // fixed shuffle tree: bit-stable
        if (lane == 0) dot_out[row] = dot;
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float d = v[i][e] - mean; q = fmaf(d, d, q); }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(w + c)), bb = __ldg(reinterpret_cast<const float4*>(b + c));
        float y[4];
        y[0] = (v[i][0] - mean) * rstd * g.x + bb.x; y[1] = (v[i][1] - mean) * rstd * g.y + bb.y;
        y[2] = (v[i][2] - mean) * rstd * g.z + bb.z; y[3] = (v[i][3] - mean) * rstd * g.w + bb.w;
        if constexpr (GELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) y[e] = gelu<EXACT>(y[e]);
        }
        if (out) store4<TO>(out + row * C + c, y);
        if (out2) {
            const float4 sv = __ldg(reinterpret_cast<const float4*>(sub + c));
            float z[4] = {y[0] - sv.x, y[1] - sv.y, y[2] - sv.z, y[3] - sv.w};
            store4<TO2>(out2 + row * C + c, z);
        }
    }
}

template <typename TI, typename TO, typename TO2, int NV>
int ln_launch(const LnArgs& a, cudaStream_t stream) {
    const unsigned grid = (unsigned)((a.rows + 7) / 8);
    const TI* in = static_cast<const TI*>(a.in); TO* out = static_cast<TO*>(a.out); TO2* out2 = static_cast<TO2*>(a.out2);
    const bf16* add = static_cast<const bf16*>(a.add);
    if (a.gelu) {
        if (a.dot_out) { set_error("layernorm: dot_out is not combined with gelu"); return -1; }
        if (a.exact_gelu) SLSB_CUDA_CHECK(launch_pdl(ln_kernel<TI, TO, TO2, NV, true, true, false>, dim3(grid), dim3(256), 0, stream, in, add, a.sum_out, out, out2, a.sub, a.w, a.b, a.dot_w, a.dot_out, static_cast<bf16*>(a.copy_out), a.rows, a.eps));
        else SLSB_CUDA_CHECK(launch_pdl(ln_kernel<TI, TO, TO2, NV, true, false, false>, dim3(grid), dim3(256), 0, stream, in, add, a.sum_out, out, out2, a.sub, a.w, a.b, a.dot_w, a.dot_out, static_cast<bf16*>(a.copy_out), a.rows, a.eps));
    } else if (a.dot_out) {
        SLSB_CUDA_CHECK(launch_pdl(ln_kernel<TI, TO, TO2, NV, false, true, true>, dim3(grid), dim3(256), 0, stream, in, add, a.sum_out, out, out2, a.sub, a.w, a.b, a.dot_w, a.dot_out, static_cast<bf16*>(a.copy_out), a.rows, a.eps));
    } else {
        SLSB_CUDA_CHECK(launch_pdl(ln_kernel<TI, TO, TO2, NV, false, true, false>, dim3(grid), dim3(256), 0, stream, in, add, a.sum_out, out, out2, a.sub, a.w, a.b, a.dot_w, a.dot_out, static_cast<bf16*>(a.copy_out), a.rows, a.eps));
    }
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
// ------------------------------------------------------------------------------------------------
// Encoder LayerNorm(1024) of the fp32 residual stream -> bf16, persistent and fed by bulk async copies.
// The one-warp-per-row kernel above is latency-bound (ncu: 0.3 instructions / clk / scheduler, 33-40 % of the DRAM peak): a
// warp walks ~340 dependent-ish instructions per row and the register file caps the SM at 32 such warps whose loads are all
// issued wave by wave.  Here a producer warp streams 8-row blocks (32 KB, contiguous) into a 3-stage shared memory ring with
// cp.async.bulk + mbarrier transaction counts, and TWO consumer warps share a row (512 channels each, 16 values per lane):
// half the dependent chain per warp, 32 consumer warps per SM in 56 registers, partial sums exchanged through shared memory
// with one 64-thread named barrier per row.  2 CTAs / SM x 96 KB of ring = up to 192 KB of loads in flight per SM.
// ------------------------------------------------------------------------------------------------
constexpr int kLsRows = 8, kLsC = 1024, kLsStages = 3, kLsStageBytes = kLsRows * kLsC * 4;
constexpr int kLsConsumers = 2 * kLsRows, kLsThreads = (kLsConsumers + 1) * 32;
constexpr int kLsSmem = kLsStages * kLsStageBytes + 2 * kLsC * 4 + kLsRows * 2 * 16 + 64;     // ring + LN weight / bias + exchange + barriers

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool DOT>
__global__ void __launch_bounds__(kLsThreads, 2) ln_stream_kernel(const float* __restrict__ in, bf16* __restrict__ out,
                                                                  const float* __restrict__ w, const float* __restrict__ b,
                                                                  const float* __restrict__ dot_w, float* __restrict__ dot_out,
                                                                  bf16* __restrict__ copy_out, long long rows, float eps) {
    constexpr int C = kLsC, NVH = C / 256;                              // float4 vectors per lane (half a row per warp)
    extern __shared__ __align__(128) uint8_t ls_smem[];
    float* w_s = reinterpret_cast<float*>(ls_smem + kLsStages * kLsStageBytes);
    float* b_s = w_s + C;
    float4* xch = reinterpret_cast<float4*>(b_s + C);                   // [row][half] = {sum, dot, sum of squared deviations, -}
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(xch + kLsRows * 2);
    uint64_t* empty_bar = full_bar + kLsStages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long num_blocks = (rows + kLsRows - 1) / kLsRows;
    griddep_launch();
    if (threadIdx.x == 0) {
        for (int s = 0; s < kLsStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kLsConsumers); }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < C; i += blockDim.x) { w_s[i] = __ldg(w + i); b_s[i] = __ldg(b + i); }   // parameters: not written by the previous kernel
    __syncthreads();
    griddep_wait();
    if (warp == kLsConsumers) {
        // ===================== producer: one bulk copy per 8-row block =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long blk = blockIdx.x; blk < num_blocks; blk += gridDim.x) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const long long r0 = blk * kLsRows;
                const uint32_t bytes = (uint32_t)((rows - r0 < kLsRows ? rows - r0 : kLsRows) * C * 4);
                mbar_expect_tx(&full_bar[stage], bytes);
                bulk_load_1d(ls_smem + stage * kLsStageBytes, in + r0 * C, bytes, &full_bar[stage]);
                if (++stage == kLsStages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }
    // ===================== consumers: warps 2r, 2r + 1 = halves of row r of the block =====================
    const int r = warp >> 1, h = warp & 1;
    const int c0 = h * (C / 2) + lane * 4;                               // this lane's channels: c0 + 128 i, i < NVH
    const int bar_id = 1 + r;
    int stage = 0; uint32_t phase = 0;
    for (long long blk = blockIdx.x; blk < num_blocks; blk += gridDim.x) {
        const long long row = blk * kLsRows + r;
        mbar_wait(&full_bar[stage], phase);
        float v[NVH][4];
        if (row < rows) {
            const float* x = reinterpret_cast<const float*>(ls_smem + stage * kLsStageBytes) + r * C + c0;
#pragma unroll
            for (int i = 0; i < NVH; ++i) load4<float>(x + 128 * i, v[i]);
        }
        uint64_t* release = &empty_bar[stage];
        if (++stage == kLsStages) { stage = 0; phase ^= 1; }
        if (row >= rows) {                                              // both warps of the row skip together; nothing was read
            __syncwarp();
            if (lane == 0) mbar_arrive(release);
            continue;
        }
        float s = 0.f, dot = 0.f;
#pragma unroll
        for (int i = 0; i < NVH; ++i) {
            if (copy_out) store4<bf16>(copy_out + row * C + c0 + 128 * i, v[i]);
            if constexpr (DOT) {
                const float4 dw = __ldg(reinterpret_cast<const float4*>(dot_w + c0 + 128 * i));
                dot = fmaf(v[i][0], dw.x, dot); dot = fmaf(v[i][1], dw.y, dot); dot = fmaf(v[i][2], dw.z, dot); dot = fmaf(v[i][3], dw.w, dot);
            }
            s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
        }
        s = warp_sum(s);                                                // fixed shuffle trees + fixed half order: bit-stable
        if constexpr (DOT) dot = warp_sum(dot);
        // the stage is released only now: the reduction above consumed every value the warp loaded, so no LDS of the stage is in
        // flight any more (an arrive issued right after the loads let the next bulk copy overwrite bytes that had not been read yet)
        if (lane == 0) mbar_arrive_after(release, s);
        if (lane == 0) { xch[r * 2 + h].x = s; xch[r * 2 + h].y = dot; }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const float mean = (xch[r * 2].x + xch[r * 2 + 1].x) * (1.0f / C);
        if constexpr (DOT) {
            if (h == 0 && lane == 0) dot_out[row] = xch[r * 2].y + xch[r * 2 + 1].y;
        }
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NVH; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float d = v[i][e] - mean; q = fmaf(d, d, q); }
        q = warp_sum(q);
        if (lane == 0) xch[r * 2 + h].z = q;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const float rstd = 1.0f / sqrtf((xch[r * 2].z + xch[r * 2 + 1].z) * (1.0f / C) + eps);
#pragma unroll
        for (int i = 0; i < NVH; ++i) {
            const float4 g = *reinterpret_cast<const float4*>(w_s + c0 + 128 * i), bb = *reinterpret_cast<const float4*>(b_s + c0 + 128 * i);
            float y[4];
            y[0] = (v[i][0] - mean) * rstd * g.x + bb.x; y[1] = (v[i][1] - mean) * rstd * g.y + bb.y;
            y[2] = (v[i][2] - mean) * rstd * g.z + bb.z; y[3] = (v[i][3] - mean) * rstd * g.w + bb.w;
            store4<bf16>(out + row * C + c0 + 128 * i, y);
        }
    }
}

static bool ln_stream_enabled() {
    static const bool on = !(getenv("SLSB_LN_STREAM") && atoi(getenv("SLSB_LN_STREAM")) == 0);
    return on;
}

static int ln_stream_launch(const LnArgs& a, cudaStream_t stream) {
    static int num_sms = 0;
    static unsigned long long configured_on = 0;       // bit d: function attributes set on device d (they are per device)
    if (first_use_on_device(&configured_on)) {
        int dev = 0;
        SLSB_CUDA_CHECK(cudaGetDevice(&dev));
        SLSB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(ln_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLsSmem));
        SLSB_CUDA_CHECK(cudaFuncSetAttribute(ln_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLsSmem));
    }
    const long long blocks = (a.rows + kLsRows - 1) / kLsRows;
    const unsigned grid = (unsigned)(blocks < 2ll * num_sms ? blocks : 2ll * num_sms);
    const float* in = static_cast<const float*>(a.in);
    if (a.dot_out) SLSB_CUDA_CHECK(launch_pdl(ln_stream_kernel<true>, dim3(grid), dim3(kLsThreads), (size_t)kLsSmem, stream, in, static_cast<bf16*>(a.out), a.w, a.b, a.dot_w, a.dot_out, static_cast<bf16*>(a.copy_out), a.rows, a.eps));
    else SLSB_CUDA_CHECK(launch_pdl(ln_stream_kernel<false>, dim3(grid), dim3(kLsThreads), (size_t)kLsSmem, stream, in, static_cast<bf16*>(a.out), a.w, a.b, a.dot_w, a.dot_out, static_cast<bf16*>(a.copy_out), a.rows, a.eps));
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <typename TI, typename TO, typename TO2>
int ln_dispatch_c(const LnArgs& a, cudaStream_t stream) {
    if (a.C == 512) return ln_launch<TI, TO, TO2, 4>(a, stream);
    if (a.C == 1024) return ln_launch<TI, TO, TO2, 8>(a, stream);
    set_error("layernorm: C=%d unsupported (512 or 1024)", a.C);
    return -1;
}

template <typename TO>
__global__ void pad_frames_kernel(const float* __restrict__ x, TO* __restrict__ out, int T, int D, int left, int Tp, const int* __restrict__ lens) {
    const int b = blockIdx.z, j = blockIdx.y;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (c >= D) return;
    const int t = j - left;
    const int len = lens ? lens[b] : T;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (t >= 0 && t < T && t < len) load4<float>(x + ((long long)b * T + t) * D + c, v);
    store4<TO>(out + ((long long)b * Tp + j) * D + c, v);
}

__global__ void zero_padded_kernel(float* __restrict__ x, int T, int D, const int* __restrict__ lens) {
    const int b = blockIdx.z, t = blockIdx.y;
    if (t < lens[b]) return;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (c < D) *reinterpret_cast<float4*>(x + ((long long)b * T + t) * D + c) = make_float4(0, 0, 0, 0);
}

__global__ void cvt_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const long long step = (long long)gridDim.x * blockDim.x * 4;
    for (; i + 3 < n; i += step) {
        float v[4];
        load4<float>(in + i, v);
        store4<bf16>(out + i, v);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long t = n & ~3ll; t < n; ++t) out[t] = __float2bfloat16_rn(in[t]);
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void synth_kernel(float* __restrict__ out, long long first_utt, int samples, float scale) {
    const int u = blockIdx.y;
    const unsigned long long key = splitmix64((unsigned long long)(0x5EED0000ll + first_utt + u) * 0xD6E8FEB86659FD93ull);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < samples; i += gridDim.x * blockDim.x) {
        const unsigned long long h = splitmix64((unsigned long long)i ^ key);
        const unsigned s = (unsigned)(h & 0xFFFF) + (unsigned)((h >> 16) & 0xFFFF) + (unsigned)((h >> 32) & 0xFFFF) + (unsigned)(h >> 48);
        out[(long long)u * samples + i] = __fmul_rn(__fsub_rn(__uint2float_rn(s), 131070.0f), scale);
    }
}

// 16-bit PCM -> fp32 / 32768 with the reference's pad(): first S samples, or tile-repeat a shorter clip up to S
// (data_utils_SSL.py:58-65, :109-115).  grid (ceil(S / 1024), B); HBM-bound: 2 B read + 4 B written per sample.
__global__ void __launch_bounds__(256) ingest_pcm16_kernel(const int16_t* __restrict__ pcm, const long long* __restrict__ offsets,
                                                           const int* __restrict__ lens, float* __restrict__ out, int S) {
    const int b = blockIdx.y;
    const int16_t* src = pcm + offsets[b];
    const int len = lens[b];
    const int i0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 >= S) return;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = i0 + e;
        const int j = len >= S ? i : i % len;
        v[e] = i < S ? __int2float_rn((int)src[j]) * (1.0f / 32768.0f) : 0.f;
    }
    float* o = out + (long long)b * S + i0;
    if (i0 + 3 < S && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    else for (int e = 0; e < 4 && i0 + e < S; ++e) o[e] = v[e];
}

struct ConvCfg { int n; int k[8]; int s[8]; };
__global__ void frame_len_kernel(const int* __restrict__ sl, int* __restrict__ fl, int B, ConvCfg c) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int n = sl[b];
    for (int i = 0; i < c.n; ++i) n = n < c.k[i] ? 0 : (n - c.k[i]) / c.s[i] + 1;   // floor((n - k) / s) + 1, 0 below the kernel width
    fl[b] = n;
}

}  // namespace

int conv0_ln_gelu(const float* wav, int B, int S, int L0, int k, int stride, const float* w, const float* bias,
                  const float* ln_w, const float* ln_b, void* out, int out_bf16, int C, bool exact_gelu, cudaStream_t stream) {
    if (C != C0 || k != K0) { set_error("conv0: only C=512,k=10 supported (got C=%d k=%d)", C, k); return -1; }
    dim3 grid((L0 + 8 * FRAMES_PER_WARP - 1) / (8 * FRAMES_PER_WARP), B);
    if (out_bf16) {
        if (exact_gelu) conv0_kernel<bf16, true><<<grid, 256, 0, stream>>>(wav, S, L0, stride, w, bias, ln_w, ln_b, static_cast<bf16*>(out));
        else conv0_kernel<bf16, false><<<grid, 256, 0, stream>>>(wav, S, L0, stride, w, bias, ln_w, ln_b, static_cast<bf16*>(out));
    } else {
        if (exact_gelu) conv0_kernel<float, true><<<grid, 256, 0, stream>>>(wav, S, L0, stride, w, bias, ln_w, ln_b, static_cast<float*>(out));
        else conv0_kernel<float, false><<<grid, 256, 0, stream>>>(wav, S, L0, stride, w, bias, ln_w, ln_b, static_cast<float*>(out));
    }
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int layernorm(const LnArgs& a, cudaStream_t stream) {
    if (a.rows <= 0) return 0;
    if (!a.in_bf16 && a.out_bf16 && a.out && !a.out2 && !a.add && !a.gelu && a.C == kLsC && a.rows >= 64 && ln_stream_enabled() &&
        (reinterpret_cast<uintptr_t>(a.in) & 15) == 0)
        return ln_stream_launch(a, stream);
    const int sel = a.in_bf16 * 4 + a.out_bf16 * 2 + a.out2_bf16;
    switch (sel) {
        case 0: return ln_dispatch_c<float, float, float>(a, stream);
        case 1: return ln_dispatch_c<float, float, bf16>(a, stream);
        case 2: return ln_dispatch_c<float, bf16, float>(a, stream);
        case 3: return ln_dispatch_c<float, bf16, bf16>(a, stream);
        case 6: return ln_dispatch_c<bf16, bf16, float>(a, stream);
        case 7: return ln_dispatch_c<bf16, bf16, bf16>(a, stream);
        case 4: return ln_dispatch_c<bf16, float, float>(a, stream);
        default: return ln_dispatch_c<bf16, float, bf16>(a, stream);
    }
}

int pad_frames(const float* x, void* out, int out_bf16, int B, int T, int D, int left, int Tp, const int* lens, cudaStream_t stream) {
    dim3 grid((D / 4 + 255) / 256, Tp, B);
    if (out_bf16) pad_frames_kernel<bf16><<<grid, 256, 0, stream>>>(x, static_cast<bf16*>(out), T, D, left, Tp, lens);
    else pad_frames_kernel<float><<<grid, 256, 0, stream>>>(x, static_cast<float*>(out), T, D, left, Tp, lens);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int zero_padded_frames(float* x, int B, int T, int D, const int* lens, cudaStream_t stream) {
    if (!lens) return 0;
    dim3 grid((D / 4 + 255) / 256, T, B);
    zero_padded_kernel<<<grid, 256, 0, stream>>>(x, T, D, lens);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int convert_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t stream) {
    if (n <= 0) return 0;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    cvt_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, static_cast<bf16*>(out), n);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int synth_clips(float* out, long long first_utt, int count, int samples, cudaStream_t stream) {
    if (count <= 0) return 0;
    const float scale = (float)(1.0 / sqrt(4.0 * (65536.0 * 65536.0 - 1.0) / 12.0));
    dim3 grid(64, count);
    synth_kernel<<<grid, 256, 0, stream>>>(out, first_utt, samples, scale);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ingest_pcm16(const int16_t* pcm, const long long* offsets, const int* lens, int B, int S, float* out, cudaStream_t stream) {
    if (B <= 0 || S <= 0) return 0;
    dim3 grid((S + 1023) / 1024, B);
    ingest_pcm16_kernel<<<grid, 256, 0, stream>>>(pcm, offsets, lens, out, S);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int frame_lengths(const int* sample_lens, int* frame_lens, int B, int n_conv, const int* k, const int* s, cudaStream_t stream) {
    ConvCfg c{};
    c.n = n_conv;
    for (int i = 0; i < n_conv && i < 8; ++i) { c.k[i] = k[i]; c.s[i] = s[i]; }
    frame_len_kernel<<<(B + 127) / 128, 128, 0, stream>>>(sample_lens, frame_lens, B, c);
    SLSB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace slsb

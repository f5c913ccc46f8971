// Engine: weight arena, workspace, the forward schedule of the scoring path, and the C ABI (include/slsb200.h).
// Host code only orchestrates launches on the caller's stream; it never synchronises inside slsb_forward.
#include "common.cuh"
#include "kernels.h"
#include "../../include/slsb200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace slsb {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// tensor-map encode via the driver entry point (resolved lazily; no libcuda link dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
static int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, bool swizzle64 = false) {
    EncodeTiledFn fn = get_encode();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return -1; }
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] stride0=%llu box=[%u,%u,%u,%u] base=%p", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0,
                  rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
        return -1;
    }
    return 0;
}
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, bool sw) {
    return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, sw);
}
int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, bool sw) {
    return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, sw);
}

int encode_tmap_f32_sw64(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, false, true);
}

int encode_tmap_bf16_sw64(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, false, true);
}

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* v = getenv("SLSB_NO_PDL"); on = (v && atoi(v) != 0) ? 0 : 1; }
    return on != 0;
}

// small elementwise helper kernels private to the engine
__global__ void center_rows_kernel(const float* __restrict__ x, const float* __restrict__ sub, float* __restrict__ of, bf16* __restrict__ ob, long long n, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i] - sub[i % C];
    if (of) of[i] = v;
    if (ob) ob[i] = __float2bfloat16_rn(v);
}
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __bfloat162float(in[i]);
}

}  // namespace slsb

using namespace slsb;

// ------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------
struct Weight {
    int64_t numel = 0;
    float* f32 = nullptr;
    bf16* b16 = nullptr;
    bool gemm = false;   // gets a bf16 copy
    bool set = false;
};

struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
    int reserve(size_t need) {
        if (need <= bytes) return 0;
        if (p) { cudaDeviceSynchronize(); cudaFree(p); p = nullptr; bytes = 0; }
        need = (need + (size_t(4) << 20)) & ~size_t(255);      // slack: TMA boxes may touch one partial tile past the end
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e)); return -1; }
        cudaMemset(p, 0, need);
        bytes = need;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct slsb_engine {
    slsb_config cfg{};
    int device = 0, num_sms = 148;
    std::map<std::string, Weight> w;
    bool finalized = false;
    int64_t launches = 0;
    // workspace
    Buf fe[2], lnbuf, qkv, attn, ffn, xmid, xfinal, xc, xpad, acts, encoded, sums, votes, thr, cut, thr_w, cut_w, pool_part, wmask, flac_bytes, flac_frames, flac_status, pooled, logprob,
        sls_w, sls_in, sls_part, sls_dots, zeros, scratch, flens, wav_stage[2], lens_stage[2], score_stage[4], recon, tmp_bf16, im2col, conv0_w64, conv0_wb64, conv0_gram, ybuf, pcm_stage, off_stage;
    // pipelined host scoring (slsb_score_submit / slsb_score_wait): uploads run on a private copy stream into two staging
    // slots so the H2D copy of batch i+1 overlaps the forward of batch i; up to 4 submissions may be in flight
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_slot_free[2] = {nullptr, nullptr}, ev_done[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t submit_seq = 0;
    std::vector<Buf> X;
    // in-place stream mode (bf16 path): ONE fp32 residual stream `X[0]` advanced by reduce-add GEMM epilogues, bf16 snapshots of the
    // 24 layer results (`snap`) written by the LayerNorm kernel that reads them
    std::vector<Buf> snap;
    bool inplace = false;
    void* l2_ptr = nullptr; size_t l2_bytes = 0, l2_carve = 0; cudaStream_t l2_stream = nullptr;   // persisting-L2 window currently installed
    long long l2_max_carve = -1, l2_max_window = 0;
    // last call
    int B = 0, S = 0, T = 0, prec = 0, head = 0;
    bool have_lens = false, have_acts = false, have_sel = false, have_dots = false, have_snap = false;
    int sls_ks = 17, sls_kp = 0;      // fp32 path: 17-way split-K over Kp (a multiple of 16 * 17 * 4 = 1088, so also of the 64-wide k-blocks)
    // optional per-launch CUDA-event timing of the tensor-core kernels (bench.py roofline)
    bool profiling = false;
    struct ProfRec { cudaEvent_t a, b; double flops; int kind; };
    std::vector<ProfRec> prof;
};

// kinds 0-7 count FLOPs, kinds 8+ count algorithmic HBM bytes (HBM-bound kernels)
enum ProfKind { PK_ENC_QKV = 0, PK_ENC_OUT = 1, PK_ENC_FC1 = 2, PK_ENC_FC2 = 3, PK_CONV_GEMM = 4, PK_POS_GEMM = 5, PK_OTHER_GEMM = 6, PK_ATTN = 7,
                PK_LN = 8, PK_SLS_POOL = 9, PK_SLS_FC1 = 10, PK_SAE_GEMM = 11, PK_SAE_SELECT = 12, PK_CLS = 13, PK_COUNT = 14 };

struct ProfScope {
    slsb_engine* e; cudaStream_t st; int idx = -1;
    ProfScope(slsb_engine* e_, cudaStream_t st_, int kind, double flops) : e(e_), st(st_) {
        if (!e->profiling) return;
        slsb_engine::ProfRec r{};
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        r.flops = flops; r.kind = kind;
        cudaEventRecord(r.a, st);
        e->prof.push_back(r);
        idx = (int)e->prof.size() - 1;
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(e->prof[idx].b, st); }
};

// Every C-ABI entry that takes an engine runs on the engine's device and leaves the caller's current device as it found it
// (a process may touch several GPUs: torch's current device need not be the engine's).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int conv_len(const slsb_config& c, int n, int upto = -1) {
    const int last = upto < 0 ? c.n_conv : upto;
    for (int i = 0; i < last; ++i) n = n < c.conv_kernel[i] ? 0 : (n - c.conv_kernel[i]) / c.conv_stride[i] + 1;   // C division truncates: no frames below the kernel width
    return n;
}

static void add_weight(slsb_engine* e, const std::string& name, int64_t numel, bool gemm = false) {
    Weight wt;
    wt.numel = numel;
    wt.gemm = gemm;
    e->w[name] = wt;
}

static int build_weight_table(slsb_engine* e) {
    const slsb_config& c = e->cfg;
    const int C = c.conv_dim, D = c.embed_dim, F = c.ffn_dim;
    add_weight(e, "conv0.w", (int64_t)C * c.conv_kernel[0]);
    for (int i = 0; i < c.n_conv; ++i) {
        const std::string p = "conv" + std::to_string(i);
        if (i > 0) add_weight(e, p + ".w", (int64_t)C * c.conv_kernel[i] * C, true);
        add_weight(e, p + ".b", C);
        add_weight(e, p + ".ln.w", C);
        add_weight(e, p + ".ln.b", C);
    }
    add_weight(e, "feat_ln.w", C); add_weight(e, "feat_ln.b", C);
    add_weight(e, "proj.w", (int64_t)D * C, true); add_weight(e, "proj.b", D);
    add_weight(e, "pos.w", (int64_t)D * c.pos_kernel * (D / c.pos_groups), true); add_weight(e, "pos.b", D);
    for (int l = 0; l < c.n_layers; ++l) {
        const std::string p = "L" + std::to_string(l);
        add_weight(e, p + ".ln1.w", D); add_weight(e, p + ".ln1.b", D);
        add_weight(e, p + ".qkv.w", (int64_t)3 * D * D, true); add_weight(e, p + ".qkv.b", 3 * D);
        add_weight(e, p + ".out.w", (int64_t)D * D, true); add_weight(e, p + ".out.b", D);
        add_weight(e, p + ".ln2.w", D); add_weight(e, p + ".ln2.b", D);
        add_weight(e, p + ".fc1.w", (int64_t)F * D, true); add_weight(e, p + ".fc1.b", F);
        add_weight(e, p + ".fc2.w", (int64_t)D * F, true); add_weight(e, p + ".fc2.b", D);
    }
    add_weight(e, "enc_ln.w", D); add_weight(e, "enc_ln.b", D);
    if (c.sae_dict > 0) {
        add_weight(e, "sae.enc.w", (int64_t)c.sae_dict * D, true); add_weight(e, "sae.enc.b", c.sae_dict);
        add_weight(e, "sae.b_dec", D);
        add_weight(e, "sae.dec.w", (int64_t)D * c.sae_dict, true);
    }
    if (c.cls_in > 0) {
        add_weight(e, "cls.ln.w", c.cls_in); add_weight(e, "cls.ln.b", c.cls_in);
        add_weight(e, "cls.fc1.w", (int64_t)c.cls_hidden * c.cls_in); add_weight(e, "cls.fc1.b", c.cls_hidden);
        add_weight(e, "cls.fc2.w", (int64_t)2 * c.cls_hidden); add_weight(e, "cls.fc2.b", 2);
    }
    if (c.sls_frames > 0) {
        const int kraw = (c.sls_frames / 3) * (D / 3);
        const int q = 64 * e->sls_ks;
        e->sls_kp = (kraw + q - 1) / q * q;
        add_weight(e, "sls.fc0.w", D); add_weight(e, "sls.fc0.b", 1);
        add_weight(e, "sls.bn", 4);
        add_weight(e, "sls.fc1.w", (int64_t)c.sls_hidden * e->sls_kp, true); add_weight(e, "sls.fc1.b", c.sls_hidden);
        add_weight(e, "sls.fc3.w", (int64_t)2 * c.sls_hidden); add_weight(e, "sls.fc3.b", 2);
    }
    for (auto& kv : e->w) {
        cudaError_t err = cudaMalloc(&kv.second.f32, kv.second.numel * sizeof(float));
        if (err == cudaSuccess && kv.second.gemm) err = cudaMalloc(&kv.second.b16, kv.second.numel * sizeof(bf16));
        if (err != cudaSuccess) { set_error("weight arena cudaMalloc failed for %s: %s", kv.first.c_str(), cudaGetErrorString(err)); return -1; }
    }
    return 0;
}

#define W32(name) (e->w.at(name).f32)
#define W16(name) (e->w.at(name).b16)
#define LAUNCH(call) do { ++e->launches; if ((call) != 0) return -1; } while (0)

// C = act(A W^T + b) (+res): fp32 path -> CUDA cores, bf16 path -> tcgen05
static int linear(slsb_engine* e, bool bf, const void* A, long long lda, const std::string& wname, int N, int K, long long M,
                  const float* bias, const float* residual, long long ldr, void* out, long long ldc, int out_bf16, int act, cudaStream_t st,
                  int kind = PK_OTHER_GEMM) {
    ProfScope ps(e, st, kind, 2.0 * (double)M * N * K);
    if (bf) {
        TcGemmArgs g;
        g.a_mode = A_PLAIN; g.A = A; g.lda = lda; g.W = W16(wname); g.ldw = K; g.M = (int)M; g.N = N; g.K = K; g.batches = 1;
        g.out = out; g.ldc = ldc; g.out_bf16 = out_bf16; g.bias = bias; g.residual = residual; g.ldr = ldr; g.act = act;
        LAUNCH(tc_gemm(g, e->num_sms, st));
    } else {
        SimtGemmArgs g;
        g.A = A; g.lda = lda; g.W = W32(wname); g.ldw = K; g.M = (int)M; g.N = N; g.K = K; g.batches = 1;
        g.out = out; g.ldc = ldc; g.bias = bias; g.residual = residual; g.ldr = ldr; g.act = act; g.exact_gelu = 1;
        LAUNCH(simt_gemm(g, st));
    }
    return 0;
}

static int conv_layer(slsb_engine* e, bool bf, const void* x, const void* Wp, const float* bias, void* out, int B, int Lin, int C, int N,
                      int k, int s, cudaStream_t st) {
    const int Lout = (Lin - k) / s + 1;
    ProfScope ps(e, st, PK_CONV_GEMM, 2.0 * (double)B * Lout * N * k * C);
    if (bf) {
        TcGemmArgs g;
        g.a_mode = A_CONV; g.A = x; g.W = Wp; g.ldw = (long long)k * C; g.M = Lout; g.N = N; g.K = k * C; g.batches = B;
        g.conv_cin = C; g.conv_stride = s; g.conv_lin = Lin;
        g.out = out; g.ldc = N; g.out_batch_stride = (long long)Lout * N; g.out_bf16 = 1; g.bias = bias; g.act = ACT_NONE;
        LAUNCH(tc_gemm(g, e->num_sms, st));
    } else {
        SimtGemmArgs g;
        g.A = x; g.lda = (long long)s * C; g.a_batch_stride = (long long)Lin * C; g.W = static_cast<const float*>(Wp); g.ldw = (long long)k * C;
        g.M = Lout; g.N = N; g.K = k * C; g.batches = B;
        g.out = out; g.ldc = N; g.out_batch_stride = (long long)Lout * N; g.bias = bias; g.act = ACT_NONE;
        LAUNCH(simt_gemm(g, st));
    }
    return 0;
}

// x_out = x + gelu(pos_conv(x) + b)  (wav2vec2.py:915-917); xpad is scratch [B, T + K, D]
static int pos_conv(slsb_engine* e, bool bf, const float* x, const void* Wp, const float* bias, float* out, void* xpad, int B, int T, int D,
                    int K, int groups, const int* flens, cudaStream_t st) {
    const int gw = D / groups;
    // bf16: Toeplitz GEMM with 128 x 256 tiles (A_POS4) unless SLSB_POS_V1=1 asks for the 128 x 64 kernel; the padded stream then
    // has a multiple of 4 frames per utterance (the extra frames are zero like the rest of the padding)
    static int pos_v1 = -1;
    if (pos_v1 < 0) { const char* v = getenv("SLSB_POS_V1"); pos_v1 = (v && atoi(v) != 0) ? 1 : 0; }
    const bool pos4 = bf && !pos_v1 && gw == 64;
    const int Tp = pos4 ? (T + K + 3) / 4 * 4 : T + K;
    LAUNCH(pad_frames(x, xpad, bf ? 1 : 0, B, T, D, K / 2, Tp, flens, st));
    ProfScope ps(e, st, PK_POS_GEMM, 2.0 * (double)B * T * D * K * gw);
    if (bf) {
        if (gw != 64) { set_error("pos_conv (tcgen05): group width %d != 64", gw); return -1; }
        TcGemmArgs g;
        g.a_mode = pos4 ? A_POS4 : A_POS; g.A = xpad; g.W = Wp; g.ldw = (long long)K * gw; g.M = T; g.N = D; g.K = K * gw; g.batches = B;
        g.pos_dim = D; g.pos_tp = Tp;
        g.out = out; g.ldc = D; g.out_batch_stride = (long long)T * D; g.out_bf16 = 0; g.bias = bias;
        g.residual = x; g.ldr = D; g.res_batch_stride = (long long)T * D; g.act = ACT_GELU;
        LAUNCH(tc_gemm(g, e->num_sms, st));
    } else {
        SimtGemmArgs g;
        g.A = xpad; g.lda = D; g.a_batch_stride = (long long)Tp * D; g.a_kinner = gw; g.a_kouter = D;
        g.W = static_cast<const float*>(Wp); g.ldw = (long long)K * gw; g.w_group_stride = (long long)gw * K * gw;
        g.groups = groups; g.a_group_offset = gw; g.n_per_group = gw;
        g.M = T; g.N = gw; g.K = K * gw; g.batches = B;
        g.out = out; g.ldc = D; g.out_batch_stride = (long long)T * D; g.bias = bias;
        g.residual = x; g.ldr = D; g.res_batch_stride = (long long)T * D; g.act = ACT_GELU; g.exact_gelu = 1;
        LAUNCH(simt_gemm(g, st));
    }
    return 0;
}

static int attention(slsb_engine* e, bool bf, const void* qkv, void* out, int B, int T, int H, const int* flens, cudaStream_t st) {
    ProfScope ps(e, st, PK_ATTN, 4.0 * (double)B * H * T * T * 64);
    int impl = e->cfg.attn_impl;
    static int impl_env = -1;                     // SLSB_ATTN_IMPL=<1|2|3>: A/B override of an AUTO config (race hunts, profiling)
    if (impl_env < 0) { const char* v = getenv("SLSB_ATTN_IMPL"); impl_env = v ? atoi(v) : 0; }
    if (impl == SLSB_ATTN_AUTO && impl_env > 0) impl = impl_env;
    if (impl == SLSB_ATTN_AUTO) impl = (bf && T <= 512) ? SLSB_ATTN_TC : SLSB_ATTN_SIMT;      // 512 frames = 10.2 s: bf16 never leaves tcgen05 (config 4)
    if ((impl == SLSB_ATTN_TC && T > 512) || (impl == SLSB_ATTN_TC_V1 && T > 256)) impl = SLSB_ATTN_SIMT;
    if (impl == SLSB_ATTN_TC && bf) LAUNCH(attention_tc(qkv, out, B, T, H, flens, e->num_sms, st));
    else if (impl == SLSB_ATTN_TC_V1 && bf) LAUNCH(attention_tc_v1(qkv, out, B, T, H, flens, e->num_sms, st));
    else LAUNCH(attention_simt(qkv, out, bf ? 1 : 0, B, T, H, flens, st));
    return 0;
}

// fused conv/GEMM + LayerNorm(512) + GELU.  K >= 256 (conv1..6): CTA-pair kernel, whose double-buffered accumulators hide the
// epilogue behind the next tile's MMAs (measured conv1..6: 1208 us vs 1571 us per 64-clip batch).  K = 64 (conv0) has no
// mainloop to hide anything behind and pays the pair's statistics exchange on every tile (722 us vs 515 us): it stays on the
// one-CTA-per-row-block kernel.  SLSB_LN_GEMM_V1=1 / =2 force the old / the pair kernel everywhere (A/B measurements).
#ifndef SLSB_CONV0_V1_DEFAULT
#define SLSB_CONV0_V1_DEFAULT 0      // 1: im2col + tc_gemm_ln_kernel (round-1 conv0) unless SLSB_CONV0_V1=0
#endif
// conv0: the one-kernel form (conv0_tc.cu) for the XLS-R geometry; SLSB_CONV0_V1=1 selects im2col + tc_gemm_ln_kernel (A/B)
static bool conv0_one_kernel(const slsb_config& c) {
    static int v1 = -1;
    if (v1 < 0) { const char* v = getenv("SLSB_CONV0_V1"); v1 = v ? (atoi(v) != 0 ? 1 : 0) : SLSB_CONV0_V1_DEFAULT; }
    return !v1 && c.conv_dim == 512 && c.conv_kernel[0] >= 1 && c.conv_kernel[0] <= 15;
}

#ifndef SLSB_LN_GEMM_EPI8_DEFAULT
#define SLSB_LN_GEMM_EPI8_DEFAULT 0      // 1: conv1..6 on the 8-epilogue-warp pair kernel unless SLSB_LN_GEMM_EPI8=0
#endif
static int ln_gemm_dispatch(const TcLnGemmArgs& g, int num_sms, cudaStream_t st) {
    static int force = -1;
    if (force < 0) { const char* v = getenv("SLSB_LN_GEMM_V1"); force = v ? atoi(v) : 0; }
    const bool pair = force == 2 || (force != 1 && g.K >= 256);
    // SLSB_LN_GEMM_EPI8=1: the pair kernel with 8 epilogue warps (gemm_tc_ln2.cu) instead of 16 (gemm_tc_ln2x.cu)
    static int epi8 = -1;
    if (epi8 < 0) { const char* v = getenv("SLSB_LN_GEMM_EPI8"); epi8 = v ? (atoi(v) != 0 ? 1 : 0) : SLSB_LN_GEMM_EPI8_DEFAULT; }
    if (pair && !epi8) return tc_gemm_ln_gelu_pair16(g, num_sms, st);
    return pair ? tc_gemm_ln_gelu_pair(g, num_sms, st) : tc_gemm_ln_gelu(g, num_sms, st);
}

// LayerNorm launch with its algorithmic HBM bytes recorded (every operand is touched exactly once)
static int layernorm_timed(slsb_engine* e, const LnArgs& a, cudaStream_t st) {
    const double n = (double)a.rows * a.C;
    const double bytes = n * ((a.in_bf16 ? 2 : 4) + (a.add ? 2 : 0) + (a.sum_out ? 4 : 0) + (a.out ? (a.out_bf16 ? 2 : 4) : 0) +
                              (a.out2 ? (a.out2_bf16 ? 2 : 4) : 0) + (a.copy_out ? 2 : 0));      // copy_out: the bf16 layer-result snapshot
    ProfScope ps(e, st, PK_LN, bytes);
    LAUNCH(layernorm(a, st));
    return 0;
}

static int check_ready(slsb_engine* e) {
    if (!e) { set_error("null engine"); return -1; }
    if (!e->finalized) { set_error("weights not finalized: call slsb_set_weight for every tensor, then slsb_finalize_weights"); return -1; }
    return 0;
}

// ---- the trunk: wav -> X[0..n_layers], xfinal (+ xc = xfinal - b_dec) ---------------------------
static int run_trunk(slsb_engine* e, const float* wav, const int* slens, int B, int S, int prec, int head_flags, cudaStream_t st) {
    const slsb_config& c = e->cfg;
    const int head = head_flags & 0xFF;
    const bool retain = (head_flags & SLSB_HEAD_RETAIN) != 0;
    const bool bf = prec == SLSB_PREC_BF16;
    const size_t es = bf ? 2 : 4;
    const int C = c.conv_dim, D = c.embed_dim, F = c.ffn_dim, H = c.n_heads;
    const bool want_xc = head == SLSB_HEAD_SAE || head == SLSB_HEAD_WINDOW;
    if (B <= 0 || S <= 0) { set_error("empty batch (B=%d, S=%d)", B, S); return -1; }
    std::vector<int> L(c.n_conv);
    for (int i = 0; i < c.n_conv; ++i) L[i] = conv_len(c, S, i + 1);
    const int T = L[c.n_conv - 1];
    if (T < 1) { set_error("clip too short: %d samples give %d frames", S, T); return -1; }
    const long long M = (long long)B * T;
    const int Tp = T + c.pos_kernel;

    if (e->fe[0].reserve((size_t)B * L[0] * C * es)) return -1;
    if (e->fe[1].reserve((size_t)B * (c.n_conv > 1 ? L[1] : 1) * C * es)) return -1;
    if (e->lnbuf.reserve((size_t)M * D * es) || e->qkv.reserve((size_t)M * 3 * D * es) || e->attn.reserve((size_t)M * D * es) ||
        e->ffn.reserve((size_t)M * F * es) || e->xmid.reserve((size_t)M * D * 4) || e->xfinal.reserve((size_t)M * D * 4) ||
        e->xc.reserve((size_t)M * D * es) || e->xpad.reserve((size_t)B * (Tp + 4) * D * es) || e->flens.reserve((size_t)B * 4)) return -1;
    // SLSB_INPLACE (default 1, bf16 path): the residual stream lives in ONE fp32 buffer that out_proj / fc2 advance with TMA reduce-add
    // stores (no residual loads); layer result l is snapshotted in bf16 by the LayerNorm that reads it (the SLS head then reads
    // 2 B per element instead of 4).  0: one fp32 buffer per layer result, residual tiles loaded by TMA.
    static int inplace_env = -1;
    if (inplace_env < 0) { const char* v = getenv("SLSB_INPLACE"); inplace_env = v ? atoi(v) : 1; }
    const bool inplace = bf && inplace_env != 0;
    e->inplace = inplace;
    // bf16 layer-result snapshots (wav2vec2.py:958 layer_results): consumed by the SLS head, otherwise only on request
    const bool want_snap = head == SLSB_HEAD_SLS || retain;
    e->have_snap = !inplace || want_snap;
    if ((int)e->X.size() < c.n_layers + 1) e->X.resize(c.n_layers + 1);
    if ((int)e->snap.size() < c.n_layers) e->snap.resize(c.n_layers);
    if (inplace) {
        if (e->X[0].reserve((size_t)M * D * 4)) return -1;
        for (int l = 0; l < c.n_layers; ++l) if (e->snap[l].reserve((size_t)M * D * 2)) return -1;
        // SLSB_L2_PERSIST=1 (experiment, default off): the fp32 residual stream (52.7 MB at B = 64) is read by every LayerNorm and
        // updated by every out_proj / fc2, while fc2 streams 105 MB of activations through the same L2 in between; an access-policy
        // window on the launch stream marks the stream's lines persisting (carve-out = the stream's size).  Measured on B200:
        // fixed 4-s batches A/B/A/B 11.64 / 11.51 / 11.66 / 11.61 ms per step (LayerNorm 1.23 -> 1.19, out_proj 0.99 -> 0.95, fc2
        // 2.26 -> 2.20 ms), i.e. +0.5-1 %; but the 1-10 s length mix lost 6 % with a window per batch shape and 20 % with one window
        // over the whole allocation (a 130 MB stream cannot be resident and the 79 MB carve-out starves every other operand): not a
        // default.  slsb_destroy takes the window off again.
        static int l2p = -1;
        if (l2p < 0) { const char* v = getenv("SLSB_L2_PERSIST"); l2p = v ? atoi(v) : 0; }
        const size_t bytes = (size_t)M * D * 4;
        if (l2p && (e->l2_ptr != e->X[0].p || e->l2_bytes != bytes || e->l2_stream != st)) {
            if (e->l2_max_carve < 0) {
                int a = 0, b = 0;
                SLSB_CUDA_CHECK(cudaDeviceGetAttribute(&a, cudaDevAttrMaxPersistingL2CacheSize, e->device));
                SLSB_CUDA_CHECK(cudaDeviceGetAttribute(&b, cudaDevAttrMaxAccessPolicyWindowSize, e->device));
                e->l2_max_carve = a; e->l2_max_window = b;
            }
            const size_t carve = bytes < (size_t)e->l2_max_carve ? bytes : (size_t)e->l2_max_carve;
            const size_t win = bytes < (size_t)e->l2_max_window ? bytes : (size_t)e->l2_max_window;
            if (carve > 0 && win > 0) {
                if (carve > e->l2_carve) {                      // the carve-out only grows (variable-length batches change M every forward)
                    SLSB_CUDA_CHECK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
                    e->l2_carve = carve;
                }
                cudaStreamAttrValue v{};
                v.accessPolicyWindow.base_ptr = e->X[0].p;
                v.accessPolicyWindow.num_bytes = win;
                v.accessPolicyWindow.hitRatio = e->l2_carve >= win ? 1.0f : (float)e->l2_carve / (float)win;
                v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
                SLSB_CUDA_CHECK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v));
            }
            e->l2_ptr = e->X[0].p; e->l2_bytes = bytes; e->l2_stream = st;
        }
    } else {
        for (int l = 0; l <= c.n_layers; ++l) if (e->X[l].reserve((size_t)M * D * 4)) return -1;
    }

    int* flens = nullptr;
    if (slens) {
        flens = e->flens.as<int>();
        LAUNCH(frame_lengths(slens, flens, B, c.n_conv, c.conv_kernel, c.conv_stride, st));
    }
    e->B = B; e->S = S; e->T = T; e->prec = prec; e->have_lens = slens != nullptr; e->have_acts = false; e->have_sel = false; e->have_dots = false;

    // 1. feature extractor (wav2vec2.py:843-851)
    const bool fused_ln = bf && c.reserved[0] == 0;     // reserved[0] = 1 disables the fused conv+LN+GELU tensor-core kernel
    if (fused_ln) {
        // bf16: every layer is one tcgen05 kernel with the LayerNorm + GELU in its epilogue
        if (conv0_one_kernel(c)) {
            ProfScope ps(e, st, PK_CONV_GEMM, 2.0 * (double)B * L[0] * C * c.conv_kernel[0]);
            LAUNCH(conv0_tc(wav, e->conv0_wb64.p, e->conv0_gram.as<float>(), W32("conv0.ln.w"), W32("conv0.ln.b"), e->fe[0].p, B, S, L[0],
                            c.conv_kernel[0], c.conv_stride[0], 1e-5f, e->num_sms, st));
        } else {
            if (e->im2col.reserve((size_t)B * L[0] * 64 * 2)) return -1;
            LAUNCH(conv0_im2col(wav, e->im2col.p, B, S, L[0], c.conv_kernel[0], c.conv_stride[0], st));
            ProfScope ps(e, st, PK_CONV_GEMM, 2.0 * (double)B * L[0] * C * c.conv_kernel[0]);
            TcLnGemmArgs g;
            g.a_mode = A_PLAIN; g.A = e->im2col.p; g.lda = 64; g.W = e->conv0_w64.p; g.M = B * L[0]; g.K = 64; g.batches = 1;
            g.out = e->fe[0].p; g.bias = W32("conv0.b"); g.ln_w = W32("conv0.ln.w"); g.ln_b = W32("conv0.ln.b");
            LAUNCH(ln_gemm_dispatch(g, e->num_sms, st));
        }
        for (int i = 1; i < c.n_conv; ++i) {
            const std::string p = "conv" + std::to_string(i);
            ProfScope ps(e, st, PK_CONV_GEMM, 2.0 * (double)B * L[i] * C * c.conv_kernel[i] * C);
            TcLnGemmArgs g;
            g.a_mode = A_CONV; g.A = e->fe[(i - 1) & 1].p; g.W = W16(p + ".w"); g.M = L[i]; g.K = c.conv_kernel[i] * C; g.batches = B;
            g.conv_cin = C; g.conv_stride = c.conv_stride[i]; g.conv_lin = L[i - 1];
            g.out = e->fe[i & 1].p; g.out_batch_stride = (long long)L[i] * C;
            g.bias = W32(p + ".b"); g.ln_w = W32(p + ".ln.w"); g.ln_b = W32(p + ".ln.b");
            LAUNCH(ln_gemm_dispatch(g, e->num_sms, st));
        }
    } else {
        LAUNCH(conv0_ln_gelu(wav, B, S, L[0], c.conv_kernel[0], c.conv_stride[0], W32("conv0.w"), W32("conv0.b"), W32("conv0.ln.w"),
                             W32("conv0.ln.b"), e->fe[0].p, bf ? 1 : 0, C, !bf, st));
        for (int i = 1; i < c.n_conv; ++i) {
            const std::string p = "conv" + std::to_string(i);
            void* in = e->fe[(i - 1) & 1].p;
            void* out = e->fe[i & 1].p;
            if (conv_layer(e, bf, in, bf ? (const void*)W16(p + ".w") : (const void*)W32(p + ".w"), W32(p + ".b"), out, B, L[i - 1], C, C,
                           c.conv_kernel[i], c.conv_stride[i], st)) return -1;
            LnArgs a;
            a.in = out; a.in_bf16 = bf; a.out = out; a.out_bf16 = bf; a.w = W32(p + ".ln.w"); a.b = W32(p + ".ln.b");
            a.rows = (long long)B * L[i]; a.C = C; a.gelu = 1; a.exact_gelu = !bf;
            LAUNCH(layernorm(a, st));
        }
    }
    // 2. LayerNorm(512) + post_extract_proj (wav2vec2.py:563-564, :595-596)
    {
        LnArgs a;
        a.in = e->fe[(c.n_conv - 1) & 1].p; a.in_bf16 = bf; a.out = e->lnbuf.p; a.out_bf16 = bf; a.w = W32("feat_ln.w"); a.b = W32("feat_ln.b");
        a.rows = M; a.C = C;
        LAUNCH(layernorm(a, st));
        if (linear(e, bf, e->lnbuf.p, C, "proj.w", D, C, M, W32("proj.b"), nullptr, 0, e->xmid.p, D, 0, ACT_NONE, st)) return -1;
    }
    // 3. positional conv + GELU + residual (wav2vec2.py:910-917)
    if (pos_conv(e, bf, e->xmid.as<float>(), bf ? (const void*)W16("pos.w") : (const void*)W32("pos.w"), W32("pos.b"), e->X[0].as<float>(),
                 e->xpad.p, B, T, D, c.pos_kernel, c.pos_groups, flens, st)) return -1;
    // 4. transformer layers, pre-LN (wav2vec2.py:1044-1062)
    if (bf) {
        // bf16: the GEMMs write their branch outputs in bf16 (TMA-store epilogue, no residual traffic inside the GEMM); the fp32
        // residual stream is advanced by the LayerNorm kernels, which are coalesced and HBM-bound anyway:
        //   LN1_l : X_l = xmid_{l-1} + fc2_{l-1}   (written as layer result l-1), lnbuf = LN(X_l)
        //   LN2_l : xmid = X_l + out_proj_l,        lnbuf = LN(xmid)
        if (e->ybuf.reserve((size_t)M * D * 2)) return -1;
        // SLS head: the LayerNorm kernel that writes layer result l also emits fc0 . x_l per frame (sls_layer_weights_from_dots)
        const bool want_dots = head == SLSB_HEAD_SLS && c.sls_frames > 0 && !slens;
        float* dots = nullptr;
        if (want_dots) {
            if (e->sls_dots.reserve((size_t)c.n_layers * M * 4)) return -1;
            dots = e->sls_dots.as<float>();
            e->have_dots = true;
        }
        // SLSB_RES_IN_GEMM (default 1): out_proj / fc2 add the fp32 residual stream in their epilogue (residual tile in and sum tile
        // out through TMA, gemm_tc.cu::epilogue_tile_res_tma) and the LayerNorm kernels only read the stream (4 B) and write bf16.
        // 0: the GEMMs write bf16 branch outputs and the LayerNorm kernels advance the stream (4 + 2 B read, 4 + 2 B written).
        static int res_in_gemm = -1;
        if (res_in_gemm < 0) { const char* v = getenv("SLSB_RES_IN_GEMM"); res_in_gemm = v ? atoi(v) : 1; }
        for (int l = 0; l < c.n_layers; ++l) {
            const std::string p = "L" + std::to_string(l);
            if (inplace) {
                float* xs = e->X[0].as<float>();
                LnArgs a;      // LN1 reads the stream = layer result l - 1: snapshot it (bf16) and emit the SLS fc0 dots on the way
                a.in = xs; a.out = e->lnbuf.p; a.out_bf16 = 1; a.w = W32(p + ".ln1.w"); a.b = W32(p + ".ln1.b"); a.rows = M; a.C = D;
                if (l > 0 && want_snap) {
                    a.copy_out = e->snap[l - 1].p;
                    if (want_dots) { a.dot_w = W32("sls.fc0.w"); a.dot_out = dots + (long long)(l - 1) * M; }
                }
                if (layernorm_timed(e, a, st)) return -1;
                if (linear(e, true, e->lnbuf.p, D, p + ".qkv.w", 3 * D, D, M, W32(p + ".qkv.b"), nullptr, 0, e->qkv.p, 3 * D, 1, ACT_NONE, st, PK_ENC_QKV)) return -1;
                if (attention(e, true, e->qkv.p, e->attn.p, B, T, H, flens, st)) return -1;
                if (linear(e, true, e->attn.p, D, p + ".out.w", D, D, M, W32(p + ".out.b"), xs, D, xs, D, 0, ACT_NONE, st, PK_ENC_OUT)) return -1;
                LnArgs a2;
                a2.in = xs; a2.out = e->lnbuf.p; a2.out_bf16 = 1; a2.w = W32(p + ".ln2.w"); a2.b = W32(p + ".ln2.b"); a2.rows = M; a2.C = D;
                if (layernorm_timed(e, a2, st)) return -1;
                if (linear(e, true, e->lnbuf.p, D, p + ".fc1.w", F, D, M, W32(p + ".fc1.b"), nullptr, 0, e->ffn.p, F, 1, ACT_GELU, st, PK_ENC_FC1)) return -1;
                if (linear(e, true, e->ffn.p, F, p + ".fc2.w", D, F, M, W32(p + ".fc2.b"), xs, D, xs, D, 0, ACT_NONE, st, PK_ENC_FC2)) return -1;
                continue;
            }
            LnArgs a;
            a.out = e->lnbuf.p; a.out_bf16 = 1; a.w = W32(p + ".ln1.w"); a.b = W32(p + ".ln1.b"); a.rows = M; a.C = D;
            if (l == 0) a.in = e->X[0].p;
            else {
                if (res_in_gemm) a.in = e->X[l].p;
                else { a.in = e->xmid.p; a.add = e->ybuf.p; a.sum_out = e->X[l].as<float>(); }
                if (want_dots) { a.dot_w = W32("sls.fc0.w"); a.dot_out = dots + (long long)(l - 1) * M; }
            }
            if (layernorm_timed(e, a, st)) return -1;
            if (linear(e, true, e->lnbuf.p, D, p + ".qkv.w", 3 * D, D, M, W32(p + ".qkv.b"), nullptr, 0, e->qkv.p, 3 * D, 1, ACT_NONE, st, PK_ENC_QKV)) return -1;
            if (attention(e, true, e->qkv.p, e->attn.p, B, T, H, flens, st)) return -1;
            LnArgs a2;
            if (res_in_gemm) {
                if (linear(e, true, e->attn.p, D, p + ".out.w", D, D, M, W32(p + ".out.b"), e->X[l].as<float>(), D, e->xmid.p, D, 0, ACT_NONE, st, PK_ENC_OUT)) return -1;
                a2.in = e->xmid.p;
            } else {
                if (linear(e, true, e->attn.p, D, p + ".out.w", D, D, M, W32(p + ".out.b"), nullptr, 0, e->ybuf.p, D, 1, ACT_NONE, st, PK_ENC_OUT)) return -1;
                a2.in = e->X[l].p; a2.add = e->ybuf.p; a2.sum_out = e->xmid.as<float>();
            }
            a2.out = e->lnbuf.p; a2.out_bf16 = 1; a2.w = W32(p + ".ln2.w"); a2.b = W32(p + ".ln2.b"); a2.rows = M; a2.C = D;
            if (layernorm_timed(e, a2, st)) return -1;
            if (linear(e, true, e->lnbuf.p, D, p + ".fc1.w", F, D, M, W32(p + ".fc1.b"), nullptr, 0, e->ffn.p, F, 1, ACT_GELU, st, PK_ENC_FC1)) return -1;
            if (res_in_gemm) {
                if (linear(e, true, e->ffn.p, F, p + ".fc2.w", D, F, M, W32(p + ".fc2.b"), e->xmid.as<float>(), D, e->X[l + 1].p, D, 0, ACT_NONE, st, PK_ENC_FC2)) return -1;
            } else {
                if (linear(e, true, e->ffn.p, F, p + ".fc2.w", D, F, M, W32(p + ".fc2.b"), nullptr, 0, e->ybuf.p, D, 1, ACT_NONE, st, PK_ENC_FC2)) return -1;
            }
        }
        // 5. X_n = xmid + fc2_{n-1}; final LayerNorm on x only (wav2vec2.py:905-906); xc = x - b_dec feeds the SAE (model.py:70)
        LnArgs a;
        if (inplace) { a.in = e->X[0].p; if (want_snap) a.copy_out = e->snap[c.n_layers - 1].p; }
        else if (res_in_gemm) a.in = e->X[c.n_layers].p;
        else { a.in = e->xmid.p; a.add = e->ybuf.p; a.sum_out = e->X[c.n_layers].as<float>(); }
        a.out = e->xfinal.p; a.w = W32("enc_ln.w"); a.b = W32("enc_ln.b"); a.rows = M; a.C = D;
        if (want_xc && c.sae_dict > 0) { a.out2 = e->xc.p; a.out2_bf16 = 1; a.sub = W32("sae.b_dec"); }
        if (want_dots) { a.dot_w = W32("sls.fc0.w"); a.dot_out = dots + (long long)(c.n_layers - 1) * M; }
        if (layernorm_timed(e, a, st)) return -1;
        return 0;
    }
    for (int l = 0; l < c.n_layers; ++l) {
        const std::string p = "L" + std::to_string(l);
        float* xin = e->X[l].as<float>();
        float* xout = e->X[l + 1].as<float>();
        LnArgs a;
        a.in = xin; a.out = e->lnbuf.p; a.out_bf16 = bf; a.w = W32(p + ".ln1.w"); a.b = W32(p + ".ln1.b"); a.rows = M; a.C = D;
        LAUNCH(layernorm(a, st));
        if (linear(e, bf, e->lnbuf.p, D, p + ".qkv.w", 3 * D, D, M, W32(p + ".qkv.b"), nullptr, 0, e->qkv.p, 3 * D, bf, ACT_NONE, st, PK_ENC_QKV)) return -1;
        if (attention(e, bf, e->qkv.p, e->attn.p, B, T, H, flens, st)) return -1;
        if (linear(e, bf, e->attn.p, D, p + ".out.w", D, D, M, W32(p + ".out.b"), xin, D, e->xmid.p, D, 0, ACT_NONE, st, PK_ENC_OUT)) return -1;
        LnArgs a2;
        a2.in = e->xmid.p; a2.out = e->lnbuf.p; a2.out_bf16 = bf; a2.w = W32(p + ".ln2.w"); a2.b = W32(p + ".ln2.b"); a2.rows = M; a2.C = D;
        LAUNCH(layernorm(a2, st));
        if (linear(e, bf, e->lnbuf.p, D, p + ".fc1.w", F, D, M, W32(p + ".fc1.b"), nullptr, 0, e->ffn.p, F, bf, ACT_GELU, st, PK_ENC_FC1)) return -1;
        if (linear(e, bf, e->ffn.p, F, p + ".fc2.w", D, F, M, W32(p + ".fc2.b"), e->xmid.as<float>(), D, xout, D, 0, ACT_NONE, st, PK_ENC_FC2)) return -1;
    }
    // 5. final LayerNorm on x only (wav2vec2.py:905-906); xc = x - b_dec feeds the SAE encoder (model.py:70)
    {
        LnArgs a;
        a.in = e->X[c.n_layers].p; a.out = e->xfinal.p; a.w = W32("enc_ln.w"); a.b = W32("enc_ln.b"); a.rows = M; a.C = D;
        if (want_xc && c.sae_dict > 0) { a.out2 = e->xc.p; a.out2_bf16 = bf; a.sub = W32("sae.b_dec"); }
        LAUNCH(layernorm(a, st));
    }
    return 0;
}

// acts = relu(xc W_enc^T + b_enc) [rows, dict]   (model.py:70)
static int sae_acts(slsb_engine* e, bool bf, const void* xc, long long rows, cudaStream_t st) {
    const slsb_config& c = e->cfg;
    if (e->acts.reserve((size_t)rows * c.sae_dict * 4)) return -1;
    return linear(e, bf, xc, c.embed_dim, "sae.enc.w", c.sae_dict, c.embed_dim, rows, W32("sae.enc.b"), nullptr, 0, e->acts.p, c.sae_dict, 0, ACT_RELU, st, PK_SAE_GEMM);
}

// selection (+ pooling) pass: fills thr / cut; `pooled` (optional) receives the mean of the kept activations over the valid frames
// (model.py:245); the window variant materialises its votes only when `want_votes` (dense / compact code consumers).  `sel_out` is
// what the keep-rule looks at (activations, or the votes; nullptr when the votes were not kept).
static int sae_select(slsb_engine* e, long long rows, int T, int window, bool want_votes, float* pooled, const int* flens, const float** sel_out,
                      cudaStream_t st) {
    const slsb_config& c = e->cfg;
    const int Dd = c.sae_dict, k = c.sae_k;
    if (rows % T != 0) { set_error("top-k: rows=%lld is not a multiple of T=%d", rows, T); return -1; }
    const int B = (int)(rows / T);
    if (e->thr.reserve((size_t)rows * 4) || e->cut.reserve((size_t)rows * 4)) return -1;
    float* partial = nullptr;
    if (pooled) {
        if (e->pool_part.reserve((size_t)B * sel_chunks(T) * Dd * 4)) return -1;
        partial = e->pool_part.as<float>();
    }
    const float* acts = e->acts.as<float>();
    if (window <= 1) {
        LAUNCH(topk_select_pool(acts, B, T, Dd, k, flens, e->thr.as<float>(), e->cut.as<int>(), partial, pooled, st));
        if (pooled) ++e->launches;
        *sel_out = acts;
        return 0;
    }
    const int stride = window / 2 > 0 ? window / 2 : 1;
    if (T < window || stride >= T) { set_error("window top-k: T=%d shorter than window %d", T, window); return -1; }
    const int nw = (T - window) / stride + 1;                                  // model_window_topk.py:141
    if (e->wmask.reserve((size_t)B * nw * 256 * 4)) return -1;
    float* votes = nullptr;
    if (want_votes) {
        if (e->votes.reserve((size_t)rows * Dd * 4)) return -1;
        votes = e->votes.as<float>();
    }
    LAUNCH(window_select_pool(acts, B, T, Dd, k, window, stride, nw, e->wmask.as<uint32_t>(), e->thr.as<float>(), e->cut.as<int>(), votes,
                              partial, pooled, st));
    e->launches += pooled ? 2 : 1;
    *sel_out = votes;
    return 0;
}

static int run_head(slsb_engine* e, int head_flags, int prec, float* logprob, cudaStream_t st) {
    const slsb_config& c = e->cfg;
    const int head = head_flags & 0xFF;
    const bool retain = (head_flags & SLSB_HEAD_RETAIN) != 0;
    const bool bf = prec == SLSB_PREC_BF16;
    const int B = e->B, T = e->T, D = c.embed_dim;
    const long long M = (long long)B * T;
    const int* flens = e->have_lens ? e->flens.as<int>() : nullptr;
    if (head == SLSB_HEAD_SAE || head == SLSB_HEAD_WINDOW) {
        if (c.cls_in <= 0) { set_error("engine was created without classifier weights"); return -1; }
        if (e->pooled.reserve((size_t)B * c.cls_in * 4)) return -1;
        if (c.sae_dict > 0) {
            const int window = head == SLSB_HEAD_WINDOW ? c.sae_window : 1;
            if (window > 1 && flens) { set_error("window top-k with per-utterance lengths is not defined by the reference"); return -1; }
            if (sae_acts(e, bf, e->xc.p, M, st)) return -1;
            const float* sel = nullptr;
            {   // algorithmic bytes of selection + pooling: the fp32 activations read once
                ProfScope ps(e, st, PK_SAE_SELECT, (double)M * c.sae_dict * 4);
                const bool dense_next = c.cls_in != c.sae_dict;               // use_sparse_features=False densifies below
                if (sae_select(e, M, T, window, retain || dense_next, dense_next ? nullptr : e->pooled.as<float>(), flens, &sel, st)) return -1;
                e->have_acts = true; e->have_sel = sel != nullptr;
            }
            if (c.cls_in == c.sae_dict) {
            } else {
                // use_sparse_features=False: classifier sees the reconstruction (model.py:231-233)
                if (e->encoded.reserve((size_t)M * c.sae_dict * 4) || e->recon.reserve((size_t)M * D * 4)) return -1;
                LAUNCH(votes_densify(e->acts.as<float>(), sel, e->thr.as<float>(), e->cut.as<int>(), e->encoded.as<float>(), M, c.sae_dict, st));
                const void* A = e->encoded.p;
                if (bf) {
                    if (e->tmp_bf16.reserve((size_t)M * c.sae_dict * 2)) return -1;
                    LAUNCH(convert_f32_to_bf16(e->encoded.as<float>(), e->tmp_bf16.p, M * c.sae_dict, st));
                    A = e->tmp_bf16.p;
                }
                if (linear(e, bf, A, c.sae_dict, "sae.dec.w", D, c.sae_dict, M, W32("sae.b_dec"), nullptr, 0, e->recon.p, D, 0, ACT_NONE, st)) return -1;
                LAUNCH(mean_pool_frames(e->recon.as<float>(), e->pooled.as<float>(), B, T, D, flens, st));
            }
        } else {
            LAUNCH(mean_pool_frames(e->xfinal.as<float>(), e->pooled.as<float>(), B, T, D, flens, st));   // use_sae=False
        }
        if (e->scratch.reserve((size_t)(B * c.cls_hidden + 1024) * 4)) return -1;
        ++e->launches;
        ProfScope ps(e, st, PK_CLS, (double)c.cls_hidden * c.cls_in * 4);
        LAUNCH(classifier_head(e->pooled.as<float>(), B, c.cls_in, c.cls_hidden, W32("cls.ln.w"), W32("cls.ln.b"), W32("cls.fc1.w"), W32("cls.fc1.b"),
                               W32("cls.fc2.w"), W32("cls.fc2.b"), e->scratch.as<float>(), logprob, st));
        return 0;
    }
    if (head == SLSB_HEAD_SLS) {
        if (c.sls_frames <= 0) { set_error("engine was created without SLS weights"); return -1; }
        if (T != c.sls_frames) { set_error("SLS head: fc1 is sized for %d frames, got %d", c.sls_frames, T); return -1; }
        if (flens) { set_error("SLS head with per-utterance lengths is not defined by the reference"); return -1; }
        const int Kp = e->sls_kp, Hs = c.sls_hidden;
        // split-K over the 22 848-wide fc1 reduction: bf16 -> one tcgen05 tile per SM (n-tile x k-split), fp32 -> 17 SIMT groups
        int KS = e->sls_ks;
        if (bf) {
            const int total_kb = Kp / 64, n_tiles = Hs / 256 > 0 ? Hs / 256 : 1;
            int want = e->num_sms / n_tiles; if (want < 1) want = 1; if (want > total_kb) want = total_kb;
            const int per = (total_kb + want - 1) / want;
            KS = (total_kb + per - 1) / per;
        }
        if (e->sls_w.reserve((size_t)B * c.n_layers * 4) || e->sls_in.reserve((size_t)B * Kp * 4) ||
            e->sls_part.reserve((size_t)B * KS * Hs * 4) || e->zeros.reserve((size_t)(KS > 1 ? KS : 1) * Hs * 4)) return -1;
        std::vector<const void*> layers(c.n_layers);
        std::vector<const float*> layers_f32(c.n_layers);
        for (int l = 0; l < c.n_layers; ++l) {
            layers[l] = e->inplace ? e->snap[l].p : e->X[l + 1].p;
            layers_f32[l] = e->inplace ? nullptr : e->X[l + 1].as<float>();
        }
        if (e->have_dots) LAUNCH(sls_layer_weights_from_dots(e->sls_dots.as<float>(), c.n_layers, B, T, W32("sls.fc0.b"), e->sls_w.as<float>(), st));
        else if (e->inplace) { set_error("SLS head: the in-place stream needs the fused fc0 dots"); return -1; }
        else LAUNCH(sls_layer_weights(layers_f32.data(), c.n_layers, B, T, D, W32("sls.fc0.w"), W32("sls.fc0.b"), e->sls_w.as<float>(), nullptr, st));
        {   // every fp32 layer output is read once; the pooled [B, 67*341] matrix is written once
            ProfScope ps(e, st, PK_SLS_POOL, (double)c.n_layers * M * D * (e->inplace ? 2 : 4) + (double)B * Kp * (bf ? 2 : 4));
            LAUNCH(sls_fuse_pool(layers.data(), e->inplace ? 1 : 0, c.n_layers, e->sls_w.as<float>(), B, T, D, W32("sls.bn"), 1e-5f, e->sls_in.p, bf ? 1 : 0, Kp, st));
        }
        if (bf) {   // fc1 weight [1024, 22848] bf16 streamed once per batch; fp32 partials [B][KS][Hs]
            ProfScope ps(e, st, PK_SLS_FC1, (double)Hs * Kp * 2 + (double)B * Kp * 2 + (double)B * KS * Hs * 4);
            TcGemmArgs g;
            g.a_mode = A_PLAIN; g.A = e->sls_in.p; g.lda = Kp; g.W = W16("sls.fc1.w"); g.ldw = Kp; g.M = B; g.N = Hs; g.K = Kp; g.k_splits = KS;
            g.out = e->sls_part.p; g.ldc = (long long)KS * Hs; g.out_batch_stride = Hs; g.out_bf16 = 0; g.bias = e->zeros.as<float>(); g.act = ACT_NONE;
            LAUNCH(tc_gemm(g, e->num_sms, st));
        } else {
            SimtGemmArgs g;      // split-K: group ks handles columns [ks*Kp/KS, (ks+1)*Kp/KS) of both operands
            g.A = e->sls_in.p; g.lda = Kp; g.a_group_offset = Kp / KS; g.W = W32("sls.fc1.w"); g.ldw = Kp; g.w_group_stride = Kp / KS;
            g.groups = KS; g.n_per_group = Hs; g.M = B; g.N = Hs; g.K = Kp / KS; g.batches = 1;
            g.out = e->sls_part.p; g.ldc = (long long)KS * Hs; g.bias = e->zeros.as<float>(); g.act = ACT_NONE;
            ProfScope ps(e, st, PK_SLS_FC1, (double)Hs * Kp * 4 + (double)B * Kp * 4 + (double)B * KS * Hs * 4);
            LAUNCH(simt_gemm(g, st));
        }
        LAUNCH(sls_tail(e->sls_part.as<float>(), KS, B, Hs, W32("sls.fc1.b"), W32("sls.fc3.w"), W32("sls.fc3.b"), logprob, st));
        return 0;
    }
    set_error("unknown head %d", head);
    return -1;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int slsb_abi_version(void) { return SLSB_ABI_VERSION; }
const char* slsb_last_error(void) { return g_err; }

int slsb_create(const slsb_config* cfg, int device, slsb_engine** out) {
    if (!cfg || !out) { set_error("slsb_create: null argument"); return -1; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { set_error("slsb_create: no CUDA device (this library has no CPU fallback)"); return -2; }
    if (device < 0 || device >= n) { set_error("slsb_create: device %d out of range (%d devices)", device, n); return -1; }
    cudaDeviceProp prop;
    DeviceGuard guard(device);            // allocate the arena on `device`, hand the caller's current device back
    SLSB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_error("slsb_create: device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor); return -2; }
    if (cfg->n_conv < 1 || cfg->n_conv > 8 || cfg->embed_dim % 128 || cfg->embed_dim / cfg->n_heads != 64 || cfg->conv_dim != 512 ||
        cfg->embed_dim / cfg->pos_groups != 64 || cfg->ffn_dim % 256 || cfg->n_layers < 1 || cfg->n_layers > 31) {
        set_error("slsb_create: unsupported geometry (need conv_dim 512, head dim 64, pos group width 64, ffn %% 256 == 0, <= 31 layers)");
        return -1;
    }
    slsb_engine* e = new slsb_engine();
    e->cfg = *cfg;
    e->device = device;
    e->num_sms = prop.multiProcessorCount;
    if (build_weight_table(e)) { slsb_destroy(e); return -1; }
    *out = e;
    return 0;
}

int slsb_destroy(slsb_engine* e) {
    if (!e) return 0;
    DeviceGuard guard(e->device);
    cudaDeviceSynchronize();
    if (e->l2_ptr) {                        // take the persisting-L2 window off the stream it was installed on, release the lines
        cudaStreamAttrValue v{};
        v.accessPolicyWindow.base_ptr = nullptr; v.accessPolicyWindow.num_bytes = 0; v.accessPolicyWindow.hitRatio = 0.f;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal; v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        if (cudaStreamSetAttribute(e->l2_stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
        if (cudaCtxResetPersistingL2Cache() != cudaSuccess) cudaGetLastError();
    }
    for (auto& kv : e->w) { if (kv.second.f32) cudaFree(kv.second.f32); if (kv.second.b16) cudaFree(kv.second.b16); }
    Buf* bufs[] = {&e->fe[0], &e->fe[1], &e->lnbuf, &e->qkv, &e->attn, &e->ffn, &e->xmid, &e->xfinal, &e->xc, &e->xpad, &e->acts, &e->encoded,
                   &e->sums, &e->votes, &e->thr, &e->cut, &e->thr_w, &e->cut_w, &e->pool_part, &e->wmask, &e->flac_bytes, &e->flac_frames, &e->flac_status, &e->pooled, &e->logprob, &e->sls_w, &e->sls_in, &e->sls_part, &e->sls_dots,
                   &e->zeros, &e->scratch, &e->flens, &e->wav_stage[0], &e->wav_stage[1], &e->lens_stage[0], &e->lens_stage[1], &e->score_stage[0],
                   &e->score_stage[1], &e->score_stage[2], &e->score_stage[3], &e->recon, &e->tmp_bf16, &e->im2col, &e->conv0_w64, &e->conv0_wb64, &e->conv0_gram, &e->ybuf,
                   &e->pcm_stage, &e->off_stage};
    for (Buf* b : bufs) b->release();
    for (int i = 0; i < 2; ++i) { if (e->ev_h2d[i]) cudaEventDestroy(e->ev_h2d[i]); if (e->ev_slot_free[i]) cudaEventDestroy(e->ev_slot_free[i]); }
    for (int i = 0; i < 4; ++i) if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto& b : e->X) b.release();
    for (auto& b : e->snap) b.release();
    delete e;
    return 0;
}

int64_t slsb_weight_numel(slsb_engine* e, const char* name) {
    if (!e || !name) return -1;
    auto it = e->w.find(name);
    return it == e->w.end() ? -1 : it->second.numel;
}

int slsb_set_weight(slsb_engine* e, const char* name, const float* src, int64_t numel, void* stream) {
    if (!e || !name || !src) { set_error("slsb_set_weight: null argument"); return -1; }
    DeviceGuard guard(e->device);
    auto it = e->w.find(name);
    if (it == e->w.end()) { set_error("slsb_set_weight: unknown tensor '%s'", name); return -1; }
    if (it->second.numel != numel) { set_error("slsb_set_weight: '%s' expects %lld elements, got %lld", name, (long long)it->second.numel, (long long)numel); return -1; }
    SLSB_CUDA_CHECK(cudaMemcpyAsync(it->second.f32, src, numel * sizeof(float), cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    it->second.set = true;
    e->finalized = false;
    return 0;
}

int slsb_finalize_weights(slsb_engine* e, void* stream) {
    if (!e) { set_error("null engine"); return -1; }
    DeviceGuard guard(e->device);
    for (auto& kv : e->w) {
        if (!kv.second.set) { set_error("slsb_finalize_weights: tensor '%s' was never set", kv.first.c_str()); return -1; }
        if (kv.second.gemm) LAUNCH(convert_f32_to_bf16(kv.second.f32, kv.second.b16, kv.second.numel, static_cast<cudaStream_t>(stream)));
    }
    if (e->conv0_w64.reserve((size_t)e->cfg.conv_dim * 64 * 2)) return -1;
    LAUNCH(conv0_pack_weights(e->w.at("conv0.w").f32, e->conv0_w64.p, e->cfg.conv_dim, e->cfg.conv_kernel[0], static_cast<cudaStream_t>(stream)));
    if (conv0_one_kernel(e->cfg)) {
        if (e->conv0_wb64.reserve((size_t)e->cfg.conv_dim * 64 * 2) || e->conv0_gram.reserve((16 + 256) * 4)) return -1;
        LAUNCH(conv0_tc_pack(e->w.at("conv0.w").f32, e->w.at("conv0.b").f32, e->conv0_wb64.p, e->conv0_gram.as<float>(), e->cfg.conv_dim,
                             e->cfg.conv_kernel[0], static_cast<cudaStream_t>(stream)));
    }
    e->finalized = true;
    return 0;
}

int slsb_frames_for_samples(const slsb_engine* e, int samples) { return e ? conv_len(e->cfg, samples) : -1; }
int64_t slsb_launch_count(const slsb_engine* e) { return e ? e->launches : -1; }

int slsb_forward(slsb_engine* e, const float* wav_dev, const int32_t* sample_lens_dev, int B, int S, int head_flags, int precision,
                 float* logprob_dev, void* stream) {
    if (check_ready(e)) return -1;
    DeviceGuard guard(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int head = head_flags & 0xFF;
    if (head_flags & ~(0xFF | SLSB_HEAD_RETAIN)) { set_error("slsb_forward: unknown head flags 0x%x", head_flags); return -1; }
    if (run_trunk(e, wav_dev, sample_lens_dev, B, S, precision, head_flags, st)) return -1;
    e->head = head;
    if (head == SLSB_HEAD_NONE) return 0;
    if (!logprob_dev) { set_error("slsb_forward: logprob_dev is null"); return -1; }
    return run_head(e, head_flags, precision, logprob_dev, st);
}

int slsb_extract_feat(slsb_engine* e, const float* wav_dev, const int32_t* sample_lens_dev, int B, int S, int precision, float* x_dev, void* stream) {
    if (check_ready(e)) return -1;
    DeviceGuard guard(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (run_trunk(e, wav_dev, sample_lens_dev, B, S, precision, SLSB_HEAD_NONE, st)) return -1;
    e->head = SLSB_HEAD_NONE;
    if (x_dev) SLSB_CUDA_CHECK(cudaMemcpyAsync(x_dev, e->xfinal.p, (size_t)B * e->T * e->cfg.embed_dim * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int slsb_get_tensor(slsb_engine* e, const char* name, float* dst, int64_t numel, void* stream) {
    if (!e || !name || !dst) { set_error("slsb_get_tensor: null argument"); return -1; }
    DeviceGuard guard(e->device);
    if (e->B == 0) { set_error("slsb_get_tensor: no forward has run yet"); return -1; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const slsb_config& c = e->cfg;
    const long long M = (long long)e->B * e->T;
    const std::string n(name);
    const void* src = nullptr;
    int64_t want = 0;
    if (n == "x") { src = e->xfinal.p; want = M * c.embed_dim; }
    else if (n.rfind("layer_results.", 0) == 0) {
        const int i = atoi(n.c_str() + 14);
        if (i < 0 || i >= c.n_layers) { set_error("slsb_get_tensor: layer %d out of range", i); return -1; }
        want = M * c.embed_dim;
        if (!e->have_snap) { set_error("slsb_get_tensor: layer results were not retained by the last forward (pass head | SLSB_HEAD_RETAIN)"); return -1; }
        if (e->inplace) {      // bf16 snapshot of the layer result -> fp32
            if (numel != want) { set_error("slsb_get_tensor: '%s' has %lld elements, caller gave %lld", name, (long long)want, (long long)numel); return -1; }
            ++e->launches;
            bf16_to_f32_kernel<<<(unsigned)((want + 255) / 256), 256, 0, st>>>(e->snap[i].as<bf16>(), dst, want);
            SLSB_CUDA_CHECK(cudaGetLastError());
            return 0;
        }
        src = e->X[i + 1].p;
    } else if (n == "pos_out") {
        if (e->inplace) { set_error("slsb_get_tensor: 'pos_out' is not retained by the in-place stream (SLSB_INPLACE=0 keeps it)"); return -1; }
        src = e->X[0].p; want = M * c.embed_dim;
    }
    else if (n == "acts") { if (!e->have_acts) { set_error("no SAE activations: last forward ran no SAE head"); return -1; } src = e->acts.p; want = M * c.sae_dict; }
    else if (n == "pooled") { src = e->pooled.p; want = (int64_t)e->B * c.cls_in; }
    else if (n == "sls_weights") { src = e->sls_w.p; want = (int64_t)e->B * c.n_layers; }
    else if (n == "encoded") {
        if (!e->have_sel) { set_error(e->have_acts ? "the window head keeps its votes only on request: run the forward with head | SLSB_HEAD_RETAIN" : "no SAE selection: last forward ran no SAE head"); return -1; }
        want = M * c.sae_dict;
        if (numel != want) { set_error("slsb_get_tensor: '%s' has %lld elements, caller gave %lld", name, (long long)want, (long long)numel); return -1; }
        const float* sel = (e->head == SLSB_HEAD_WINDOW && c.sae_window > 1) ? e->votes.as<float>() : e->acts.as<float>();
        LAUNCH(votes_densify(e->acts.as<float>(), sel, e->thr.as<float>(), e->cut.as<int>(), dst, M, c.sae_dict, st));
        return 0;
    } else if (n == "features") {
        want = M * c.conv_dim;
        if (numel != want) { set_error("slsb_get_tensor: '%s' has %lld elements, caller gave %lld", name, (long long)want, (long long)numel); return -1; }
        set_error("slsb_get_tensor: 'features' is overwritten by the encoder's LayerNorm scratch; not retained");
        return -1;
    } else { set_error("slsb_get_tensor: unknown tensor '%s'", name); return -1; }
    if (numel != want) { set_error("slsb_get_tensor: '%s' has %lld elements, caller gave %lld", name, (long long)want, (long long)numel); return -1; }
    SLSB_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)want * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int slsb_get_sparse(slsb_engine* e, int32_t* idx_dev, float* val_dev, int32_t* count_dev, void* stream) {
    if (!e || !idx_dev || !val_dev) { set_error("slsb_get_sparse: null argument"); return -1; }
    DeviceGuard guard(e->device);
    if (!e->have_sel) { set_error(e->have_acts ? "slsb_get_sparse: the window head keeps its votes only on request (head | SLSB_HEAD_RETAIN)" : "slsb_get_sparse: last forward ran no SAE head"); return -1; }
    const slsb_config& c = e->cfg;
    const long long M = (long long)e->B * e->T;
    const float* sel = (e->head == SLSB_HEAD_WINDOW && c.sae_window > 1) ? e->votes.as<float>() : e->acts.as<float>();
    LAUNCH(votes_compact(e->acts.as<float>(), sel, e->thr.as<float>(), e->cut.as<int>(), idx_dev, val_dev, count_dev, M, c.sae_dict, c.sae_k,
                         static_cast<cudaStream_t>(stream)));
    return 0;
}

int slsb_sae_encode(slsb_engine* e, const float* x_dev, int64_t rows, int T, int window, int precision, float* encoded_dev, void* stream) {
    if (check_ready(e)) return -1;
    DeviceGuard guard(e->device);
    const slsb_config& c = e->cfg;
    if (c.sae_dict <= 0) { set_error("engine has no SAE weights"); return -1; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool bf = precision == SLSB_PREC_BF16;
    if (rows <= 0) return 0;
    if (e->xc.reserve((size_t)rows * c.embed_dim * (bf ? 2 : 4))) return -1;
    const long long n = rows * c.embed_dim;
    ++e->launches;
    center_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x_dev, W32("sae.b_dec"), bf ? nullptr : e->xc.as<float>(), bf ? e->xc.as<bf16>() : nullptr, n, c.embed_dim);
    SLSB_CUDA_CHECK(cudaGetLastError());
    if (sae_acts(e, bf, e->xc.p, rows, st)) return -1;
    const float* sel = nullptr;
    if (sae_select(e, rows, T > 0 ? T : 1, window, true, nullptr, nullptr, &sel, st)) return -1;
    LAUNCH(votes_densify(e->acts.as<float>(), sel, e->thr.as<float>(), e->cut.as<int>(), encoded_dev, rows, c.sae_dict, st));
    return 0;
}

int slsb_sae_decode(slsb_engine* e, const float* encoded_dev, int64_t rows, int precision, float* recon_dev, void* stream) {
    if (check_ready(e)) return -1;
    DeviceGuard guard(e->device);
    const slsb_config& c = e->cfg;
    if (c.sae_dict <= 0) { set_error("engine has no SAE weights"); return -1; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool bf = precision == SLSB_PREC_BF16;
    if (rows <= 0) return 0;
    const void* A = encoded_dev;
    if (bf) {
        if (e->tmp_bf16.reserve((size_t)rows * c.sae_dict * 2)) return -1;
        LAUNCH(convert_f32_to_bf16(encoded_dev, e->tmp_bf16.p, rows * c.sae_dict, st));
        A = e->tmp_bf16.p;
    }
    return linear(e, bf, A, c.sae_dict, "sae.dec.w", c.embed_dim, c.sae_dict, rows, W32("sae.b_dec"), nullptr, 0, recon_dev, c.embed_dim, 0, ACT_NONE, st);
}

int slsb_sae_loss(slsb_engine* e, int precision, float* loss_dev, void* stream) {
    if (check_ready(e)) return -1;
    DeviceGuard guard(e->device);
    const slsb_config& c = e->cfg;
    if (!e->have_sel) { set_error(e->have_acts ? "slsb_sae_loss: the window head keeps its votes only on request (head | SLSB_HEAD_RETAIN)" : "slsb_sae_loss: last forward ran no SAE head"); return -1; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long M = (long long)e->B * e->T;
    if (e->encoded.reserve((size_t)M * c.sae_dict * 4) || e->recon.reserve((size_t)M * c.embed_dim * 4) || e->scratch.reserve((size_t)(e->B * c.cls_hidden + 1024) * 4)) return -1;
    const float* sel = (e->head == SLSB_HEAD_WINDOW && c.sae_window > 1) ? e->votes.as<float>() : e->acts.as<float>();
    LAUNCH(votes_densify(e->acts.as<float>(), sel, e->thr.as<float>(), e->cut.as<int>(), e->encoded.as<float>(), M, c.sae_dict, st));
    if (slsb_sae_decode(e, e->encoded.as<float>(), M, precision, e->recon.as<float>(), stream)) return -1;
    e->launches += 1;
    LAUNCH(mse_loss(e->recon.as<float>(), e->xfinal.as<float>(), M * c.embed_dim, loss_dev, e->scratch.as<float>(), st));
    return 0;
}

int64_t slsb_score_submit(slsb_engine* e, const float* wav_host, const int32_t* lens_host, int B, int S, int head, int precision,
                          float* scores_host, void* stream) {
    if (check_ready(e)) return -1;
    DeviceGuard guard(e->device);
    if (!wav_host || !scores_host) { set_error("slsb_score_submit: null buffer"); return -1; }
    if ((head & 0xFF) == SLSB_HEAD_NONE) { set_error("slsb_score_submit: a classifier head is required"); return -1; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!e->copy_stream) {
        SLSB_CUDA_CHECK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            SLSB_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_h2d[i], cudaEventDisableTiming));
            SLSB_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_slot_free[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 4; ++i) SLSB_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
    }
    const int64_t ticket = e->submit_seq;
    const int slot = (int)(ticket & 1), ring = (int)(ticket & 3);
    if (ticket >= 4) SLSB_CUDA_CHECK(cudaEventSynchronize(e->ev_done[ring]));      // bound the submissions in flight (scores_host reuse is the caller's business)
    if (e->wav_stage[slot].reserve((size_t)B * S * 4) || e->lens_stage[slot].reserve((size_t)B * 4) || e->logprob.reserve((size_t)B * 2 * 4) ||
        e->score_stage[ring].reserve((size_t)B * 4)) return -1;
    // upload on the copy stream once the forward that last read this slot has finished
    if (ticket >= 2) SLSB_CUDA_CHECK(cudaStreamWaitEvent(e->copy_stream, e->ev_slot_free[slot], 0));
    SLSB_CUDA_CHECK(cudaMemcpyAsync(e->wav_stage[slot].p, wav_host, (size_t)B * S * 4, cudaMemcpyHostToDevice, e->copy_stream));
    const int32_t* lens_dev = nullptr;
    if (lens_host) {
        SLSB_CUDA_CHECK(cudaMemcpyAsync(e->lens_stage[slot].p, lens_host, (size_t)B * 4, cudaMemcpyHostToDevice, e->copy_stream));
        lens_dev = e->lens_stage[slot].as<int32_t>();
    }
    SLSB_CUDA_CHECK(cudaEventRecord(e->ev_h2d[slot], e->copy_stream));
    SLSB_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev_h2d[slot], 0));
    if (slsb_forward(e, e->wav_stage[slot].as<float>(), lens_dev, B, S, head, precision, e->logprob.as<float>(), stream)) return -1;
    SLSB_CUDA_CHECK(cudaEventRecord(e->ev_slot_free[slot], st));
    LAUNCH(scores_from_logprob(e->logprob.as<float>(), e->score_stage[ring].as<float>(), B, st));
    SLSB_CUDA_CHECK(cudaMemcpyAsync(scores_host, e->score_stage[ring].p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    SLSB_CUDA_CHECK(cudaEventRecord(e->ev_done[ring], st));
    e->submit_seq = ticket + 1;
    return ticket;
}

int slsb_score_wait(slsb_engine* e, int64_t ticket) {
    if (!e) { set_error("null engine"); return -1; }
    DeviceGuard guard(e->device);
    if (ticket >= e->submit_seq) { set_error("slsb_score_wait: ticket %lld was never submitted", (long long)ticket); return -1; }
    const int64_t lo = ticket < 0 ? (e->submit_seq > 4 ? e->submit_seq - 4 : 0) : ticket;
    const int64_t hi = ticket < 0 ? e->submit_seq : ticket + 1;
    for (int64_t t = lo; t < hi; ++t) {
        if (t + 4 < e->submit_seq) continue;                 // its ring slot was re-used, which already waited for it
        SLSB_CUDA_CHECK(cudaEventSynchronize(e->ev_done[t & 3]));
    }
    return 0;
}

int slsb_score_host(slsb_engine* e, const float* wav_host, const int32_t* lens_host, int B, int S, int head, int precision, float* scores_host, void* stream) {
    const int64_t t = slsb_score_submit(e, wav_host, lens_host, B, S, head, precision, scores_host, stream);
    if (t < 0) return -1;
    return slsb_score_wait(e, t);
}

int slsb_ingest_pcm16(const int16_t* pcm_dev, const int64_t* offsets_dev, const int32_t* lens_dev, int B, int S, float* wav_dev, void* stream) {
    if (!pcm_dev || !offsets_dev || !lens_dev || !wav_dev) { set_error("slsb_ingest_pcm16: null buffer"); return -1; }
    return ingest_pcm16(pcm_dev, reinterpret_cast<const long long*>(offsets_dev), lens_dev, B, S, wav_dev, static_cast<cudaStream_t>(stream));
}

int slsb_score_pcm16_host(slsb_engine* e, const int16_t* pcm_host, int64_t total_samples, const int64_t* offsets_host, const int32_t* lens_host,
                          int B, int S, int head, int precision, float* scores_host, void* stream) {
    if (check_ready(e)) return -1;
    DeviceGuard guard(e->device);
    if (!pcm_host || !offsets_host || !lens_host || !scores_host) { set_error("slsb_score_pcm16_host: null buffer"); return -1; }
    if ((head & 0xFF) == SLSB_HEAD_NONE) { set_error("slsb_score_pcm16_host: a classifier head is required"); return -1; }
    for (int b = 0; b < B; ++b) {
        if (lens_host[b] < 1 || offsets_host[b] < 0 || offsets_host[b] + lens_host[b] > total_samples) {
            set_error("slsb_score_pcm16_host: clip %d (offset %lld, %d samples) is empty or outside the %lld-sample buffer", b,
                      (long long)offsets_host[b], lens_host[b], (long long)total_samples);
            return -1;
        }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (e->pcm_stage.reserve((size_t)total_samples * 2) || e->off_stage.reserve((size_t)B * 8) || e->lens_stage[0].reserve((size_t)B * 4) ||
        e->wav_stage[0].reserve((size_t)B * S * 4) || e->logprob.reserve((size_t)B * 2 * 4) || e->score_stage[0].reserve((size_t)B * 4)) return -1;
    SLSB_CUDA_CHECK(cudaMemcpyAsync(e->pcm_stage.p, pcm_host, (size_t)total_samples * 2, cudaMemcpyHostToDevice, st));
    SLSB_CUDA_CHECK(cudaMemcpyAsync(e->off_stage.p, offsets_host, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    SLSB_CUDA_CHECK(cudaMemcpyAsync(e->lens_stage[0].p, lens_host, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    LAUNCH(ingest_pcm16(e->pcm_stage.as<int16_t>(), e->off_stage.as<long long>(), e->lens_stage[0].as<int>(), B, S, e->wav_stage[0].as<float>(), st));
    // pad() has already brought every clip to S samples (the reference's eval path never passes lengths down)
    if (slsb_forward(e, e->wav_stage[0].as<float>(), nullptr, B, S, head, precision, e->logprob.as<float>(), stream)) return -1;
    LAUNCH(scores_from_logprob(e->logprob.as<float>(), e->score_stage[0].as<float>(), B, st));
    SLSB_CUDA_CHECK(cudaMemcpyAsync(scores_host, e->score_stage[0].p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    SLSB_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
}

int slsb_score_flac_host(slsb_engine* e, const uint8_t* bytes_host, int64_t nbytes, const slsb_flac_frame* frames_host, int n_frames,
                         int64_t total_samples, const int64_t* offsets_host, const int32_t* lens_host, int B, int S, int head, int precision,
                         float* scores_host, int32_t* status_host, void* stream) {
    if (check_ready(e)) return -1;
    DeviceGuard guard(e->device);
    if (!bytes_host || !frames_host || !offsets_host || !lens_host || !scores_host || !status_host) { set_error("slsb_score_flac_host: null buffer"); return -1; }
    if ((head & 0xFF) == SLSB_HEAD_NONE) { set_error("slsb_score_flac_host: a classifier head is required"); return -1; }
    if (n_frames < 1 || nbytes < 1) { set_error("slsb_score_flac_host: empty batch"); return -1; }
    for (int b = 0; b < B; ++b) {
        if (lens_host[b] < 1 || offsets_host[b] < 0 || offsets_host[b] + lens_host[b] > total_samples) {
            set_error("slsb_score_flac_host: clip %d (offset %lld, %d samples) is empty or outside the %lld-sample buffer", b,
                      (long long)offsets_host[b], lens_host[b], (long long)total_samples);
            return -1;
        }
    }
    for (int i = 0; i < n_frames; ++i) {
        const slsb_flac_frame& f = frames_host[i];
        if (f.byte_off < 0 || f.byte_len < 1 || f.byte_off + f.byte_len > nbytes || f.keep < 0 || f.out_off < 0 || f.out_off + f.keep > total_samples) {
            set_error("slsb_score_flac_host: frame %d lies outside its buffers", i);
            return -1;
        }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (e->flac_bytes.reserve((size_t)nbytes + 8) || e->flac_frames.reserve((size_t)n_frames * sizeof(slsb_flac_frame)) ||
        e->flac_status.reserve((size_t)n_frames * 4) || e->pcm_stage.reserve((size_t)total_samples * 2) || e->off_stage.reserve((size_t)B * 8) ||
        e->lens_stage[0].reserve((size_t)B * 4) || e->wav_stage[0].reserve((size_t)B * S * 4) || e->logprob.reserve((size_t)B * 2 * 4) ||
        e->score_stage[0].reserve((size_t)B * 4)) return -1;
    SLSB_CUDA_CHECK(cudaMemcpyAsync(e->flac_bytes.p, bytes_host, (size_t)nbytes, cudaMemcpyHostToDevice, st));
    SLSB_CUDA_CHECK(cudaMemsetAsync(static_cast<uint8_t*>(e->flac_bytes.p) + nbytes, 0, 8, st));          // the word reader may touch the next 4-byte boundary
    SLSB_CUDA_CHECK(cudaMemcpyAsync(e->flac_frames.p, frames_host, (size_t)n_frames * sizeof(slsb_flac_frame), cudaMemcpyHostToDevice, st));
    SLSB_CUDA_CHECK(cudaMemcpyAsync(e->off_stage.p, offsets_host, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    SLSB_CUDA_CHECK(cudaMemcpyAsync(e->lens_stage[0].p, lens_host, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    LAUNCH(flac_decode_frames(e->flac_bytes.as<uint8_t>(), e->flac_frames.as<slsb_flac_frame>(), n_frames, e->pcm_stage.as<int16_t>(),
                              e->flac_status.as<int32_t>(), st));
    SLSB_CUDA_CHECK(cudaMemcpyAsync(status_host, e->flac_status.p, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, st));
    LAUNCH(ingest_pcm16(e->pcm_stage.as<int16_t>(), e->off_stage.as<long long>(), e->lens_stage[0].as<int>(), B, S, e->wav_stage[0].as<float>(), st));
    if (slsb_forward(e, e->wav_stage[0].as<float>(), nullptr, B, S, head, precision, e->logprob.as<float>(), stream)) return -1;
    LAUNCH(scores_from_logprob(e->logprob.as<float>(), e->score_stage[0].as<float>(), B, st));
    SLSB_CUDA_CHECK(cudaMemcpyAsync(scores_host, e->score_stage[0].p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    SLSB_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int i = 0; i < n_frames; ++i) {
        if (status_host[i] < 0) {       // the scores of this call are void: the caller decodes the batch on the host (slsb_score_pcm16_host)
            set_error("slsb_score_flac_host: frame %d was refused by the device decoder (code %d)", i, status_host[i]);
            return -2;
        }
    }
    return 0;
}

int slsb_profile_enable(slsb_engine* e, int on) {
    if (!e) { set_error("null engine"); return -1; }
    DeviceGuard guard(e->device);
    for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    e->prof.clear();
    e->profiling = on != 0;
    return 0;
}

int slsb_profile_read(slsb_engine* e, int kind, double* ms_out, double* flops_out, int64_t* launches_out) {
    if (!e) { set_error("null engine"); return -1; }
    DeviceGuard guard(e->device);
    SLSB_CUDA_CHECK(cudaDeviceSynchronize());
    double ms = 0.0, fl = 0.0; int64_t n = 0;
    for (auto& r : e->prof) {
        if (r.kind != kind) continue;
        float t = 0.f;
        SLSB_CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
        ms += t; fl += r.flops; ++n;
    }
    if (ms_out) *ms_out = ms;
    if (flops_out) *flops_out = fl;
    if (launches_out) *launches_out = n;
    return 0;
}

int slsb_synth_clips(float* wav_dev, int64_t first_utt, int count, int samples, void* stream) {
    return synth_clips(wav_dev, first_utt, count, samples, static_cast<cudaStream_t>(stream));
}

// ---- single-op entry points -------------------------------------------------------------------
static int device_sms() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

int slsb_op_gemm(int precision, const void* A, const void* W, const float* bias, const float* residual, void* out, int M, int N, int K,
                 int act, int out_bf16, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (precision == SLSB_PREC_BF16) {
        TcGemmArgs g;
        g.A = A; g.lda = K; g.W = W; g.ldw = K; g.M = M; g.N = N; g.K = K; g.out = out; g.ldc = N; g.out_bf16 = out_bf16;
        g.bias = bias; g.residual = residual; g.ldr = N; g.act = act;
        return tc_gemm(g, device_sms(), st);
    }
    SimtGemmArgs g;
    g.A = A; g.lda = K; g.W = static_cast<const float*>(W); g.ldw = K; g.M = M; g.N = N; g.K = K; g.out = out; g.ldc = N; g.out_bf16 = out_bf16;
    g.bias = bias; g.residual = residual; g.ldr = N; g.act = act;
    return simt_gemm(g, st);
}

int slsb_op_gemm_splitk(const void* A, const void* W, float* partial, int M, int N, int K, int k_splits, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* zeros = nullptr;
    SLSB_CUDA_CHECK(cudaMalloc(&zeros, (size_t)N * 4));
    SLSB_CUDA_CHECK(cudaMemsetAsync(zeros, 0, (size_t)N * 4, st));
    TcGemmArgs g;
    g.A = A; g.lda = K; g.W = W; g.ldw = K; g.M = M; g.N = N; g.K = K; g.k_splits = k_splits;
    g.out = partial; g.ldc = (long long)k_splits * N; g.out_batch_stride = N; g.out_bf16 = 0; g.bias = zeros; g.act = ACT_NONE;
    const int rc = tc_gemm(g, device_sms(), st);
    cudaStreamSynchronize(st);
    cudaFree(zeros);
    return rc;
}

int slsb_op_conv(int precision, const void* x, const void* W, const float* bias, void* out, int B, int L_in, int C, int N, int k, int stride, void* stream) {
    slsb_engine tmp;      // only num_sms / launches are touched
    tmp.num_sms = device_sms();
    return conv_layer(&tmp, precision == SLSB_PREC_BF16, x, W, bias, out, B, L_in, C, N, k, stride, static_cast<cudaStream_t>(stream));
}

int slsb_op_conv_ln_gelu(const void* x, const void* W, const float* bias, const float* ln_w, const float* ln_b, void* out, int B, int L_in,
                         int C, int k, int stride, void* stream) {
    TcLnGemmArgs g;
    const int Lout = (L_in - k) / stride + 1;
    g.a_mode = A_CONV; g.A = x; g.W = W; g.M = Lout; g.K = k * C; g.batches = B; g.conv_cin = C; g.conv_stride = stride; g.conv_lin = L_in;
    g.out = out; g.out_batch_stride = (long long)Lout * 512; g.bias = bias; g.ln_w = ln_w; g.ln_b = ln_b;
    return ln_gemm_dispatch(g, device_sms(), static_cast<cudaStream_t>(stream));
}

int slsb_op_conv0_tc(const float* wav, const float* w, const float* bias, const float* ln_w, const float* ln_b, void* out, void* scratch,
                     int B, int S, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int L0 = (S - 10) / 5 + 1;
    char* w64 = static_cast<char*>(scratch);                       // [512, 64] bf16
    char* cols = w64 + 512 * 64 * 2;                               // [B*L0, 64] bf16
    if (conv0_pack_weights(w, w64, 512, 10, st)) return -1;
    if (conv0_im2col(wav, cols, B, S, L0, 10, 5, st)) return -1;
    TcLnGemmArgs g;
    g.a_mode = A_PLAIN; g.A = cols; g.lda = 64; g.W = w64; g.M = B * L0; g.K = 64; g.batches = 1;
    g.out = out; g.bias = bias; g.ln_w = ln_w; g.ln_b = ln_b;
    return ln_gemm_dispatch(g, device_sms(), st);
}

int slsb_op_conv0_fused(const float* wav, const float* w, const float* bias, const float* ln_w, const float* ln_b, void* out, void* scratch,
                        int B, int S, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int L0 = (S - 10) / 5 + 1;
    char* w64 = static_cast<char*>(scratch);                       // [512, 64] bf16
    float* gram = reinterpret_cast<float*>(w64 + 512 * 64 * 2);    // 272 floats
    if (conv0_tc_pack(w, bias, w64, gram, 512, 10, st)) return -1;
    return conv0_tc(wav, w64, gram, ln_w, ln_b, out, B, S, L0, 10, 5, 1e-5f, device_sms(), st);
}

int slsb_op_posconv(int precision, const float* x, const void* W, const float* bias, float* out, void* scratch, int B, int T, int D, int K,
                    const int32_t* frame_lens_dev, void* stream) {
    slsb_engine tmp;
    tmp.num_sms = device_sms();
    return pos_conv(&tmp, precision == SLSB_PREC_BF16, x, W, bias, out, scratch, B, T, D, K, D / 64, frame_lens_dev, static_cast<cudaStream_t>(stream));
}

int slsb_op_conv0(int out_bf16, const float* wav, const float* w, const float* bias, const float* ln_w, const float* ln_b, void* out, int B, int S,
                  int exact_gelu, void* stream) {
    const int L0 = (S - 10) / 5 + 1;
    return conv0_ln_gelu(wav, B, S, L0, 10, 5, w, bias, ln_w, ln_b, out, out_bf16, 512, exact_gelu != 0, static_cast<cudaStream_t>(stream));
}

int slsb_op_layernorm(const void* in, int in_bf16, void* out, int out_bf16, const float* w, const float* b, int64_t rows, int C, int gelu,
                      int exact_gelu, void* stream) {
    LnArgs a;
    a.in = in; a.in_bf16 = in_bf16; a.out = out; a.out_bf16 = out_bf16; a.w = w; a.b = b; a.rows = rows; a.C = C; a.gelu = gelu; a.exact_gelu = exact_gelu;
    return layernorm(a, static_cast<cudaStream_t>(stream));
}

int slsb_debug_pair_schedule(int M, int N, int num_pairs, int32_t* items, int max_items, int32_t* split_out) {
    if (!items && max_items > 0) { set_error("slsb_debug_pair_schedule: null items"); return -1; }
    return pair_schedule(M, N, num_pairs, items, max_items, split_out);
}

int slsb_op_layernorm_taps(const float* in, void* out_bf16, const float* w, const float* b, const float* dot_w, float* dot_out,
                           void* copy_out_bf16, int64_t rows, int C, void* stream) {
    LnArgs a;
    a.in = in; a.in_bf16 = 0; a.out = out_bf16; a.out_bf16 = 1; a.w = w; a.b = b; a.rows = rows; a.C = C;
    a.dot_w = dot_w; a.dot_out = dot_out; a.copy_out = copy_out_bf16;
    if (!in || !out_bf16 || !w || !b || (dot_out && !dot_w)) { set_error("slsb_op_layernorm_taps: null buffer"); return -1; }
    return layernorm(a, static_cast<cudaStream_t>(stream));
}

int slsb_op_attention(int impl, int io_bf16, const void* qkv, void* out, int B, int T, int H, const int32_t* frame_lens_dev, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (impl == SLSB_ATTN_TC || impl == SLSB_ATTN_TC_V1) {
        if (!io_bf16) { set_error("attention_tc is bf16 only"); return -1; }
        if (impl == SLSB_ATTN_TC_V1) return attention_tc_v1(qkv, out, B, T, H, frame_lens_dev, device_sms(), st);
        return attention_tc(qkv, out, B, T, H, frame_lens_dev, device_sms(), st);
    }
    return attention_simt(qkv, out, io_bf16, B, T, H, frame_lens_dev, st);
}

int slsb_op_attention_trace(const void* qkv, void* out, int B, int T, int H, int64_t* trace_dev, void* stream) {
    return attention_tc(qkv, out, B, T, H, nullptr, device_sms(), static_cast<cudaStream_t>(stream), reinterpret_cast<long long*>(trace_dev));
}

int slsb_op_topk(const float* x, int64_t rows, int D, int k, float* thr, int32_t* tie_cut, float* encoded_or_null, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (topk_threshold(x, rows, D, k, thr, tie_cut, st)) return -1;
    if (encoded_or_null) return topk_densify(x, thr, tie_cut, encoded_or_null, rows, D, st);
    return 0;
}

int slsb_op_topk_pool(const float* acts, int B, int T, int D, int k, const int32_t* lens, float* thr, int32_t* tie_cut, float* partial, float* pooled,
                      void* stream) {
    return topk_select_pool(acts, B, T, D, k, lens, thr, tie_cut, partial, pooled, static_cast<cudaStream_t>(stream));
}

int slsb_op_window_pool(const float* acts, int B, int T, int D, int k, int window, uint32_t* wmask, float* thr, int32_t* tie_cut, float* votes_or_null,
                        float* partial, float* pooled, void* stream) {
    const int stride = window / 2 > 0 ? window / 2 : 1;
    if (window < 2 || T < window || stride >= T) { set_error("slsb_op_window_pool: need 2 <= window <= T"); return -1; }
    const int nw = (T - window) / stride + 1;
    return window_select_pool(acts, B, T, D, k, window, stride, nw, wmask, thr, tie_cut, votes_or_null, partial, pooled, static_cast<cudaStream_t>(stream));
}

int slsb_op_pool_chunks(int T) { return sel_chunks(T); }

}  // extern "C"

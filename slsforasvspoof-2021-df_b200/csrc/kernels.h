// Host-side launchers of every kernel in the library (internal header; the public C ABI is include/slsb200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slsb {

enum AMode : int { A_PLAIN = 0, A_CONV = 1, A_POS = 2, A_POS4 = 3 };

// ---------------------------------------------------------------- tcgen05 bf16 GEMM (gemm_tc.cu)
struct TcGemmArgs {
    int a_mode = A_PLAIN;
    const void* A = nullptr;      // bf16
    long long lda = 0;            // A_PLAIN: row stride in elements
    const void* W = nullptr;      // bf16 [N, K], row stride ldw
    long long ldw = 0;
    int M = 0, N = 0, K = 0;      // M = rows per batch item
    int batches = 1;
    int k_splits = 1;             // A_PLAIN only: split s computes k-blocks [s*ceil(K/64/k_splits), ...) into out + s*out_batch_stride
    // A_CONV
    int conv_cin = 0, conv_stride = 0, conv_lin = 0;
    // A_POS
    int pos_dim = 0, pos_tp = 0;
    // epilogue
    void* out = nullptr; long long ldc = 0, out_batch_stride = 0; int out_bf16 = 0;
    const float* bias = nullptr;
    const float* residual = nullptr; long long ldr = 0, res_batch_stride = 0;
    int act = 0;
};
int tc_gemm(const TcGemmArgs& g, int num_sms, cudaStream_t stream);
// host replay of the CTA-pair kernel's static tile schedule: items[5 * i] = {pair, round, row0, col0, cols}; returns the item count
int pair_schedule(int M, int N, int num_pairs, int* items, int max_items, int* split_out);

// ---------------------------------------------------------------- tcgen05 GEMM + LayerNorm(512) + GELU epilogue (gemm_tc_ln.cu)
struct TcLnGemmArgs {
    int a_mode = A_PLAIN;
    const void* A = nullptr; long long lda = 0;
    const void* W = nullptr;                 // bf16 [512, K]
    int M = 0, N = 512, K = 0, batches = 1;
    int conv_cin = 0, conv_stride = 0, conv_lin = 0;
    void* out = nullptr; long long out_batch_stride = 0;   // bf16 [.., 512]
    const float* bias = nullptr; const float* ln_w = nullptr; const float* ln_b = nullptr;
    float eps = 1e-5f;
};
int tc_gemm_ln_gelu(const TcLnGemmArgs& g, int num_sms, cudaStream_t stream);
// CTA-pair version (gemm_tc_ln2.cu): each CTA of a 2-CTA cluster owns 256 of the 512 channels, row statistics are exchanged
// through distributed shared memory, so two accumulators fit in TMEM and the epilogue overlaps the next tile's MMAs
int tc_gemm_ln_gelu_pair(const TcLnGemmArgs& g, int num_sms, cudaStream_t stream);
// the same with 16 epilogue warps of 64 columns each (gemm_tc_ln2x.cu): the 8-warp epilogue, not the operand feed, bounded the pair kernel
int tc_gemm_ln_gelu_pair16(const TcLnGemmArgs& g, int num_sms, cudaStream_t stream);
// conv0 on tensor cores: raw audio -> hi/lo-split bf16 im2col rows [B*L0, 64]; weights [C, k] -> [C, 64] (hi | lo | hi | 0)
int conv0_im2col(const float* wav, void* out, int B, int S, int L0, int k, int stride, cudaStream_t stream);
int conv0_pack_weights(const float* w, void* out, int C, int k, cudaStream_t stream);
// conv0 as one kernel (conv0_tc.cu): audio -> bf16 [B*L0, 512]; row statistics from the (k+1) x (k+1) Gram matrix of [w | b], bias in
// spare K columns, weights resident in shared memory, A tiles built in shared memory (no im2col buffer).
// w64: [512, 64] bf16; gram: 272 floats (s[16] | G[16][16])
int conv0_tc_pack(const float* w, const float* bias, void* w64, float* gram, int C, int k, cudaStream_t stream);
int conv0_tc(const float* wav, const void* w64, const float* gram, const float* ln_w, const float* ln_b, void* out, int B, int S, int L0, int k,
             int stride, float eps, int num_sms, cudaStream_t stream);

// ---------------------------------------------------------------- fp32 SIMT GEMM (gemm_simt.cu)
// C[z][m][n] = act(sum_k A[z][m][k] * W[zw][n][k] + bias[n_off + n]) (+ residual); A/W/out element types selectable.
struct SimtGemmArgs {
    const void* A = nullptr; int a_bf16 = 0;
    long long lda = 0, a_batch_stride = 0;     // elements
    int a_kinner = 0; long long a_kouter = 0;  // k -> (k / kinner) * kouter + k % kinner   (kinner = 0: contiguous)
    const float* W = nullptr; long long ldw = 0, w_group_stride = 0;   // fp32 weights
    int groups = 1;                             // z = batch * groups + group ; group selects W block, A column offset, out column offset
    long long a_group_offset = 0; int n_per_group = 0;
    int M = 0, N = 0, K = 0, batches = 1;
    void* out = nullptr; int out_bf16 = 0; long long ldc = 0, out_batch_stride = 0;
    const float* bias = nullptr;
    const float* residual = nullptr; long long ldr = 0, res_batch_stride = 0;
    int act = 0; int exact_gelu = 1;
};
int simt_gemm(const SimtGemmArgs& g, cudaStream_t stream);

// ---------------------------------------------------------------- feature extractor / norms (frontend.cu)
// conv0 (C_in = 1) + LayerNorm(512) + GELU fused, channels-last output [B, L0, 512]
int conv0_ln_gelu(const float* wav, int B, int S, int L0, int k, int stride, const float* w /*[C,k]*/, const float* bias,
                  const float* ln_w, const float* ln_b, void* out, int out_bf16, int C, bool exact_gelu, cudaStream_t stream);
// row-wise LayerNorm over C (C in {512, 1024}), optional GELU, optional second output (y - sub[c]); in/out fp32 or bf16
struct LnArgs {
    const void* in = nullptr; int in_bf16 = 0;
    const void* add = nullptr; float* sum_out = nullptr;   // optional: rows = in + add (bf16), written to sum_out (fp32) before the norm
    void* out = nullptr; int out_bf16 = 0;
    void* out2 = nullptr; int out2_bf16 = 0; const float* sub = nullptr;   // out2 = y - sub
    const float* w = nullptr; const float* b = nullptr;
    const float* dot_w = nullptr; float* dot_out = nullptr;   // optional: dot_out[row] = sum_c (in + add)[row, c] * dot_w[c]
    void* copy_out = nullptr;                                 // optional: bf16 copy of the (in + add) rows (layer-result snapshot)
    long long rows = 0; int C = 0; int gelu = 0; int exact_gelu = 1; float eps = 1e-5f;
};
int layernorm(const LnArgs& a, cudaStream_t stream);
// x[B, T, D] (fp32) -> zero-padded [B, T + K, D] (fp32 or bf16), `left` zero frames in front; frames >= lens[b] zeroed
int pad_frames(const float* x, void* out, int out_bf16, int B, int T, int D, int left, int Tp, const int* lens, cudaStream_t stream);
// zero frames t >= lens[b] of x[B, T, D] in place (wav2vec2.py:912-913)
int zero_padded_frames(float* x, int B, int T, int D, const int* lens, cudaStream_t stream);
int convert_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t stream);
// deterministic synthetic clips (same integer hash as oracle.trunk.hash_normal)
int synth_clips(float* out, long long first_utt, int count, int samples, cudaStream_t stream);
// 16-bit PCM clips (concatenated, offsets[b] / lens[b]) -> fp32 [B, S] in [-1, 1), truncated or tile-repeated to S samples
// (data_utils_SSL.py:58-65 pad, :109-115 __getitem__)
int ingest_pcm16(const int16_t* pcm, const long long* offsets, const int* lens, int B, int S, float* out, cudaStream_t stream);
// frame lengths from sample lengths (wav2vec2.py:523-538)
int frame_lengths(const int* sample_lens, int* frame_lens, int B, int n_conv, const int* k, const int* s, cudaStream_t stream);

// ---------------------------------------------------------------- attention (attention.cu)
// qkv [B*T, 3*D] (q pre-scaled), heads of 64; out [B*T, D]; keys >= lens[b] masked (lens may be null)
int attention_simt(const void* qkv, void* out, int io_bf16, int B, int T, int H, const int* lens, cudaStream_t stream);
// trace (optional, device): 64 units x 16 clock64 stamps of CTA 0's pipeline events (slsb_op_attention_trace)
int attention_tc(const void* qkv_bf16, void* out_bf16, int B, int T, int H, const int* lens, int num_sms, cudaStream_t stream,
                 long long* trace = nullptr);
int attention_tc_v1(const void* qkv_bf16, void* out_bf16, int B, int T, int H, const int* lens, int num_sms, cudaStream_t stream);

// ---------------------------------------------------------------- heads (heads.cu)
// per-row top-k threshold of non-negative fp32 rows: writes thr[row] (k-th largest value) and tie_cut[row]
// (entries == thr are kept only while index < tie_cut[row]) -- canonical lowest-index-wins rule
int topk_threshold(const float* acts, long long rows, int D, int k, float* thr, int* tie_cut, cudaStream_t stream);
// dense encoded[row][f] = keep ? acts : 0 (API parity with AutoEncoderTopK.encode, model.py:68-79)
int topk_densify(const float* acts, const float* thr, const int* tie_cut, float* encoded, long long rows, int D, cudaStream_t stream);
// pooled[b][f] = (1/len_b) * sum_{t < len_b} kept(acts[b,t,f])   (model.py:245), fixed summation order
int topk_mean_pool(const float* acts, const float* thr, const int* tie_cut, float* pooled, int B, int T, int D, const int* lens, cudaStream_t stream);
// Fused scoring paths (activations read once per selection; canonical pooling order = chunks of 8 frames, see heads.cu):
//   rows = B * T rows of `acts`; thr / tie_cut as topk_threshold; partial: [B, sel_chunks(T), D] scratch; pooled: [B, D] mean of the
//   kept activations over the frames < len_b.  partial == pooled == nullptr: selection only.
int sel_chunks(int T);
int topk_select_pool(const float* acts, int B, int T, int D, int k, const int* lens, float* thr, int* tie_cut, float* partial, float* pooled,
                     cudaStream_t stream);
//   window variant: wmask [B, nw, 256] words of scratch; votes_or_null [B*T, D] is written only when the caller wants the votes
int window_select_pool(const float* acts, int B, int T, int D, int k, int window, int stride, int nw, uint32_t* wmask, float* thr, int* tie_cut,
                       float* votes_or_null, float* partial, float* pooled, cudaStream_t stream);
// window top-k (model_window_topk.py:118-203): window sums -> per-window top-k -> votes -> per-frame top-k (unfused reference kernels)
int window_sums(const float* acts, float* sums, int B, int T, int D, int window, int stride, int nw, cudaStream_t stream);
int window_votes(const float* acts, const float* sums, const float* thr_w, const int* cut_w, float* votes,
                 int B, int T, int D, int window, int stride, int nw, cudaStream_t stream);
// keep-by-votes: encoded = acts * mask(votes) ; pooled likewise
int votes_densify(const float* acts, const float* votes, const float* thr, const int* tie_cut, float* encoded, long long rows, int D, cudaStream_t stream);
int votes_mean_pool(const float* acts, const float* votes, const float* thr, const int* tie_cut, float* pooled, int B, int T, int D, const int* lens, cudaStream_t stream);
// compact (index, value) form of the kept entries, ascending feature index, k slots per row (unused: idx -1, val 0)
int votes_compact(const float* acts, const float* votes, const float* thr, const int* tie_cut, int* idx_out, float* val_out, int* count_out,
                  long long rows, int D, int k, cudaStream_t stream);
// classifier: LN(D) -> Linear(D,Hd) -> ReLU -> Linear(Hd,2) -> log_softmax   (model.py:183-189, :246-247)
int classifier_head(const float* pooled, int B, int D, int Hd, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* hidden /*[B, Hd] scratch*/, float* logprob, cudaStream_t stream);
// mean over valid frames of a [B, T, D] fp32 stream (use_sae=False / use_sparse_features=False pooling)
int mean_pool_frames(const float* x, float* pooled, int B, int T, int D, const int* lens, cudaStream_t stream);
// SLS (model_backup.py:186-202 + upstream classifier)
int sls_layer_weights(const float* const* layers, int n_layers, int B, int T, int D, const float* fc0_w, const float* fc0_b,
                      float* layer_w /*[B, n_layers]*/, const int* lens, cudaStream_t stream);
// same weights from dots[l][b*T + t] = fc0_w . x_l[b, t, :] (emitted by the LayerNorm kernels, LnArgs::dot_out)
int sls_layer_weights_from_dots(const float* dots, int n_layers, int B, int T, const float* fc0_b, float* layer_w, cudaStream_t stream);
int sls_fuse_pool(const void* const* layers, int layers_bf16, int n_layers, const float* layer_w, int B, int T, int D, const float* bn /*w,b,rm,rv*/,
                  float bn_eps, void* out /*[B, (T/3)*(D/3) zero-padded to ldo], fp32 or bf16*/, int out_bf16, int ldo, cudaStream_t stream);
// fc1 split-K partials [B][KS][Hd] -> selu(sum + b1) -> fc3 -> selu -> log_softmax
int sls_tail(const float* partial, int KS, int B, int Hd, const float* b1, const float* w3, const float* b3, float* logprob, cudaStream_t stream);
// scores = exp(logprob[:, 1])   (main.py:183-184)
int scores_from_logprob(const float* logprob, float* scores, int B, cudaStream_t stream);
// mse between recon and x (model.py:224-225), deterministic two-stage reduction
int mse_loss(const float* a, const float* b, long long n, float* out_scalar, float* scratch, cudaStream_t stream);

// ---------------------------------------------------------------- device FLAC decode (flac_gpu.cu)
}  // namespace slsb
#include "../../include/slsb200.h"
namespace slsb {
int flac_decode_frames(const uint8_t* bytes_dev, const slsb_flac_frame* frames_dev, int n_frames, int16_t* pcm_dev, int32_t* status_dev, cudaStream_t stream);

}  // namespace slsb

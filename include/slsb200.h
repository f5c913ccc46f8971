/* slsb200 -- C ABI of the B200-native XLS-R-300M + {TopK-SAE | window-TopK | SLS} scoring path.
 *
 * This is the drop-in boundary for the reference's `Model(args, device).forward(x) -> log-probs`
 * (SLSforASVspoof-2021-DF, citations relative to /root/reference):
 *
 *   slsb_forward        <->  model.py:195-260  Model.forward            (H-SAE, what main.py --is_eval runs)
 *                            model_window_topk.py:324-393               (H-WIN, --use_window_topk)
 *                            model_backup.py:167-202 + upstream SLS     (H-SLS)
 *   slsb_extract_feat   <->  model.py:128-141  SSLModel.extract_feat    (fairseq Wav2Vec2Model.forward(mask=False,
 *                            features_only=True)['x'], wav2vec/wav2vec2.py:540-647)
 *   slsb_get_tensor     <->  ...['layer_results'] (wav2vec2.py:958), sae activations (model.py:236-240)
 *   slsb_sae_encode     <->  model.py:68-79 / model_window_topk.py:68-116  AutoEncoderTopK.encode
 *   slsb_sae_decode     <->  model.py:81-83  AutoEncoderTopK.decode
 *   slsb_score_host     <->  main.py:172-193  produce_evaluation_file's per-batch body
 *                            (batch_x.to(device); model(...); exp(out)[:,1].cpu())
 *
 * Conventions: plain pointers and sizes only; every *_dev pointer is device memory owned by the CALLER;
 * the engine owns its weight arena and workspace; `stream` is a cudaStream_t passed as void*;
 * no call synchronises the host except slsb_score_host (which returns host scores) and slsb_create/destroy;
 * one engine per GPU per process, not thread-safe per engine.  Every function returns 0 on success and a
 * negative code on failure; slsb_last_error() gives the message (thread-local).  There is no CPU fallback:
 * without a CUDA device slsb_create fails.
 */
#ifndef SLSB200_H
#define SLSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLSB_ABI_VERSION 1

enum { SLSB_HEAD_NONE = 0, SLSB_HEAD_SAE = 1, SLSB_HEAD_WINDOW = 2, SLSB_HEAD_SLS = 3 };
/* OR-ed into `head`: keep this forward's intermediates readable afterwards through slsb_get_tensor("layer_results.<i>" / "acts" /
 * "encoded"), slsb_get_sparse and slsb_sae_loss.  Without it a forward keeps only what the score needs ("x", "pooled",
 * "sls_weights"): the bf16 layer-result snapshots are written only for the SLS head (which consumes them), and the SAE heads run
 * their fused select + pool path, which never materialises per-row thresholds or dense codes (model.py:236-240 and :224-225 are the
 * reference paths that need them: return_interpretability / return_sae_loss). */
#define SLSB_HEAD_RETAIN 0x100
enum { SLSB_PREC_FP32 = 0, SLSB_PREC_BF16 = 1 };
enum { SLSB_ATTN_AUTO = 0, SLSB_ATTN_SIMT = 1, SLSB_ATTN_TC = 2 /* persistent, P in TMEM */, SLSB_ATTN_TC_V1 = 3 /* one CTA per query tile, P in smem */ };

typedef struct slsb_config {
    int32_t n_conv;              /* 7 */
    int32_t conv_dim;            /* 512 */
    int32_t conv_kernel[8];      /* 10,3,3,3,3,2,2 */
    int32_t conv_stride[8];      /* 5,2,2,2,2,2,2 */
    int32_t embed_dim;           /* 1024 */
    int32_t ffn_dim;             /* 4096 */
    int32_t n_heads;             /* 16 (head dim must be 64) */
    int32_t n_layers;            /* 24 */
    int32_t pos_kernel;          /* 128 */
    int32_t pos_groups;          /* 16 (group width must be 64) */
    int32_t sae_dict;            /* 4096; 0 = no SAE head weights */
    int32_t sae_k;               /* 128 */
    int32_t sae_window;          /* 8 for H-WIN, 1 otherwise */
    int32_t cls_in;              /* classifier input dim: sae_dict (sparse features) or embed_dim */
    int32_t cls_hidden;          /* 256 */
    int32_t sls_frames;          /* 201: frame count the SLS fc1 was sized for; 0 = no SLS head weights */
    int32_t sls_hidden;          /* 1024 */
    int32_t attn_impl;           /* SLSB_ATTN_* */
    int32_t reserved[8];
} slsb_config;

typedef struct slsb_engine slsb_engine;

int slsb_abi_version(void);
const char* slsb_last_error(void);

/* life cycle */
int slsb_create(const slsb_config* cfg, int device, slsb_engine** out);
int slsb_destroy(slsb_engine* e);

/* weights: `name` is one of the packed tensor names listed in INTEGRATION.md ("conv1.w", "L3.qkv.w", ...);
 * `src` is fp32, host or device memory (cudaMemcpyDefault); `numel` must match the engine's expectation.
 * slsb_finalize_weights builds the bf16 copies the tensor-core path reads. */
int slsb_set_weight(slsb_engine* e, const char* name, const float* src, int64_t numel, void* stream);
int slsb_finalize_weights(slsb_engine* e, void* stream);
int64_t slsb_weight_numel(slsb_engine* e, const char* name);   /* -1 if unknown */

/* shape helper: frames produced by the conv stack for `samples` input samples (wav2vec2.py:523-538) */
int slsb_frames_for_samples(const slsb_engine* e, int samples);

/* The hot path.  wav_dev: fp32 [B, S]; sample_lens_dev: int32 [B] valid samples per clip, or NULL (all S);
 * logprob_dev: fp32 [B, 2] (ignored for SLSB_HEAD_NONE).  Asynchronous on `stream`. */
int slsb_forward(slsb_engine* e, const float* wav_dev, const int32_t* sample_lens_dev, int B, int S,
                 int head, int precision, float* logprob_dev, void* stream);

/* SSLModel.extract_feat: x_dev fp32 [B, T, embed_dim] (final LayerNorm applied, wav2vec2.py:905-906) */
int slsb_extract_feat(slsb_engine* e, const float* wav_dev, const int32_t* sample_lens_dev, int B, int S,
                      int precision, float* x_dev, void* stream);

/* Tensors left in the workspace by the last forward/extract_feat, copied into caller memory (fp32):
 *   "x" [B,T,D] | "layer_results.<i>" [B,T,D] raw residual stream after layer i | "features" [B,T,conv_dim]
 *   "acts" [B,T,dict] post-ReLU pre-top-k | "encoded" [B,T,dict] | "pooled" [B,cls_in] | "sls_weights" [B,n_layers] */
int slsb_get_tensor(slsb_engine* e, const char* name, float* dst_dev, int64_t numel, void* stream);

/* Sparse form of the last forward's SAE code (model.py:236-240 `last_sparse_features`, what the 13 analyze_*.py readers
 * consume) without materialising the dense [B,T,dict] tensor: idx_dev int32 [B*T, sae_k] feature indices in ascending
 * order (-1 = unused slot), val_dev fp32 [B*T, sae_k], count_dev int32 [B*T] (may be NULL). */
int slsb_get_sparse(slsb_engine* e, int32_t* idx_dev, float* val_dev, int32_t* count_dev, void* stream);

/* AutoEncoderTopK.encode / decode on caller activations: x_dev fp32 [rows, embed_dim] with rows = B*T.
 * window > 1 applies the window top-k over T frames per utterance (rows must be a multiple of T). */
int slsb_sae_encode(slsb_engine* e, const float* x_dev, int64_t rows, int T, int window, int precision,
                    float* encoded_dev, void* stream);
int slsb_sae_decode(slsb_engine* e, const float* encoded_dev, int64_t rows, int precision, float* recon_dev, void* stream);
/* mean((decode(encoded) - x)^2) of the last forward (model.py:224-225); loss_dev: fp32 [1] */
int slsb_sae_loss(slsb_engine* e, int precision, float* loss_dev, void* stream);

/* End-to-end scoring of one batch with HOST buffers (pinned or pageable): H2D copy of the clips, forward,
 * score = exp(logprob[:,1]) (main.py:183-184), D2H copy of B floats, stream synchronised before returning. */
int slsb_score_host(slsb_engine* e, const float* wav_host, const int32_t* sample_lens_host, int B, int S,
                    int head, int precision, float* scores_host, void* stream);

/* The same, pipelined: slsb_score_submit enqueues upload + forward + score download and returns a ticket (>= 0) without
 * waiting; the upload runs on a private copy stream into one of two staging slots, so the H2D copy of batch i+1 overlaps
 * the forward of batch i.  At most 4 submissions are in flight (the 5th submit waits for the 1st).  wav_host / scores_host
 * must stay valid until slsb_score_wait(e, ticket) returns (ticket < 0 waits for everything submitted so far).
 * slsb_score_host == submit + wait.  (DataLoader prefetch + batch_x.to(device, non_blocking=True), main.py:165-178.) */
int64_t slsb_score_submit(slsb_engine* e, const float* wav_host, const int32_t* sample_lens_host, int B, int S,
                          int head, int precision, float* scores_host, void* stream);
int slsb_score_wait(slsb_engine* e, int64_t ticket);

/* Audio ingest on the device (replaces the float conversion + pad() of Dataset_*_eval.__getitem__, data_utils_SSL.py:109-115
 * and :58-65): 16-bit PCM -> float32 x / 32768 (what librosa/soundfile deliver for 16-bit FLAC/WAV); a clip with >= S samples
 * keeps its first S, a shorter one is tile-repeated up to S.  pcm_dev: the clips back to back; offsets_dev int64 [B] first
 * sample of clip b; lens_dev int32 [B] (>= 1); wav_dev fp32 [B, S]. */
int slsb_ingest_pcm16(const int16_t* pcm_dev, const int64_t* offsets_dev, const int32_t* lens_dev, int B, int S,
                      float* wav_dev, void* stream);
/* slsb_score_host for 16-bit PCM host buffers: uploads 2 bytes per sample of the UN-padded clips, ingests, scores. */
int slsb_score_pcm16_host(slsb_engine* e, const int16_t* pcm_host, int64_t total_samples, const int64_t* offsets_host,
                          const int32_t* lens_host, int B, int S, int head, int precision, float* scores_host, void* stream);

/* Native FLAC decode (host code, no GPU): replaces librosa.load -> libsndfile in Dataset_*_eval.__getitem__
 * (data_utils_SSL.py:109-113) for the FLAC corpora the reference scores.  data: a whole .flac file in memory.  max_samples > 0
 * stops after that many samples per channel (pad() keeps only the first 64 600, data_utils_SSL.py:60-61).  Frame CRC-8 / CRC-16 are
 * always checked; verify_md5 != 0 also checks STREAMINFO's MD5 when the stream is decoded completely.
 *  slsb_flac_decode: interleaved int32 samples at the stream's own bit depth; info int32 [6] = {sample rate, channels, bits per
 *    sample, total samples (low 31 bits), MD5 verified (0/1), total samples >> 31}.  pcm_out NULL: STREAMINFO only, returns 0.
 *  slsb_flac_decode_mono16: 16-bit streams only, channels averaged to mono (halves away from zero), what the scorer ingests.
 * Both return the samples per channel decoded, or < 0: -2 not FLAC, -3 truncated, -4 bad header, -5 CRC-8, -6 CRC-16, -7 reserved
 * bit pattern, -8 MD5 mismatch, -9 unsupported (mid-stream format change, bit depth), -10 STREAMINFO, -11 bad argument. */
int64_t slsb_flac_decode(const uint8_t* data, int64_t nbytes, int64_t max_samples, int verify_md5, int32_t* pcm_out,
                         int64_t pcm_capacity, int32_t* info);
int64_t slsb_flac_decode_mono16(const uint8_t* data, int64_t nbytes, int64_t max_samples, int verify_md5, int16_t* pcm_out,
                                int64_t pcm_capacity, int32_t* sample_rate);

/* ---- device FLAC decode (csrc/flac_gpu.cu, csrc/flac_frame.h): replaces the per-item librosa decode of data_utils_SSL.py:109-113 for
 * the corpus format (16-bit mono FLAC).  The host scans a stream once (frame boundaries, header CRC-8, frame CRC-16, no sample is
 * decoded); the device decodes one frame per thread.
 *  slsb_flac_scan: frame table of the frames covering the first max_samples samples (0 = all).  info int32 [8] = {sample rate, channels,
 *    bits per sample, total samples (low 31 bits), max block size, total >> 31, samples covered by the table (capped at max_samples), 0}.
 *    frame_off / frame_len: byte offset of each frame's sync code in `data` and its length incl. the CRC-16; frame_samples: block sizes.
 *    Returns the number of frames (<= cap) or a negative FLAC error code (same codes as slsb_flac_decode).
 *  slsb_flac_frame: one decode job.  byte_off: frame start in the byte buffer (which must stay readable 8 bytes past its last frame);
 *    keep: samples of the frame to store (the clip head may end inside its last frame); out_off: first output sample in pcm; bps: STREAMINFO's.
 *  slsb_flac_decode_frames: device buffers; status[i] = block size of frame i (> 0) or a negative error (-9: not 16-bit mono - decode this
 *    clip with slsb_flac_decode_mono16 instead).  slsb_flac_decode_frames_host: the same frame core on host buffers (test twin). */
typedef struct slsb_flac_frame {
    int64_t byte_off;
    int64_t out_off;
    int32_t byte_len;
    int32_t keep;
    int32_t bps;
    int32_t reserved;
} slsb_flac_frame;
int64_t slsb_flac_scan(const uint8_t* data, int64_t nbytes, int64_t max_samples, int32_t* info, int64_t* frame_off, int32_t* frame_len,
                       int32_t* frame_samples, int64_t cap);
int slsb_flac_decode_frames(const uint8_t* bytes_dev, const slsb_flac_frame* frames_dev, int n_frames, int16_t* pcm_dev, int32_t* status_dev,
                            void* stream);
/* FLAC bytes -> scores: the compressed frames of B clips (frame table from slsb_flac_scan, out_off / offsets / lens in samples of one
 * int16 buffer of total_samples) are uploaded, decoded on the device, converted + padded like slsb_score_pcm16_host and scored.
 * status_host int32 [n_frames] receives the per-frame decoder status.  Returns 0, -1 (error) or -2 (a frame was refused by the device
 * decoder: the scores are void, decode this batch on the host instead). */
int slsb_score_flac_host(slsb_engine* e, const uint8_t* bytes_host, int64_t nbytes, const slsb_flac_frame* frames_host, int n_frames,
                         int64_t total_samples, const int64_t* offsets_host, const int32_t* lens_host, int B, int S, int head, int precision,
                         float* scores_host, int32_t* status_host, void* stream);
int slsb_flac_decode_frames_host(const uint8_t* bytes, const slsb_flac_frame* frames, int n_frames, int16_t* pcm, int32_t* status);

/* synthetic clips keyed by utterance index, bit-identical to oracle.trunk.synth_clips */
int slsb_synth_clips(float* wav_dev, int64_t first_utt, int count, int samples, void* stream);

/* how many kernels the engine launched since creation (bench.py's gpu_launches) */
int64_t slsb_launch_count(const slsb_engine* e);

/* Per-launch CUDA-event timing of the tensor-core kernels (events recorded on the launch stream).
 * kind: 0 qkv, 1 out_proj, 2 fc1, 3 fc2 (encoder GEMMs), 4 conv-stack implicit GEMMs, 5 positional conv, 6 other GEMMs, 7 attention
 * (these report algorithmic FLOPs); 8 LayerNorm(+residual) kernels, 9 SLS weighted-sum/BN/SELU/max-pool, 10 SLS fc1 (these
 * report algorithmic HBM BYTES in flops_out).  slsb_profile_read synchronises the device and sums elapsed ms / work / launches. */
int slsb_profile_enable(slsb_engine* e, int on);
int slsb_profile_read(slsb_engine* e, int kind, double* ms_out, double* flops_out, int64_t* launches_out);

/* ---- single-op entry points (unit tests call each kernel through the same ABI) ---- */
/* out = act(A[M,K] * W[N,K]^T + bias[N]) (+ residual[M,N]); act: 0 none, 1 gelu, 2 relu.
 * precision fp32: A, W, out fp32 (CUDA cores).  bf16: A, W bf16 (tcgen05), out bf16 if out_bf16 else fp32. */
int slsb_op_gemm(int precision, const void* A, const void* W, const float* bias, const float* residual, void* out,
                 int M, int N, int K, int act, int out_bf16, void* stream);
/* split-K tcgen05 GEMM (SLS fc1): A bf16 [M,K], W bf16 [N,K] -> fp32 partials [M][k_splits][N]; split s covers k-blocks
 * [s*ceil(K/64/k_splits), ...); k_splits must leave no empty split.  Synchronises the stream (test hook). */
int slsb_op_gemm_splitk(const void* A, const void* W, float* partial, int M, int N, int K, int k_splits, void* stream);
/* Conv1d(C->N, k, stride) over channels-last x[B, L_in, C] as implicit GEMM; W is [N, k*C] tap-major */
int slsb_op_conv(int precision, const void* x, const void* W, const float* bias, void* out, int B, int L_in, int C, int N,
                 int k, int stride, void* stream);
/* bf16 tensor-core feature-extractor layer: GELU(LayerNorm_512(Conv1d(x))) in one kernel; x bf16 [B, L_in, 512], out bf16 */
int slsb_op_conv_ln_gelu(const void* x, const void* W, const float* bias, const float* ln_w, const float* ln_b, void* out,
                         int B, int L_in, int C, int k, int stride, void* stream);
/* conv0 (raw audio, k = 10, stride 5) on tensor cores via a hi/lo bf16 split; scratch >= 64 KB + B*L0*128 bytes; out bf16 */
int slsb_op_conv0_tc(const float* wav, const float* w, const float* bias, const float* ln_w, const float* ln_b, void* out,
                     void* scratch, int B, int S, void* stream);
/* conv0 as ONE kernel (csrc/conv0_tc.cu; replaces the first block of wav2vec2.py:785-822): A tiles (hi/lo split of the audio window, bias
 * in spare K columns) built in shared memory, LayerNorm statistics from the 11 x 11 Gram matrix of [w | b], weights resident, 16 epilogue
 * warps; scratch >= 64 KB + 2 KB; out bf16 [B*L0, 512] */
int slsb_op_conv0_fused(const float* wav, const float* w, const float* bias, const float* ln_w, const float* ln_b, void* out,
                        void* scratch, int B, int S, void* stream);
/* grouped positional conv + GELU + residual: x fp32 [B,T,D]; W [D, K*64] (per out channel: tap-major, 64 in-channels) */
int slsb_op_posconv(int precision, const float* x, const void* W, const float* bias, float* out, void* scratch,
                    int B, int T, int D, int K, const int32_t* frame_lens_dev, void* stream);
int slsb_op_conv0(int out_bf16, const float* wav, const float* w, const float* bias, const float* ln_w, const float* ln_b,
                  void* out, int B, int S, int exact_gelu, void* stream);
int slsb_op_layernorm(const void* in, int in_bf16, void* out, int out_bf16, const float* w, const float* b,
                      int64_t rows, int C, int gelu, int exact_gelu, void* stream);
/* Host-side replay (no GPU needed) of the static tile schedule of the CTA-pair tcgen05 GEMM for an M x N output on num_pairs
 * CTA pairs: items int32 [max_items][5] = {pair, round, first row, first column, columns}; *split_out = column slices per tile
 * of the partial last round (1, 2 or 4).  Returns the number of items (may exceed max_items) or -1.  Test hook, no reference
 * counterpart: the schedule must produce every output element exactly once. */
int slsb_debug_pair_schedule(int M, int N, int num_pairs, int32_t* items, int max_items, int32_t* split_out);
/* the encoder's LayerNorm as the layer loop uses it (wav2vec2.py:1045 / :1055 on the fp32 stream): bf16 normalised rows, and
 * optionally a bf16 copy of the un-normalised rows (the layer_results snapshot, wav2vec2.py:958) and dot_out[row] =
 * <row, dot_w> (SLS fc0 before the mean over frames, model_backup.py:186-202 getAttenF). */
int slsb_op_layernorm_taps(const float* in, void* out_bf16, const float* w, const float* b, const float* dot_w, float* dot_out,
                           void* copy_out_bf16, int64_t rows, int C, void* stream);
int slsb_op_attention(int impl, int io_bf16, const void* qkv, void* out, int B, int T, int H,
                      const int32_t* frame_lens_dev, void* stream);
/* tcgen05 attention with a pipeline timeline: trace_dev int64 [64 units][16 events] of CTA 0, SM clock64 stamps
 * (0 S issue, 1 S committed, 2 P seen by MMA warp, 3 PV committed, 4 S seen by softmax, 5 max pass done, 6 P published,
 *  7 O seen, 8 epilogue done, 9 stage free seen by producer, 10 loads landed) -- tuning hook, no reference counterpart */
int slsb_op_attention_trace(const void* qkv, void* out, int B, int T, int H, int64_t* trace_dev, void* stream);
int slsb_op_topk(const float* x, int64_t rows, int D, int k, float* thr, int32_t* tie_cut, float* encoded_or_null, void* stream);
/* Fused scoring paths of H-SAE / H-WIN (heads.cu): exact per-frame top-k of acts [B*T, D] (D in 1024/2048/4096/8192) and the mean of the
 * kept activations over the frames < lens[b] (lens may be NULL), summed in the canonical order: chunks of 8 frames in frame order, then
 * the chunk partials in chunk order.  partial: [B, slsb_op_pool_chunks(T), D] scratch.  Mirrors model.py:68-79 + :245 (H-SAE) and
 * model_window_topk.py:118-203 + :365 (H-WIN; wmask: [B, nw, 256] words of scratch, votes_or_null [B*T, D] optional). */
int slsb_op_topk_pool(const float* acts, int B, int T, int D, int k, const int32_t* lens, float* thr, int32_t* tie_cut, float* partial, float* pooled,
                      void* stream);
int slsb_op_window_pool(const float* acts, int B, int T, int D, int k, int window, uint32_t* wmask, float* thr, int32_t* tie_cut, float* votes_or_null,
                        float* partial, float* pooled, void* stream);
int slsb_op_pool_chunks(int T);

#ifdef __cplusplus
}
#endif
#endif /* SLSB200_H */

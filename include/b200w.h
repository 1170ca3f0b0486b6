/* libb200whisper.so -- C ABI of the B200-native Whisper hot path.
 *
 * This is the drop-in boundary for the reference's ASR backend.  The reference (pure Python)
 * reaches its hot path through third-party Python/C++ packages:
 *   /root/reference/asr_components.py:279-288   self.model.transcribe(...)  -> faster-whisper
 *        -> FeatureExtractor.__call__ (NumPy log-mel)                       [b200w_logmel*]
 *        -> WhisperModel.encode -> ctranslate2 Whisper.encode               [b200w_encode]
 *        -> ctranslate2 Whisper.generate (cross K/V cached at step 0,       [b200w_cross_kv,
 *           greedy / beam search, logit rules, no_speech_prob)               b200w_decode]
 *        -> ctranslate2 Whisper.align (cross-attention DTW word timing)      [b200w_align]
 *        -> ctranslate2 Whisper.detect_language                              [b200w_detect_language]
 *   /root/reference/audio_deduplicator.py:60-181  _extract_features (NumPy) [b200w_dedup_features]
 * The Python host side (realtime-whisper-asr_b200/) binds these with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 on success and -1 on failure (message via
 * b200w_last_error(), thread-local).  "dev" pointers are CUDA device pointers on the current
 * device, "host" pointers are ordinary host memory.  `stream` is a cudaStream_t passed as void*
 * (NULL = default stream).  No CPU fallback exists: without a CUDA device every compute entry
 * point fails.
 */
#ifndef B200W_H_
#define B200W_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200W_ABI_VERSION 3

const char* b200w_last_error(void);
int b200w_abi_version(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
uint64_t b200w_launch_count(void);

/* Measurement hook: record CUDA events around every GEMM launch of one tile-width class
 * (bn = 64 | 128 | 256) until _end, which synchronises those events and returns the summed device
 * time (ms), the summed algorithmic flops (2*M*N*K per launch) and the launch count. */
/* Development aid: every CTA of later GEMM launches writes 16 int64 timestamps (clock64 / globaltimer)
 * to dev_buf; NULL switches it off. */
int b200w_gemm_set_trace(void* dev_buf);
/* Per-op timing of the decoder passes (bench.py's roofline leg for the decode kernels): while enabled,
 * decode steps are launched eagerly (no CUDA graph) with CUDA events around every op on the launching
 * stream; totals accumulate per op id until the next enable(1).  read(): op 0..15 -> name, total ms,
 * launches and algorithmic work (bytes for op 7 = decode cross attention; 0 where not tracked). */
struct b200w_engine;
int b200w_op_profile_enable(struct b200w_engine* e, int on);
int b200w_op_profile_read(struct b200w_engine* e, int op, const char** name, double* total_ms, int64_t* launches,
                          double* work);
/* Step probe (always on): CUDA events around the decode cross attention of the middle decoder layer in every decode
 * step, recorded INSIDE the captured step graph (external event nodes).  Totals since the last reset: elapsed ms,
 * launches, algorithmic bytes (K and V of every live (stream, head) once). */
int b200w_step_probe_reset(struct b200w_engine* e);
int b200w_step_probe_read(struct b200w_engine* e, double* total_ms, int64_t* launches, double* bytes);
int b200w_gemm_profile_begin(int bn);
int b200w_gemm_profile_end(double* total_ms, double* total_flops, int64_t* launches);
/* The same for the encoder self-attention kernel (algorithmic flops 4 T^2 d per window and launch). */
int b200w_attn_profile_begin(void);
int b200w_attn_profile_end(double* total_ms, double* total_flops, int64_t* launches);

/* ---------------------------------------------------------------- front end (HBM-bound) ------ */

/* Log-mel features of n_streams clips (faster-whisper FeatureExtractor.__call__, padding=160).
 *   audio_dev        fp32 concatenated 16 kHz samples
 *   offsets_dev      int64 [n_streams+1] sample offsets into audio_dev
 *   out_offsets_dev  int64 [n_streams]   element offsets into out_dev
 *   tile_prefix_dev  int32 [n_streams+1] cumulative ceil(frames_s / 32), frames_s = N_s/160 + 1
 *   out_dev          fp32; stream s gets a row-major [n_mels][frames_s] matrix
 *   stream_max_dev   int32 [n_streams] scratch (per-call running max)
 *   max_frames       max_s frames_s            n_mels 80 or 128 */
int b200w_logmel(const float* audio_dev, const int64_t* offsets_dev, const int64_t* out_offsets_dev,
                 const int32_t* tile_prefix_dev, int n_streams, int total_tiles, int n_mels,
                 float* out_dev, int32_t* stream_max_dev, int max_frames, void* stream);

/* Cut encoder windows out of log-mel features: window w = features of stream win_stream[w],
 * frames [seeks[w], seeks[w]+3000), zero-padded (pad_or_trim on the FEATURES), transposed to
 * time-major bf16 [n_windows][3000][128] (channels >= n_mels are zero). */
int b200w_pack_windows(const float* feats_dev, const int64_t* feat_offsets_dev,
                       const int32_t* n_frames_dev, const int32_t* seeks_dev,
                       const int32_t* win_stream_dev, int n_windows, int n_mels,
                       void* windows_bf16_dev, void* stream);

/* AudioDeduplicator._extract_features for n_chunks chunks of the same audio buffer:
 * [rms, centroid, zcr, rolloff85, bandwidth] / max|.|  -> out_dev fp32 [n_chunks][5];
 * raw_out_dev (optional, may be NULL) receives the un-normalised values.
 * max_points = max over chunks of (N even ? N/2 : N); must be <= 27648.  A chunk whose length has a
 * prime factor > 64 is reported as NaN features. */
int b200w_dedup_features(const float* audio_dev, const int64_t* offsets_dev, int n_chunks,
                         int max_points, int sample_rate, float* out_dev, float* raw_out_dev,
                         void* stream);

/* Band-pass pre-filter of n_chunks chunks: scipy.signal.filtfilt(b, a, chunk) with scipy's defaults (odd extension
 * by 27 samples, steady-state initial conditions) for an 8th-order IIR -- what the reference applies to every
 * incoming chunk (/root/reference/vocal_separation.py:335-365 SimpleFilterSeparator.separate: 4th-order Butterworth
 * band-pass 85-3400 Hz; call site 一键实时识别麦克风.py:1474-1482).
 *   coeffs_host  26 doubles (HOST memory): b[9], a[9] with a[0] == 1, zi[8] (scipy.signal.lfilter_zi)
 *   out_dev      fp32, same layout as audio_dev       scratch_dev  fp32 [n_chunks][27]
 * Chunks of <= 27 samples are copied unchanged (scipy refuses them; the reference then keeps the original audio). */
int b200w_bandpass_filtfilt(const float* audio_dev, const int64_t* offsets_dev, int n_chunks,
                            const double* coeffs_host, float* out_dev, float* scratch_dev, void* stream);

/* Batched VAD gate: the reference's DynamicVADIterator (/root/reference/asr_components.py:12-78 = upstream
 * FixedVADIterator re-chunking + VADIterator state machine with a live-tunable min_silence_samples) for many streams.
 *   b200w_vad_window_energy  built-in scorer: probs[w] = rms / (rms + k) of window w (fp32 [n_windows][512]); any
 *                            other scorer (a batched Silero network) fills the same probs buffer instead
 *   b200w_vad_gate           one thread per stream: windows [win_offsets[s], win_offsets[s+1]) in order through the
 *                            state machine; state int64 [n_streams][3] = (triggered, temp_end, current_sample) is
 *                            updated in place; min_silence_dev double [n_streams] (samples); events int64
 *                            [n_streams][4] = (has_start, start, has_end, end) merged over the call's windows as
 *                            FixedVADIterator.__call__ merges them */
int b200w_vad_window_energy(const float* windows_dev, int n_windows, float k, float* probs_dev, void* stream);
int b200w_vad_gate(const float* probs_dev, const int32_t* win_offsets_dev, int n_streams, int64_t* state_dev,
                   const double* min_silence_dev, float threshold, int64_t speech_pad_samples, int64_t* events_dev,
                   void* stream);

/* ---------------------------------------------------------------- building blocks (tests) ---- */

/* out[M][N] = act(A[M][K] * W[N][K]^T + bias) + residual.  bf16 operands, fp32 accumulate on
 * tcgen05.  flags: bit0 out fp32, bit1 residual fp32, bit2 GELU, bits 8..15 split-K count s (> 1: out
 * receives fp32 partial sums [s][M][ldo], no bias / activation / residual).  force_bn: 0 | 64 | 128 | 256. */
int b200w_gemm_bf16(const void* A_dev, const void* W_dev, const float* bias_dev,
                    const void* residual_dev, void* out_dev, int M, int N, int K, int64_t lda,
                    int64_t ldw, int64_t ldr, int64_t ldo, int flags, int force_bn, void* stream);

/* Decode-step (skinny) GEMM, M <= 256: out[M][N] = act(X[M][K] * W[N][K]^T + bias), weight rows on the
 * MMA's M dimension (swap-AB), see csrc/skinny.cu.  flags: bit0 out fp32, bit2 GELU.  splits > 1: out
 * receives fp32 partial sums [splits][M][ldo] (bias must be NULL, no GELU). */
int b200w_skinny_gemm(const void* X_dev, const void* W_dev, const float* bias_dev, void* out_dev, int M, int N,
                      int K, int64_t ldx, int64_t ldw, int64_t ldo, int flags, int splits, void* stream);

/* Decode-step cross attention on tcgen05 (csrc/attention.cu, cross_attn_sp_kernel): groups of <= 8 consecutive
 * query rows of one stream share one pass over that stream's keys / values.
 *   cross_kv_dev  bf16 [n_streams][n_layers][k|v][n_heads][len][64]       q_dev  bf16 [rows][n_heads*64]
 *   stream_dev    int32 [rows] stream of each row       g_start_dev / g_count_dev  int32 [n_groups]
 *   n_live_dev    int32 [1]: groups [0, n_live) are processed             out_dev bf16 [rows][n_heads*64]
 *   scratch_dev   b200w_cross_attention_tc_scratch_bytes() bytes, zeroed once by the caller
 *   pieces        key ranges every (group, head) item is cut into (a divisor of ceil(len/128), else 1 is used):
 *                 a function of the shapes only, so a stream's result never depends on its batch neighbours */
int64_t b200w_cross_attention_tc_scratch_bytes(void);
int b200w_cross_attention_tc(const void* cross_kv_dev, int n_streams, int n_layers, int layer, int n_heads, int len,
                             const void* q_dev, const int32_t* stream_dev, const int32_t* g_start_dev,
                             const int32_t* g_count_dev, const int32_t* n_live_dev, int n_groups, void* out_dev,
                             float scale, void* scratch_dev, int pieces, void* stream);

/* Cross attention for passes with many query rows per stream (prompts, alignment): flash kernel, one CTA per
 * (128 query rows of a segment, head).  Segments = runs of consecutive query rows of one stream.
 *   q_dev bf16 [n_q_rows][n_heads*64]   row_stream_dev int32 [n_q_rows]   seg_start_dev / seg_count_dev int32 [n_segs] */
int b200w_cross_attention_flash(const void* cross_kv_dev, int n_streams, int n_layers, int layer, int n_heads, int len,
                                const void* q_dev, int n_q_rows, const int32_t* row_stream_dev,
                                const int32_t* seg_start_dev, const int32_t* seg_count_dev, int n_segs,
                                int max_seg_rows, void* out_dev, void* stream);

/* Decode-step self-attention (csrc/attention.cu, qkv_attn_kernel): split-K partial sums of the fused QKV
 * projection -> bias, bf16 -> the row's new key / value into the paged cache -> attention over the row's history
 * (through its page-table row) plus its own position; one warp per (row, head), n_heads % 4 == 0.  rows_per_group:
 * how many consecutive rows belong to one stream (a hint, unused by the shipped kernel).
 *   partial_dev fp32 [n_splits][rows][3d]   bias_dev fp32 [3d]   slot_dev int32 [rows] slot of the row's own token
 *   kpool_dev / vpool_dev bf16 [slots][d]   table_dev int32 [table rows][table_stride] position -> slot
 *   trow_dev int32 [rows] page-table row    kvlen_dev int32 [rows] visible keys INCLUDING the row's own position
 *   out_dev bf16 [rows][d]                  max_len >= every kvlen */
int b200w_step_self_attention(const float* partial_dev, int n_splits, const float* bias_dev, int rows, int d,
                              int n_heads, const int32_t* slot_dev, void* kpool_dev, void* vpool_dev,
                              const int32_t* table_dev, int table_stride, const int32_t* trow_dev,
                              const int32_t* kvlen_dev, int max_len, int rows_per_group, void* out_dev, void* stream);

/* Fused logit rules + log-softmax + top-k of the decode step (csrc/kernels.cu, sample_row_kernel): per row, the
 * static suppress mask (bit0 always, bit1 at the first sampled position), Whisper's timestamp rules driven by
 * state[row] = (flags, lowest admissible timestamp id, 0, 0) with flags bit0 rules on, bit1 first position, bit2
 * no timestamp may follow, bit3 only timestamps / EOT may follow; log-probabilities renormalised over the
 * timestamps when their total mass beats every text token; the k best (ties: lower id) -> ids / logprobs [rows][k].
 *   logits_dev fp32 [rows][ld] (ld % 8 == 0)   mask_dev uint8 [n_vocab]   state_dev int32 [rows][4]
 *   streaming != 0 selects the kernel that streams the row from global memory three times (fallback for rows that
 *   do not fit in shared memory); both give bit-identical results.
 *   temperature > 0 (k must be 1): ONE token per row drawn from softmax(rule-masked logits / temperature) and its
 *   log-probability under that distribution; the draw is a pure function of (seed, state[row][3] = sample stream,
 *   state[row][2] = draw index). */
int b200w_sample_topk(const float* logits_dev, int rows, int ld, int n_vocab, const uint8_t* mask_dev,
                      const int32_t* state_dev, int eot, int no_timestamps, int timestamp_begin,
                      int max_initial_ts, int k, int32_t* out_ids_dev, float* out_logprobs_dev, int streaming,
                      float temperature, uint32_t seed, void* stream);

/* conv1d (kernel 3, padding 1, stride 1 or 2) as implicit GEMM over time-major activations:
 * x bf16 [B][T][C] -> GELU(conv(x) + bias) (+ addend fp32 [T_out][N], e.g. positions) ->
 * out [B][T_out][N] bf16 or fp32.  w bf16 [N][3*C] tap-major.  C % 64 == 0. */
int b200w_conv1d_k3(const void* x_dev, const void* w_dev, const float* bias_dev,
                    const float* addend_dev, void* out_dev, int B, int T, int C, int N, int stride,
                    int out_fp32, void* stream);

/* Encoder self-attention: qkv bf16 [B][T][3d] -> out bf16 [B][T][d], heads of 64. */
int b200w_enc_attention(const void* qkv_dev, int B, int T, int d, int n_heads, void* out_dev,
                        void* stream);

/* Profiling aid: with B200W_ENC_ATTN_DBG=2 in the environment one CTA of the encoder attention kernel records
 * clock64 stamps of its pipeline phases, [role 0..3][key block 0..31][stamp 0..7] (roles: softmax tile 0 / 1, MMA
 * issuer, TMA producer).  Copies the first n values to host_out. */
int b200w_enc_attention_trace_read(int64_t* host_out, int n);

/* ---------------------------------------------------------------- engine --------------------- */

typedef struct b200w_engine b200w_engine;

typedef struct b200w_model_config {
  int32_t n_mels;        /* 80 | 128 */
  int32_t d_model;       /* 384 .. 1280, multiple of 64 */
  int32_t n_heads;       /* d_model / 64 */
  int32_t n_enc_layers;
  int32_t n_dec_layers;
  int32_t n_vocab;       /* 51865 | 51866 */
  int32_t n_audio_ctx;   /* 1500 */
  int32_t n_text_ctx;    /* 448 */
  int32_t max_batch;     /* windows per encode / streams per decode */
  int32_t max_beam;      /* >= 1 */
} b200w_model_config;

int b200w_engine_create(const b200w_model_config* cfg, b200w_engine** out);
int b200w_engine_destroy(b200w_engine* e);

/* Register one weight tensor (device memory owned by the caller, kept alive for the engine's
 * lifetime).  dtype: 0 = fp32, 1 = bf16.  Names: see realtime-whisper-asr_b200/weights.py. */
int b200w_engine_set_weight(b200w_engine* e, const char* name, const void* dev_ptr, int64_t numel,
                            int dtype);
/* All weights registered: validate and allocate workspaces. */
int b200w_engine_finalize(b200w_engine* e);

/* Whisper encoder on B windows: bf16 [B][3000][128] -> bf16 [B][1500][d]. */
int b200w_encode(b200w_engine* e, const void* windows_dev, int B, void* enc_out_dev, void* stream);

/* Cross-attention K/V of every decoder layer, once per chunk (one GEMM over all layers):
 *  enc_out bf16 [B][1500][d] -> cross_kv bf16 [B][n_dec_layers][2][heads][1500][64] (head-major). */
int b200w_cross_kv(b200w_engine* e, const void* enc_out_dev, int B, void* cross_kv_dev, void* stream);

typedef struct b200w_decode_opts {
  int32_t beam_size;             /* 1 = greedy */
  float patience;                /* 1.0 */
  float length_penalty;          /* 1.0 */
  int32_t max_new_tokens;        /* cap on sampled tokens (<= n_text_ctx - prompt) ; 0 = no cap */
  int32_t suppress_blank;        /* 1 */
  int32_t without_timestamps;    /* 0 */
  int32_t max_initial_timestamp_index; /* 50 */
  int32_t eot, sot, no_speech, no_timestamps, timestamp_begin; /* token ids */
  const int32_t* suppress_ids;   /* host, static suppress list */
  int32_t n_suppress;
  const int32_t* blank_ids;      /* host, tokens suppressed at the first sampled position */
  int32_t n_blank;
  /* temperature > 0: sampling instead of search (faster-whisper's fallback temperatures): best_of independent
   * samples per chunk drawn from softmax(rule-masked logits / temperature), the one with the best length-normalised
   * log-probability is returned; beam_size is ignored.  Draws are a pure function of (seed, chunk row, position). */
  float temperature;             /* 0 = greedy / beam search */
  int32_t best_of;               /* samples per chunk (<= max_beam) */
  uint32_t seed;
} b200w_decode_opts;

/* Autoregressive decode of B chunks (greedy, beam search or temperature sampling, with Whisper's logit rules).
 *   prompts_host      int32 concatenated prompt tokens, prompt_lens_host int32 [B]
 *   tokens_out_host   int32 [B][max_out]   sampled tokens (without the prompt, without EOT)
 *   n_tokens_out_host int32 [B]
 *   sum_logprob_out_host fp32 [B]  cumulative log-prob of the returned hypothesis (incl. EOT)
 *   no_speech_out_host   fp32 [B]  softmax prob of <|nospeech|> at the <|startoftranscript|> slot */
int b200w_decode(b200w_engine* e, const void* cross_kv_dev, int B, const int32_t* prompts_host,
                 const int32_t* prompt_lens_host, const b200w_decode_opts* opts,
                 int32_t* tokens_out_host, int32_t max_out, int32_t* n_tokens_out_host,
                 float* sum_logprob_out_host, float* no_speech_out_host, void* stream);

/* Teacher-forced pass for word timing (CTranslate2 Whisper.align): for each of B sequences
 * (tokens = sot_sequence + [no_timestamps] + text + [eot]) returns the DTW path between text
 * tokens and encoder frames plus per-token probabilities.
 *   align_heads_host  int32 [n_align][2] (layer, head)
 *   num_frames_host   int32 [B]   content frames of the chunk (mel frames; //2 = encoder frames)
 *   sot_len           tokens before the first text token (sot sequence + no_timestamps)
 *   path_out_host     int32 [B][max_path][2] (text index, time index), path_len_out_host [B]
 *   tok_prob_out_host fp32 [B][max_tok]  probability of each text token (teacher forced) */
int b200w_align(b200w_engine* e, const void* cross_kv_dev, int B, const int32_t* tokens_host,
                const int32_t* token_lens_host, const int32_t* num_frames_host, int32_t sot_len,
                const int32_t* align_heads_host, int32_t n_align, int32_t median_width,
                int32_t* path_out_host, int32_t max_path, int32_t* path_len_out_host,
                float* tok_prob_out_host, int32_t max_tok, void* stream);

/* Language detection (faster-whisper detect_language): one decoder step on <|sot|>, softmax over
 * the language tokens.  probs_out_host fp32 [B][n_lang]. */
int b200w_detect_language(b200w_engine* e, const void* cross_kv_dev, int B, int32_t sot,
                          const int32_t* lang_ids_host, int32_t n_lang, float* probs_out_host,
                          void* stream);

/* Debug / parity hooks: full-vocabulary logits of one teacher-forced pass.
 * tokens_host concatenated, lens [B]; logits_out_dev fp32 [sum(lens)][n_vocab_padded]
 * (n_vocab_padded = n_vocab rounded up to 8). */
int b200w_decoder_logits(b200w_engine* e, const void* cross_kv_dev, int B, const int32_t* tokens_host,
                         const int32_t* token_lens_host, float* logits_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200W_H_ */

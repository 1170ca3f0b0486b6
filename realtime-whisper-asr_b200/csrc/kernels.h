// Launchers for the small memory-bound kernels (kernels.cu).  All pointers are device pointers.
#pragma once
#include "common.h"

namespace b200w {

// Pending split-K partial sums of the GEMM that feeds the residual stream: [n][rows][d] fp32 slabs
// `stride` elements apart, plus that GEMM's bias.  LayerNorm folds them into x (in slab order).
struct SplitReduce {
  const float* partial = nullptr;
  int n = 0;
  long long stride = 0;
  const float* bias = nullptr;
};
void layernorm(float* x, const int* row_idx, const float* gamma, const float* beta, bf16* y, int rows, int d,
               cudaStream_t stream, const SplitReduce* sr = nullptr);
void embed_tokens(const int* tok, const int* pos, const bf16* tok_emb, const float* pos_emb, float* x,
                  int n, int d, cudaStream_t stream);
void kv_scatter(const bf16* qkv, const int* slot, bf16* kpool, bf16* vpool, int n, int d,
                cudaStream_t stream);

// Reduce split-K slabs of a GEMM whose consumer is not a LayerNorm: out = bf16(act(bias + sum of slabs)).
// slot != null: fused-QKV rows, K and V columns are scattered into the KV-cache pools instead.
void split_finalize(const float* partial, int n_splits, long long stride, const float* bias, int rows, int N, int act,
                    bf16* out, long long ldo, const int* slot, bf16* kpool, bf16* vpool, int d, cudaStream_t stream);

// per-row sampling state flags (state[row].x); state[row].y = lowest admissible timestamp id
enum : int {
  SF_RULES = 1,       // apply Whisper timestamp rules
  SF_FIRST = 2,       // first sampled position
  SF_SUP_ALL_TS = 4,  // last two tokens were timestamps: no timestamp may follow
  SF_SUP_TEXT = 8,    // last token was an unpaired timestamp: only timestamps / EOT may follow
};
static constexpr int MAX_TOPK = 16;

struct SampleParams {
  const float* logits;  // [rows][ld]
  int ld, n_vocab;
  const unsigned char* mask;  // [n_vocab] bit0: always suppressed, bit1: suppressed at first position
  const int4* state;          // [rows]
  int eot, no_timestamps, timestamp_begin, max_initial_ts;
  int k;
  int* out_ids;         // [rows][k]
  float* out_logprobs;  // [rows][k]
  // temperature > 0: instead of the k best, ONE token per row drawn from softmax(masked logits / temperature)
  // (k must be 1); the draw of a row is a pure function of (seed, state[row].z = draw index, state[row].w = sample
  // stream), the log-probability returned is the token's under the tempered distribution
  float temperature = 0.f;
  unsigned seed = 0;
};
void sample_topk(const SampleParams& p, int rows, cudaStream_t stream, bool streaming = false);

void token_probs(const float* logits, int ld, int n_vocab, const int* rows, const int* tok,
                 const int* ids, int n_ids, float* out, int n_rows, cudaStream_t stream);

void align_postprocess(float* weights, int n_align, int n_q, int stride, const int* seq_q0,
                       const int* seq_ntok, const int* seq_nf, int n_seq, int max_ntok, int max_nf,
                       int median_width, float* matrix, int row_lo, int row_hi_trim,
                       signed char* trace, long long trace_stride, int* path, int max_path,
                       int* path_len, cudaStream_t stream);

void gather_fork_tables(const int* old_t, const int* parent, const int* row, const int* pos,
                        const int* slot, int* new_t, int n, int stride, cudaStream_t stream);

}  // namespace b200w

// Attention launchers (attention.cu).  All pointers are device pointers.
#pragma once
#include "common.h"

namespace b200w {

// Encoder self-attention over the fused QKV activation: qkv bf16 [B][T][3d] (q | k | v, heads of
// 64 inside each third) -> out bf16 [B][T][d].  softmax(q k^T / 8) v, non-causal.
void enc_attention(const bf16* qkv, int B, int T, int d, int n_heads, bf16* out, cudaStream_t stream);

// Per-launch CUDA-event timing of the encoder attention (algorithmic flops 4 T^2 d per window); see bench.py.
void enc_attention_profile_begin();
void enc_attention_profile_end(double* total_ms, double* total_flops, long long* launches);

// Decoder attention for a handful of queries; one warp per (query, head).
struct DecAttnParams {
  const bf16* q = nullptr;   // query i, head h at q + i * q_stride + h * 64
  int q_stride = 0;
  const bf16* kbase = nullptr;  // key of slot s, head h at kbase + base_off + s * kv_row_stride + h * 64
  const bf16* vbase = nullptr;
  long long kv_row_stride = 0;
  const int* table = nullptr;   // [rows][table_stride] position -> slot; null = identity
  int table_stride = 0;
  const int* trow = nullptr;    // [n_queries] table row (self) / stream index (cross)
  const long long* kv_base_off = nullptr;  // [rows] element offset per table row / stream, or null
  const int* kvlen = nullptr;   // [n_queries] number of visible keys
  bf16* out = nullptr;          // [n_queries][out_stride]
  int out_stride = 0;
  float scale = 0.125f;
  int n_queries = 0;
  int max_len = 0;              // upper bound on kvlen (sizes the score scratch)
  float* probs_out = nullptr;   // optional [n_export][n_queries][probs_stride]
  const int* head_export = nullptr;  // [n_heads] -> export index or -1
  int probs_stride = 0;
};
void dec_attention(const DecAttnParams& p, int n_heads, cudaStream_t stream);
// phase trace of one CTA of the encoder attention kernel (B200W_ENC_ATTN_DBG=2): [role][block][stamp] clock64 values
void enc_attention_trace_read(long long* host_out, int n);

// Decode steps: split-K QKV partial sums -> (bias, bf16) -> new K/V into the cache + self-attention of the row.
struct QkvAttnParams {
  const float* partial = nullptr;  // [n_splits][rows][3d] fp32 slabs
  int n_splits = 0;
  long long split_stride = 0;
  const float* bias = nullptr;     // [3d]
  int d = 0;
  const int* slot = nullptr;       // [rows] cache slot of the row's own token
  bf16* kpool = nullptr;           // this layer's pools [slots][d]
  bf16* vpool = nullptr;
  const int* table = nullptr;      // [table rows][table_stride] position -> slot
  int table_stride = 0;
  const int* trow = nullptr;       // [rows]
  const int* kvlen = nullptr;      // [rows] visible keys including the row's own position
  bf16* out = nullptr;
  int out_stride = 0;
  float scale = 0.125f;
  int max_len = 0;
};
void qkv_attention(const QkvAttnParams& p, int n_rows, int n_heads, cudaStream_t stream);

// Cross attention for the decoder: groups of up to 8 consecutive query rows of the SAME stream
// (beams of one chunk, or consecutive prompt / teacher-forced tokens) share one pass over that
// stream's K/V.  K/V of position t, head h: kbase/vbase + kv_base_off[stream] + h * head_stride + t * kv_row_stride.
struct CrossAttnParams {
  const bf16* q = nullptr;
  int q_stride = 0;
  const bf16* kbase = nullptr;
  const bf16* vbase = nullptr;
  long long kv_row_stride = 0;             // elements between consecutive positions
  long long head_stride = 64;              // elements between consecutive heads
  const long long* kv_base_off = nullptr;  // [streams]
  const int* stream = nullptr;             // [n_queries] stream of each query row
  const int* g_start = nullptr;            // [n_groups] first query row of the group
  const int* g_count = nullptr;            // [n_groups] rows in the group (1..8)
  int len = 0;                             // keys per stream (1500)
  bf16* out = nullptr;
  int out_stride = 0;
  float scale = 0.125f;
  int n_queries = 0;
  float* probs_out = nullptr;              // optional [n_export][n_queries][probs_stride]
  const int* head_export = nullptr;        // [n_heads] -> export index or -1
  int probs_stride = 0;
  int export_only = 0;  // 1: only heads with head_export[h] >= 0 are computed (the others were done elsewhere)
};
// max_group: upper bound on g_count (selects the specialisation of the streaming kernel)
// split_scratch (optional, cross_attention_scratch_bytes(64, n_heads) bytes): enables the split-KV path
// for passes with at most 64 groups (decode steps).
void cross_attention(const CrossAttnParams& p, int n_groups, int n_heads, int max_group, cudaStream_t stream,
                     float* split_scratch = nullptr);
size_t cross_attention_scratch_bytes(int n_groups, int n_heads);
static constexpr int CROSS_GROUP = 8;

// Cross attention for passes with many rows per stream (prompts, alignment): flash kernel (tcgen05), one CTA per
// (128 query rows of a segment, head).  Segments = runs of consecutive query rows of one stream.
void cross_attention_flash(const bf16* cross_kv, int n_streams, int n_layers, int layer, int n_heads, int len,
                           const bf16* q, int n_q_rows, const int* row_stream, const int* seg_start,
                           const int* seg_count, int n_segs, int max_seg_rows, bf16* out, cudaStream_t stream);

// Decode-step cross attention on tcgen05 (no probability export).  cross_kv is the whole per-chunk
// buffer [n_streams][n_layers][k|v][n_heads][len][64]; groups [0, *n_live_groups) are processed, each
// with g_count <= 8 live rows starting at g_start.
// pieces: every (group, head) item is cut into this many key ranges (must divide the tile count ceil(len/128),
// else 1), computed by different CTAs and merged in piece order: a function of the shapes only, never of how many
// items share the launch, so a stream's result does not depend on what is batched next to it.
// scratch: cross_attention_tc_scratch_bytes() bytes, zeroed once (partial results of the pieces and their
// self-resetting arrival counters).
size_t cross_attention_tc_scratch_bytes();
// byte offset (inside the scratch) of the int the kernel sets when a score leaves its exponent range
size_t cross_attention_tc_error_offset();
void cross_attention_tc(const bf16* cross_kv, int n_streams, int n_layers, int layer, int n_heads, int len,
                        const bf16* q, int q_stride, const int* stream, const int* g_start, const int* g_count,
                        const int* n_live_groups, int n_groups, bf16* out, int out_stride, float scale, void* scratch,
                        cudaStream_t cs, int pieces = 1);

}  // namespace b200w

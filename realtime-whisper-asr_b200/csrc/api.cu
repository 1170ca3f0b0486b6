// extern "C" surface of libb200whisper.so: front end + building blocks.  (Engine entry points
// live in engine.cu.)  Declarations: include/b200w.h.
#include "../../include/b200w.h"

#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "attention.h"
#include "common.h"
#include "dstep.h"
#include "frontend.h"
#include "gemm.h"
#include "kernels.h"
#include "skinny.h"

namespace b200w {

std::atomic<unsigned long long> g_launch_count{0};

static thread_local std::string t_last_error;
void set_last_error(const std::string& s) { t_last_error = s; }
const char* get_last_error() { return t_last_error.c_str(); }

bool pdl_enabled() {
  static const bool on = getenv("B200W_NO_PDL") == nullptr;
  return on;
}

int num_sms() {
  static std::atomic<int> cache[64];  // per device ordinal; 0 = not read yet
  int dev = 0;
  B200W_CHECK_CUDA(cudaGetDevice(&dev));
  const int slot = dev & 63;
  int n = cache[slot].load(std::memory_order_relaxed);
  if (n == 0) {
    B200W_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    cache[slot].store(n, std::memory_order_relaxed);
  }
  return n;
}

void* alloc_descriptor_record(size_t bytes) {
  static std::mutex mu;
  static char* chunk[64] = {nullptr};
  static size_t used[64] = {0};
  constexpr size_t CHUNK = 1 << 20;
  bytes = (bytes + 127) & ~static_cast<size_t>(127);
  B200W_REQUIRE(bytes <= CHUNK, "descriptor record too large");
  int dev = 0;
  B200W_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  const int s = dev & 63;
  if (!chunk[s] || used[s] + bytes > CHUNK) {
    void* p = nullptr;
    B200W_CHECK_CUDA(cudaMalloc(&p, CHUNK));  // (the previous chunk stays allocated: its records are in use)
    chunk[s] = static_cast<char*>(p);
    used[s] = 0;
  }
  void* r = chunk[s] + used[s];
  used[s] += bytes;
  return r;
}

void ensure_dynamic_smem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> done;
  int dev = 0;
  B200W_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = done[{dev, func}];
  if (bytes > have) {
    B200W_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
}

}  // namespace b200w

using namespace b200w;

extern "C" {

const char* b200w_last_error(void) { return get_last_error(); }
int b200w_abi_version(void) { return B200W_ABI_VERSION; }
uint64_t b200w_launch_count(void) { return g_launch_count.load(); }

int b200w_gemm_set_trace(void* dev_buf) {
  if (getenv("B200W_TRACE_DSTEP")) dstep_set_trace(dev_buf, atoi(getenv("B200W_TRACE_DSTEP")));
  else if (getenv("B200W_TRACE_SKINNY")) skinny_set_trace(dev_buf);
  else gemm_set_trace(dev_buf);
  return 0;
}

int b200w_gemm_profile_begin(int bn) {
  B200W_API_BEGIN
  gemm_profile_begin(bn);
  B200W_API_END
}
int b200w_gemm_profile_end(double* total_ms, double* total_flops, int64_t* launches) {
  B200W_API_BEGIN
  long long n = 0;
  gemm_profile_end(total_ms, total_flops, &n);
  *launches = n;
  B200W_API_END
}

int b200w_attn_profile_begin(void) {
  B200W_API_BEGIN
  enc_attention_profile_begin();
  B200W_API_END
}
int b200w_attn_profile_end(double* total_ms, double* total_flops, int64_t* launches) {
  B200W_API_BEGIN
  long long n = 0;
  enc_attention_profile_end(total_ms, total_flops, &n);
  *launches = n;
  B200W_API_END
}

int b200w_logmel(const float* audio_dev, const int64_t* offsets_dev, const int64_t* out_offsets_dev,
                 const int32_t* tile_prefix_dev, int n_streams, int total_tiles, int n_mels,
                 float* out_dev, int32_t* stream_max_dev, int max_frames, void* stream) {
  B200W_API_BEGIN
  logmel_launch(audio_dev, reinterpret_cast<const long long*>(offsets_dev),
                reinterpret_cast<const long long*>(out_offsets_dev), tile_prefix_dev, n_streams,
                total_tiles, n_mels, out_dev, stream_max_dev, max_frames,
                static_cast<cudaStream_t>(stream));
  B200W_API_END
}

int b200w_pack_windows(const float* feats_dev, const int64_t* feat_offsets_dev,
                       const int32_t* n_frames_dev, const int32_t* seeks_dev,
                       const int32_t* win_stream_dev, int n_windows, int n_mels,
                       void* windows_bf16_dev, void* stream) {
  B200W_API_BEGIN
  pack_windows_launch(feats_dev, reinterpret_cast<const long long*>(feat_offsets_dev), n_frames_dev,
                      seeks_dev, win_stream_dev, n_windows, n_mels,
                      static_cast<bf16*>(windows_bf16_dev), static_cast<cudaStream_t>(stream));
  B200W_API_END
}

int b200w_dedup_features(const float* audio_dev, const int64_t* offsets_dev, int n_chunks,
                         int max_points, int sample_rate, float* out_dev, float* raw_out_dev,
                         void* stream) {
  B200W_API_BEGIN
  dedup_features_launch(audio_dev, reinterpret_cast<const long long*>(offsets_dev), n_chunks,
                        max_points, sample_rate, out_dev, raw_out_dev,
                        static_cast<cudaStream_t>(stream));
  B200W_API_END
}

int b200w_bandpass_filtfilt(const float* audio_dev, const int64_t* offsets_dev, int n_chunks,
                            const double* coeffs_host, float* out_dev, float* scratch_dev, void* stream) {
  B200W_API_BEGIN
  B200W_REQUIRE(audio_dev && offsets_dev && coeffs_host && out_dev && scratch_dev, "null argument");
  bandpass_filtfilt_launch(audio_dev, reinterpret_cast<const long long*>(offsets_dev), n_chunks, coeffs_host, out_dev,
                           scratch_dev, static_cast<cudaStream_t>(stream));
  B200W_API_END
}

int b200w_vad_window_energy(const float* windows_dev, int n_windows, float k, float* probs_dev, void* stream) {
  B200W_API_BEGIN
  vad_window_energy_launch(windows_dev, n_windows, k, probs_dev, static_cast<cudaStream_t>(stream));
  B200W_API_END
}
int b200w_vad_gate(const float* probs_dev, const int32_t* win_offsets_dev, int n_streams, int64_t* state_dev,
                   const double* min_silence_dev, float threshold, int64_t speech_pad_samples, int64_t* events_dev,
                   void* stream) {
  B200W_API_BEGIN
  B200W_REQUIRE(probs_dev && win_offsets_dev && state_dev && min_silence_dev && events_dev, "null argument");
  vad_gate_launch(probs_dev, win_offsets_dev, n_streams, reinterpret_cast<long long*>(state_dev), min_silence_dev,
                  threshold, (long long)speech_pad_samples, reinterpret_cast<long long*>(events_dev),
                  static_cast<cudaStream_t>(stream));
  B200W_API_END
}

int b200w_gemm_bf16(const void* A_dev, const void* W_dev, const float* bias_dev,
                    const void* residual_dev, void* out_dev, int M, int N, int K, int64_t lda,
                    int64_t ldw, int64_t ldr, int64_t ldo, int flags, int force_bn, void* stream) {
  B200W_API_BEGIN
  GemmDesc d;
  d.A = static_cast<const bf16*>(A_dev);
  d.a_rows = M;
  d.a_cols = K;
  d.batches = 1;
  d.a_row_stride = lda;
  d.W = static_cast<const bf16*>(W_dev);
  d.N = N;
  d.K = K;
  d.w_row_stride = ldw;
  d.out_rows = M;
  d.bias = bias_dev;
  d.act = (flags & 4) ? 1 : 0;
  d.residual = residual_dev;
  d.res_fp32 = (flags & 2) ? 1 : 0;
  d.res_row_stride = ldr;
  d.out = out_dev;
  d.out_fp32 = (flags & 1) ? 1 : 0;
  d.out_row_stride = ldo;
  d.force_bn = force_bn;
  if ((flags >> 8) & 0xff) {  // split-K: out = fp32 partial sums [splits][M][ldo]
    d.split_k = (flags >> 8) & 0xff;
    d.split_stride = (int64_t)M * ldo;
  }
  gemm_bf16(d, static_cast<cudaStream_t>(stream));
  B200W_API_END
}

int b200w_skinny_gemm(const void* X_dev, const void* W_dev, const float* bias_dev, void* out_dev, int M, int N,
                      int K, int64_t ldx, int64_t ldw, int64_t ldo, int flags, int splits, void* stream) {
  B200W_API_BEGIN
  SkinnyDesc d;
  d.X = static_cast<const bf16*>(X_dev);
  d.ldx = ldx;
  d.M = M;
  d.W = static_cast<const bf16*>(W_dev);
  d.ldw = ldw;
  d.N = N;
  d.K = K;
  d.bias = bias_dev;
  d.act = (flags & 4) ? 1 : 0;
  d.out = out_dev;
  d.out_fp32 = (flags & 1) ? 1 : 0;
  d.ldo = ldo;
  d.splits = splits & 0xff;
  d.tile_rows = (splits >> 8) == 64 ? 64 : 128;  // test hook: bits 8.. of `splits` select 64-row tiles
  d.split_stride = (int64_t)M * ldo;
  skinny_gemm(d, static_cast<cudaStream_t>(stream));
  B200W_API_END
}

int b200w_step_self_attention(const float* partial_dev, int n_splits, const float* bias_dev, int rows, int d,
                              int n_heads, const int32_t* slot_dev, void* kpool_dev, void* vpool_dev,
                              const int32_t* table_dev, int table_stride, const int32_t* trow_dev,
                              const int32_t* kvlen_dev, int max_len, int rows_per_group, void* out_dev, void* stream) {
  B200W_API_BEGIN
  QkvAttnParams qa;
  qa.partial = partial_dev; qa.n_splits = n_splits; qa.split_stride = (long long)rows * 3 * d;
  qa.bias = bias_dev; qa.d = d; qa.slot = slot_dev;
  qa.kpool = static_cast<bf16*>(kpool_dev); qa.vpool = static_cast<bf16*>(vpool_dev);
  qa.table = table_dev; qa.table_stride = table_stride; qa.trow = trow_dev; qa.kvlen = kvlen_dev;
  qa.out = static_cast<bf16*>(out_dev); qa.out_stride = d; qa.scale = 0.125f; qa.max_len = max_len;
  (void)rows_per_group;  // (grouping hint of an experimental kernel: the shipped kernel works row by row)
  qkv_attention(qa, rows, n_heads, static_cast<cudaStream_t>(stream));
  B200W_API_END
}

int b200w_sample_topk(const float* logits_dev, int rows, int ld, int n_vocab, const uint8_t* mask_dev,
                      const int32_t* state_dev, int eot, int no_timestamps, int timestamp_begin,
                      int max_initial_ts, int k, int32_t* out_ids_dev, float* out_logprobs_dev, int streaming,
                      float temperature, uint32_t seed, void* stream) {
  B200W_API_BEGIN
  B200W_REQUIRE(ld % 8 == 0 && ld >= n_vocab, "sample_topk: ld must be a multiple of 8 and >= n_vocab");
  SampleParams sp;
  sp.logits = logits_dev; sp.ld = ld; sp.n_vocab = n_vocab; sp.mask = mask_dev;
  sp.state = reinterpret_cast<const int4*>(state_dev);
  sp.eot = eot; sp.no_timestamps = no_timestamps; sp.timestamp_begin = timestamp_begin;
  sp.max_initial_ts = max_initial_ts; sp.k = k;
  sp.out_ids = out_ids_dev; sp.out_logprobs = out_logprobs_dev;
  sp.temperature = temperature; sp.seed = seed;
  sample_topk(sp, rows, static_cast<cudaStream_t>(stream), streaming != 0);
  B200W_API_END
}

int b200w_enc_attention_trace_read(int64_t* host_out, int n) {
  B200W_API_BEGIN
  enc_attention_trace_read(reinterpret_cast<long long*>(host_out), n);
  B200W_API_END
}

int64_t b200w_cross_attention_tc_scratch_bytes(void) { return (int64_t)cross_attention_tc_scratch_bytes(); }

int b200w_cross_attention_tc(const void* cross_kv_dev, int n_streams, int n_layers, int layer, int n_heads, int len,
                             const void* q_dev, const int32_t* stream_dev, const int32_t* g_start_dev,
                             const int32_t* g_count_dev, const int32_t* n_live_dev, int n_groups, void* out_dev,
                             float scale, void* scratch_dev, int pieces, void* stream) {
  B200W_API_BEGIN
  cross_attention_tc(static_cast<const bf16*>(cross_kv_dev), n_streams, n_layers, layer, n_heads, len,
                     static_cast<const bf16*>(q_dev), n_heads * 64, stream_dev, g_start_dev, g_count_dev, n_live_dev,
                     n_groups, static_cast<bf16*>(out_dev), n_heads * 64, scale, scratch_dev,
                     static_cast<cudaStream_t>(stream), pieces);
  B200W_API_END
}

int b200w_cross_attention_flash(const void* cross_kv_dev, int n_streams, int n_layers, int layer, int n_heads, int len,
                                const void* q_dev, int n_q_rows, const int32_t* row_stream_dev,
                                const int32_t* seg_start_dev, const int32_t* seg_count_dev, int n_segs,
                                int max_seg_rows, void* out_dev, void* stream) {
  B200W_API_BEGIN
  cross_attention_flash(static_cast<const bf16*>(cross_kv_dev), n_streams, n_layers, layer, n_heads, len,
                        static_cast<const bf16*>(q_dev), n_q_rows, row_stream_dev, seg_start_dev, seg_count_dev, n_segs,
                        max_seg_rows, static_cast<bf16*>(out_dev), static_cast<cudaStream_t>(stream));
  B200W_API_END
}

int b200w_conv1d_k3(const void* x_dev, const void* w_dev, const float* bias_dev,
                    const float* addend_dev, void* out_dev, int B, int T, int C, int N, int stride,
                    int out_fp32, void* stream) {
  B200W_API_BEGIN
  conv1d_k3(static_cast<const bf16*>(x_dev), static_cast<const bf16*>(w_dev), bias_dev, addend_dev,
            out_dev, B, T, C, N, stride, out_fp32, static_cast<cudaStream_t>(stream));
  B200W_API_END
}

int b200w_enc_attention(const void* qkv_dev, int B, int T, int d, int n_heads, void* out_dev,
                        void* stream) {
  B200W_API_BEGIN
  enc_attention(static_cast<const bf16*>(qkv_dev), B, T, d, n_heads, static_cast<bf16*>(out_dev),
                static_cast<cudaStream_t>(stream));
  B200W_API_END
}

}  // extern "C"

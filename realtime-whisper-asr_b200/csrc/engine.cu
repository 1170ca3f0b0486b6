// Whisper engine: encoder, cross-K/V, autoregressive decode (greedy / beam, Whisper logit rules,
// token-granular paged KV cache), word-timing pass and language detection, orchestrated in C++
// over the kernels of this library.  Replaces the CTranslate2 runtime the reference reaches
// through faster-whisper (/root/reference/asr_components.py:279-288).
//
// KV cache layout ("pages" of one token): per decoder layer two pools K,V bf16 [slots][d].  A
// hypothesis is a row of a page table (int32 [rows][n_text_ctx], position -> slot).  Prompt slots
// are written once per chunk and shared by every beam; a beam step writes one fresh slot per
// hypothesis and forks by copying table rows (ints), never K/V bytes.
#include "../../include/b200w.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "attention.h"
#include "common.h"
#include "dstep.h"
#include "gemm.h"
#include "kernels.h"
#include "skinny.h"

using namespace b200w;

namespace {

constexpr int HEAD_DIM = 64;

struct Weight {
  const void* ptr;
  int64_t numel;
  int dtype;  // 0 fp32, 1 bf16
};

struct Lin {
  const bf16* w = nullptr;
  const float* b = nullptr;
};
struct Norm {
  const float* g = nullptr;
  const float* b = nullptr;
};
struct EncLayer {
  Norm ln1, ln2;
  Lin qkv, out, fc1, fc2;
};
struct DecLayer {
  Norm ln1, ln2, ln3;
  Lin qkv, out, cq, cout, fc1, fc2;
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  void alloc(size_t n) {
    if (n > bytes) {
      if (p) cudaFree(p);
      p = nullptr;
      B200W_CHECK_CUDA(cudaMalloc(&p, n));
      bytes = n;
    }
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

struct PinnedBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool grew = false;  // set when the address changed (captured graphs must be dropped)
  void alloc(size_t n) {
    if (n > bytes) {
      grew = p != nullptr;
      if (p) cudaFreeHost(p);
      p = nullptr;
      B200W_CHECK_CUDA(cudaMallocHost(&p, n));
      bytes = n;
    }
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

// Optional per-op timing of the decoder pass (B200W_DECODE_PROFILE=1; eager launches only): CUDA events
// around every op, summed per op type and printed when the decode call returns.
struct OpTimer {
  bool on = false;
  bool forced = false;  // b200w_op_profile_enable: decode steps run eagerly so that every op can be bracketed
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { int id; cudaEvent_t a, b; double work; };
  std::vector<Rec> recs;
  double tot_ms[16] = {0}, tot_work[16] = {0};
  long long tot_n[16] = {0};
  double next_work = 0;  // algorithmic bytes / flops of the op about to be recorded (set by the caller)
  const char* names[16] = {"embed", "layernorm", "qkv_gemm", "kv_scatter", "self_attn", "out_gemm", "cq_gemm",
                           "cross_attn", "cout_gemm", "fc1_gemm", "fc2_gemm", "logits_gemm", "sample", "fork",
                           "final_ln", "other"};
  cudaEvent_t ev() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
  void collect(const char* tag) {
    if (!on || recs.empty()) return;
    double tot[16] = {0};
    int cnt[16] = {0};
    for (auto& r : recs) {
      float ms = 0;
      cudaEventSynchronize(r.b);
      cudaEventElapsedTime(&ms, r.a, r.b);
      tot[r.id] += ms;
      cnt[r.id]++;
      tot_ms[r.id] += ms;
      tot_work[r.id] += r.work;
      tot_n[r.id]++;
    }
    if (getenv("B200W_DECODE_PROFILE")) {
      fprintf(stderr, "[b200w op profile] %s\n", tag);
      for (int i = 0; i < 16; ++i)
        if (cnt[i]) fprintf(stderr, "   %-12s n=%6d total %9.3f ms  avg %8.2f us\n", names[i], cnt[i], tot[i], 1000.0 * tot[i] / cnt[i]);
    }
    recs.clear();
    used = 0;
  }
};
}  // namespace

struct b200w_engine {
  b200w_model_config cfg{};
  std::unordered_map<std::string, Weight> weights;
  bool finalized = false;

  // resolved weights
  Lin conv1, conv2;
  const float* enc_pos = nullptr;
  std::vector<EncLayer> enc;
  Norm enc_ln_post;
  const bf16* tok_emb = nullptr;
  const float* dec_pos = nullptr;
  std::vector<DecLayer> dec;
  Lin cross_kv;
  Norm dec_ln;

  int d = 0, H = 0, Le = 0, Ld = 0, V = 0, ldv = 0, Tctx = 0, Tenc = 0;
  int n_max = 0;       // max decoder tokens per forward pass
  int slots_per_stream = 0;
  int logit_rows = 0;  // capacity of the logits buffer (rows)

  // workspaces
  DevBuf e_x, e_h, e_qkv, e_attn, e_mlp;              // encoder
  DevBuf d_x, d_h, d_qkv, d_attn, d_cq, d_mlp, d_hsel;  // decoder
  DevBuf kpool, vpool;                                // [Ld][slots][d] bf16 each
  DevBuf tables[2];                                   // int32 [rows][Tctx]
  DevBuf logits;                                      // fp32 [logit_rows][ldv]
  DevBuf meta;                                        // per-pass int metadata
  DevBuf cross_off, const_len, mask, topk_ids, topk_lp, probs, head_export;
  DevBuf al_w, al_mat, al_trace, al_path, al_plen, scratch_i;
  DevBuf x_split;        // split-KV partial results of the decode-step cross attention
  DevBuf x_tc;           // tensor-core cross attention: split-item partial results + arrival counters
  DevBuf d_split;        // fp32 split-K partial sums of the GEMM feeding the residual stream
  SplitReduce pending;   // partial sums of the last decoder GEMM, folded in by the next LayerNorm
  PinnedBuf h_meta, h_out;
  long long total_slots = 0;

  // decode-step CUDA graphs, keyed by everything that is baked into the captured launches
  struct StepGraph {
    const void* ckv;
    int B, K, parity;
    int opts_hash;
    cudaGraphExec_t exec;
  };
  std::vector<StepGraph> step_graphs;
  bool use_graphs = true;
  bool step_pass = false;  // the pass being enqueued is a decode step (selects the skinny GEMM path)
  unsigned long long launches_per_step = 0;
  cudaStream_t own_stream = nullptr;  // decode runs here (the legacy default stream cannot be captured)
  cudaEvent_t join_event = nullptr;
  // in-graph probe: events around the decode cross attention of one layer, every decode step (bench roofline)
  cudaEvent_t xev0 = nullptr, xev1 = nullptr;
  bool xev_recorded = false;
  double probe_ms = 0, probe_bytes = 0;  // totals of that probe since the last reset
  long long probe_n = 0;
  OpTimer optimer;        // per-op event timing of eager decode passes (profiling aid)
  std::mutex mu;          // one host thread at a time inside an engine call (workspaces are per engine)
  SkinnyCache* skinny_cache = nullptr;  // launch records of this engine's decode-step GEMMs (freed with it)
  DstepCache* dstep_cache = nullptr;    // records + grid-barrier words of the fused decode-step segments
};

namespace {

const void* need(b200w_engine* e, const std::string& name, int64_t numel, int dtype) {
  auto it = e->weights.find(name);
  if (it == e->weights.end()) throw Error("missing weight: " + name);
  if (it->second.numel != numel || it->second.dtype != dtype) {
    char buf[256];
    snprintf(buf, sizeof(buf), "weight %s: expected numel %lld dtype %d, got %lld dtype %d", name.c_str(),
             (long long)numel, dtype, (long long)it->second.numel, it->second.dtype);
    throw Error(buf);
  }
  if (reinterpret_cast<uintptr_t>(it->second.ptr) & 15) throw Error("weight " + name + " is not 16-byte aligned");
  return it->second.ptr;
}
Lin need_lin(b200w_engine* e, const std::string& p, int64_t n, int64_t k) {
  Lin l;
  l.w = static_cast<const bf16*>(need(e, p + ".w", n * k, 1));
  l.b = static_cast<const float*>(need(e, p + ".b", n, 0));
  return l;
}
Norm need_norm(b200w_engine* e, const std::string& p, int64_t d) {
  Norm n;
  n.g = static_cast<const float*>(need(e, p + ".g", d, 0));
  n.b = static_cast<const float*>(need(e, p + ".b", d, 0));
  return n;
}

void linear(const bf16* A, int M, int K, const Lin& l, int N, void* out, int64_t ldo, int out_fp32, int act,
            const void* residual, int res_fp32, int64_t ldr, cudaStream_t s) {
  GemmDesc g;
  g.A = A;
  g.a_rows = M;
  g.a_cols = K;
  g.a_row_stride = K;
  g.W = l.w;
  g.N = N;
  g.K = K;
  g.w_row_stride = K;
  g.out_rows = M;
  g.bias = l.b;
  g.act = act;
  g.residual = residual;
  g.res_fp32 = res_fp32;
  g.res_row_stride = ldr;
  g.out = out;
  g.out_fp32 = out_fp32;
  g.out_row_stride = ldo;
  gemm_bf16(g, s);
}

struct OpScope {
  OpTimer& t;
  int id;
  cudaStream_t s;
  cudaEvent_t a = nullptr;
  double work = 0;
  OpScope(OpTimer& t_, int id_, cudaStream_t s_) : t(t_), id(id_), s(s_) {
    if (t.on) { a = t.ev(); work = t.next_work; t.next_work = 0; cudaEventRecord(a, s); }
  }
  ~OpScope() {
    if (a) { cudaEvent_t b = t.ev(); cudaEventRecord(b, s); t.recs.push_back({id, a, b, work}); }
  }
};
#define OP(id) OpScope _op_scope_##id(e->optimer, id, s)
// Step probe: CUDA events around the tensor-core cross attention of ONE decoder layer in every decode step --
// inside the captured step graph as external event-record nodes -- so that bench.py can time the dominant
// kernel inside its timed region (a graph replay cannot be bracketed from outside).  Totals since the last reset.
static void probe_record(cudaEvent_t ev, cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  B200W_CHECK_CUDA(cudaStreamIsCapturing(s, &st));
  if (st == cudaStreamCaptureStatusActive) B200W_CHECK_CUDA(cudaEventRecordWithFlags(ev, s, cudaEventRecordExternal));
  else B200W_CHECK_CUDA(cudaEventRecord(ev, s));
}
// Development aid: B200W_ABLATE=<bit mask over the op ids above> drops those ops from the decoder pass
// (results are garbage) so that an op's real cost inside the captured step can be read off the step time.
static unsigned ablate_mask() {
  static const unsigned m = getenv("B200W_ABLATE") ? (unsigned)strtoul(getenv("B200W_ABLATE"), nullptr, 0) : 0u;
  return m;
}
#define RUN(id) if (!((ablate_mask() >> (id)) & 1u))

// x += A W^T + b for the decoder's residual stream.  Skinny calls (decode steps) are split along K so
// that enough CTAs stream the weights; the partial sums are handed to the next LayerNorm.
// Decode steps (at most 256 rows): weight-row-major tcgen05 GEMM (skinny.cu), split along K into fp32 slabs
// in e->d_split that the consumer reduces.  Returns the slab description.
// Key-range pieces per (stream, head) item of the decode-step cross attention: a constant (never a function of
// the batch), so results are batch-invariant.  Measured on the bench workload (640 items, 296 CTAs; ms per decode
// step, profiles/r02_cross_pieces.txt): 1 piece 3.99, 2: 4.33, 3: 4.49, 4: 4.66, 6: 5.23 -- every extra unit costs a
// softmax bubble in the load stream, which outweighs the better balance, so items are computed whole.
static int cross_step_pieces() {
  static const int n = getenv("B200W_XT_PIECES") ? atoi(getenv("B200W_XT_PIECES")) : 1;
  return n;
}
static bool use_skinny() {
  static const bool on = getenv("B200W_NO_SKINNY") == nullptr;
  return on;
}
SplitReduce skinny_split(b200w_engine* e, const bf16* A, int M, int K, const Lin& l, int N, int max_splits,
                         cudaStream_t s) {
  const int splits = std::max(2, skinny_suggest_splits(N, K, max_splits));
  B200W_REQUIRE((size_t)splits * M * N * 4 <= e->d_split.bytes, "split buffer too small");
  // more than 256 rows: one launch per 256-row chunk (each row's arithmetic does not depend on the chunking)
  for (int r0 = 0; r0 < M; r0 += 256) {
    SkinnyDesc g;
    g.X = A + (long long)r0 * K; g.ldx = K; g.M = std::min(256, M - r0);
    g.W = l.w; g.ldw = K; g.N = N; g.K = K;
    g.out = e->d_split.as<float>() + (long long)r0 * N; g.out_fp32 = 1; g.ldo = N;
    g.splits = splits; g.split_stride = (int64_t)M * N;
    g.cache = e->skinny_cache;
    skinny_gemm(g, s);
  }
  SplitReduce r;
  r.partial = e->d_split.as<float>();
  r.n = splits;
  r.stride = (long long)M * N;
  r.bias = l.b;
  return r;
}

SplitReduce residual_linear(b200w_engine* e, const bf16* A, int M, int K, const Lin& l, float* x, cudaStream_t s) {
  const int d = e->d;
  // measured (tools/gemm_skinny_bench.py): 7 splits for K = d, 10 for K = 4d; more only adds slab traffic
  if (use_skinny() && e->step_pass && d % 128 == 0) return skinny_split(e, A, M, K, l, d, K > 2 * d ? 10 : 7, s);
  // prompt / alignment / language passes: one unsplit GEMM with the residual in its epilogue.  (A split-K factor
  // chosen from the row count would make a stream's rounding depend on how many streams share the pass.)
  linear(A, M, K, l, d, x, d, 1, 0, x, 1, d, s);
  return SplitReduce{};
}

// ----------------------------------------------------------------------------------- decoder pass
struct PassMeta {  // device pointers into e->meta for one forward pass of n tokens
  int n = 0;
  int *tok, *pos, *slot, *trow, *kvlen, *stream, *parent;
  int *g_start, *g_count;  // cross-attention groups: runs of <= 8 rows of one stream (live groups first)
  int* n_live = nullptr;   // device: number of groups with at least one live row
  int n_live_host = 0;
  int *seg_start = nullptr, *seg_count = nullptr;  // runs of consecutive rows of one stream (flash cross attention)
  int n_segs = 0, max_seg = 0;
  int n_groups = 0, max_group = 1;
  int4* state;
};

// host-side staging of the metadata: [tok|pos|slot|trow|kvlen|stream] x n ints, then state int4 x n
struct HostMeta {
  std::vector<int> tok, pos, slot, trow, kvlen, stream, parent;
  std::vector<char> idle;  // fixed-shape decode steps: rows that only pad the batch (no cross attention)
  std::vector<int4> state;
  void clear() {
    tok.clear(); pos.clear(); slot.clear(); trow.clear(); kvlen.clear(); stream.clear(); parent.clear();
    idle.clear();
    state.clear();
  }
  int n() const { return (int)tok.size(); }
  void push(int t, int p, int s, int tr, int kl, int st, int4 sstate = make_int4(0, 0, 0, 0), int par = 0,
            bool is_idle = false) {
    tok.push_back(t); pos.push_back(p); slot.push_back(s); trow.push_back(tr); kvlen.push_back(kl);
    stream.push_back(st); parent.push_back(par); state.push_back(sstate); idle.push_back(is_idle ? 1 : 0);
  }
};

// Stage the metadata of one pass in the pinned buffer and return the device-side view.  The layout
// depends only on (n, number of groups), so a decode step of fixed shape always produces the same
// pointers -- which is what lets the step be replayed as a CUDA graph.
PassMeta fill_host_meta(b200w_engine* e, const HostMeta& h, size_t* copy_bytes) {
  const int n = h.n();
  B200W_REQUIRE(n <= e->n_max, "too many decoder tokens in one pass");
  // cross-attention groups: consecutive rows of the same stream, at most CROSS_GROUP per group
  std::vector<int> gs, gc;
  for (int i = 0; i < n;) {
    int j = i + 1;
    while (j < n && j - i < CROSS_GROUP && h.stream[j] == h.stream[i]) ++j;
    // idle rows sit behind the live rows of their stream: the group covers the same rows every step
    // (fixed shape) but only its leading live rows are attended, and a fully idle group reads no K/V
    int live = 0;
    while (i + live < j && !h.idle[i + live]) ++live;
    for (int r = i + live; r < j; ++r) B200W_REQUIRE(h.idle[r], "live decoder row behind an idle one");
    gs.push_back(i);
    gc.push_back(live);
    i = j;
  }
  const int ng = (int)gs.size();
  int max_cap = 1;
  for (int g = 0; g < ng; ++g) max_cap = std::max(max_cap, (g + 1 < ng ? gs[g + 1] : n) - gs[g]);
  // live groups first (the order of the groups is free): the tensor-core kernel walks [0, n_live)
  int n_live = 0;
  {
    std::vector<int> gs2, gc2;
    for (int pass = 0; pass < 2; ++pass)
      for (int g = 0; g < ng; ++g)
        if ((gc[g] > 0) == (pass == 0)) { gs2.push_back(gs[g]); gc2.push_back(gc[g]); }
    for (int g = 0; g < ng; ++g) n_live += gc[g] > 0;
    gs.swap(gs2);
    gc.swap(gc2);
  }
  // segments: maximal runs of consecutive rows of one stream (idle padding rows included: only many-row passes,
  // which have none, use them)
  std::vector<int> ss, sc;
  for (int i = 0; i < n;) {
    int j = i + 1;
    while (j < n && h.stream[j] == h.stream[i]) ++j;
    ss.push_back(i);
    sc.push_back(j - i);
    i = j;
  }
  const int nseg = (int)ss.size();
  const size_t ints = (size_t)n * 7 + 2 * (size_t)ng + 1 + 2 * (size_t)nseg, bytes = ints * 4 + (size_t)n * 16 + 16;
  e->h_meta.alloc(bytes);
  int* hp = e->h_meta.as<int>();
  memcpy(hp, h.tok.data(), n * 4);
  memcpy(hp + n, h.pos.data(), n * 4);
  memcpy(hp + 2 * n, h.slot.data(), n * 4);
  memcpy(hp + 3 * n, h.trow.data(), n * 4);
  memcpy(hp + 4 * n, h.kvlen.data(), n * 4);
  memcpy(hp + 5 * n, h.stream.data(), n * 4);
  memcpy(hp + 6 * n, h.parent.data(), n * 4);
  memcpy(hp + 7 * n, gs.data(), ng * 4);
  memcpy(hp + 7 * n + ng, gc.data(), ng * 4);
  hp[7 * n + 2 * ng] = n_live;
  memcpy(hp + 7 * n + 2 * ng + 1, ss.data(), nseg * 4);
  memcpy(hp + 7 * n + 2 * ng + 1 + nseg, sc.data(), nseg * 4);
  const size_t st_off = (ints * 4 + 15) / 16 * 16;
  memcpy(reinterpret_cast<char*>(hp) + st_off, h.state.data(), (size_t)n * 16);
  *copy_bytes = st_off + (size_t)n * 16;
  PassMeta m;
  m.n = n;
  int* dp = e->meta.as<int>();
  m.tok = dp; m.pos = dp + n; m.slot = dp + 2 * n; m.trow = dp + 3 * n; m.kvlen = dp + 4 * n;
  m.stream = dp + 5 * n;
  m.parent = dp + 6 * n;
  m.g_start = dp + 7 * n;
  m.g_count = dp + 7 * n + ng;
  m.n_groups = ng;
  m.n_live = dp + 7 * n + 2 * ng;
  m.n_live_host = n_live;
  m.seg_start = dp + 7 * n + 2 * ng + 1;
  m.seg_count = m.seg_start + nseg;
  m.n_segs = nseg;
  m.max_seg = 0;
  for (int c : sc) m.max_seg = std::max(m.max_seg, c);
  m.max_group = max_cap;  // capacity, not the live count: keeps the launch shape fixed
  m.state = reinterpret_cast<int4*>(reinterpret_cast<char*>(dp) + st_off);
  return m;
}

PassMeta upload_meta(b200w_engine* e, const HostMeta& h, cudaStream_t s) {
  size_t bytes = 0;
  PassMeta m = fill_host_meta(e, h, &bytes);
  B200W_CHECK_CUDA(cudaMemcpyAsync(e->meta.p, e->h_meta.p, bytes, cudaMemcpyHostToDevice, s));
  return m;
}

struct AlignExport {
  float* probs = nullptr;           // [n_align][n][Tenc]
  const int* head_table = nullptr;  // device [Ld][H] -> export index or -1
  std::vector<char> layer_has;      // host: any exported head in layer l
};

// Run every decoder layer over n tokens; hidden state (pre final-LN) ends up in e->d_x (fp32).
void decoder_forward(b200w_engine* e, const bf16* cross_kv, const PassMeta& m, const int* table,
                     int max_kvlen, const AlignExport* ax, cudaStream_t s) {
  const int n = m.n, d = e->d;
  float* x = e->d_x.as<float>();
  bf16* h = e->d_h.as<bf16>();
  bf16* qkv = e->d_qkv.as<bf16>();
  bf16* attn = e->d_attn.as<bf16>();
  bf16* cq = e->d_cq.as<bf16>();
  bf16* mlp = e->d_mlp.as<bf16>();
  RUN(0) { OP(0); embed_tokens(m.tok, m.pos, e->tok_emb, e->dec_pos, x, n, d, s); }
  SplitReduce pend;
  for (int l = 0; l < e->Ld; ++l) {
    const DecLayer& L = e->dec[l];
    bf16* kp = e->kpool.as<bf16>() + (long long)l * e->total_slots * d;
    bf16* vp = e->vpool.as<bf16>() + (long long)l * e->total_slots * d;
    // self attention
    RUN(1) { OP(1); layernorm(x, nullptr, L.ln1.g, L.ln1.b, h, n, d, s, &pend); }
    // decode steps only: the choice must not depend on how many streams are batched (batch invariance)
    const bool sk = use_skinny() && e->step_pass && d % 128 == 0;
    // reduction of the QKV slabs + cache write + self-attention in one kernel (needs heads % 4 == 0)
    static const bool use_fused = getenv("B200W_NO_FUSED_ATTN") == nullptr;
    const bool fused_attn = sk && use_fused && e->H % 4 == 0 && !((ablate_mask() >> 2) & 1u);
    if (sk) {
      // split-K QKV projection; the reduction adds the bias, rounds to bf16 and scatters K / V into the cache
      RUN(2) {
        OP(2);
        const SplitReduce r = skinny_split(e, h, n, d, L.qkv, 3 * d, 4, s);
        if (fused_attn) {
          QkvAttnParams qa;
          qa.partial = r.partial; qa.n_splits = r.n; qa.split_stride = r.stride; qa.bias = r.bias; qa.d = d;
          qa.slot = m.slot; qa.kpool = kp; qa.vpool = vp;
          qa.table = table; qa.table_stride = e->Tctx; qa.trow = m.trow; qa.kvlen = m.kvlen;
          qa.out = attn; qa.out_stride = d; qa.scale = 0.125f; qa.max_len = max_kvlen;
          qkv_attention(qa, n, e->H, s);
        } else {
          split_finalize(r.partial, r.n, r.stride, r.bias, n, 3 * d, 0, qkv, 3 * d, m.slot, kp, vp, d, s);
        }
      }
    } else {
      RUN(2) { OP(2); linear(h, n, d, L.qkv, 3 * d, qkv, 3 * d, 0, 0, nullptr, 0, 0, s); }
      RUN(3) { OP(3); kv_scatter(qkv, m.slot, kp, vp, n, d, s); }
    }
    DecAttnParams a;
    a.q = qkv; a.q_stride = 3 * d;
    a.kbase = kp; a.vbase = vp; a.kv_row_stride = d;
    a.table = table; a.table_stride = e->Tctx; a.trow = m.trow; a.kvlen = m.kvlen;
    a.out = attn; a.out_stride = d; a.scale = 0.125f; a.n_queries = n; a.max_len = max_kvlen;
    if (!fused_attn) RUN(4) { OP(4); dec_attention(a, e->H, s); }
    RUN(5) { OP(5); pend = residual_linear(e, attn, n, d, L.out, x, s); }
    // cross attention
    RUN(1) { OP(1); layernorm(x, nullptr, L.ln2.g, L.ln2.b, h, n, d, s, &pend); }
    RUN(6) {
      OP(6);
      if (sk) {
        const SplitReduce r = skinny_split(e, h, n, d, L.cq, d, 7, s);
        split_finalize(r.partial, r.n, r.stride, r.bias, n, d, 0, cq, d, nullptr, nullptr, nullptr, 0, s);
      } else {
        linear(h, n, d, L.cq, d, cq, d, 0, 0, nullptr, 0, 0, s);
      }
    }
    CrossAttnParams c;
    c.q = cq; c.q_stride = d;
    // cross K/V layout [stream][layer][k|v][head][1500][64]
    c.kbase = cross_kv + (long long)(2 * l) * e->Tenc * d;
    c.vbase = cross_kv + (long long)(2 * l + 1) * e->Tenc * d;
    c.kv_row_stride = HEAD_DIM;
    c.head_stride = (long long)e->Tenc * HEAD_DIM;
    c.kv_base_off = e->cross_off.as<long long>();
    c.stream = m.stream;
    c.g_start = m.g_start; c.g_count = m.g_count;
    c.len = e->Tenc;
    c.out = attn; c.out_stride = d; c.scale = 0.125f; c.n_queries = n;
    if (ax && ax->layer_has[l]) {
      c.head_export = ax->head_table + (long long)l * e->H;
      c.probs_out = ax->probs;
      c.probs_stride = e->Tenc;
    }
    static const bool use_tc = getenv("B200W_NO_CROSS_TC") == nullptr;
    RUN(7) {
      // Tensor-core kernel for every pass whose groups fit it (decode steps, prompts, alignment).  Heads whose
      // probabilities are exported (alignment pass) are then recomputed by the CUDA-core kernel, which writes
      // the probabilities and overwrites those heads' outputs (export_only skips every other head).
      const bool step_tc = use_tc && !c.head_export && m.max_group <= CROSS_GROUP && m.n_groups <= 64;
      const bool tc = use_tc && m.max_group <= CROSS_GROUP;
      // decode steps are op 7, with their algorithmic bytes: K and V of every live (group, head), read once;
      // prompt / alignment passes are timed under "other"
      if (step_tc) e->optimer.next_work = (double)m.n_live_host * e->H * 2.0 * e->Tenc * HEAD_DIM * 2.0;
      OpScope _op_scope_x(e->optimer, step_tc ? 7 : 15, s);
      // many rows per stream (prompts, alignment): flash kernel, K/V streamed once per (stream, head)
      // (every non-step pass, whatever its segment lengths: the kernel choice must not depend on the batch)
      static const bool use_flash = getenv("B200W_NO_CROSS_FLASH") == nullptr;
      if (use_flash && !e->step_pass) {
        cross_attention_flash(cross_kv, e->cfg.max_batch, e->Ld, l, e->H, e->Tenc, cq, n, m.stream, m.seg_start,
                              m.seg_count, m.n_segs, m.max_seg, attn, s);
        if (c.head_export) {
          c.export_only = 1;
          cross_attention(c, m.n_groups, e->H, m.max_group, s, nullptr);
        }
      } else if (tc) {
        const bool probe = step_tc && e->step_pass && l == e->Ld / 2 && !e->optimer.on;
        if (probe) probe_record(e->xev0, s);
        cross_attention_tc(cross_kv, e->cfg.max_batch, e->Ld, l, e->H, e->Tenc, cq, d, m.stream, m.g_start, m.g_count,
                           m.n_live, m.n_groups, attn, d, 0.125f, e->x_tc.p, s, e->step_pass ? cross_step_pieces() : 1);
        if (probe) { probe_record(e->xev1, s); e->xev_recorded = true; }
        if (c.head_export) {
          c.export_only = 1;
          cross_attention(c, m.n_groups, e->H, m.max_group, s, nullptr);
        }
      } else {
        cross_attention(c, m.n_groups, e->H, m.max_group, s, e->x_split.as<float>());
      }
    }
    RUN(8) { OP(8); pend = residual_linear(e, attn, n, d, L.cout, x, s); }
    // MLP
    RUN(1) { OP(1); layernorm(x, nullptr, L.ln3.g, L.ln3.b, h, n, d, s, &pend); }
    RUN(9) {
      OP(9);
      if (sk) {
        // split-K too: 40 weight tiles alone leave most SMs idle; the reduction applies bias + GELU
        const SplitReduce r = skinny_split(e, h, n, d, L.fc1, 4 * d, 3, s);
        split_finalize(r.partial, r.n, r.stride, r.bias, n, 4 * d, 1, mlp, 4 * d, nullptr, nullptr, nullptr, 0, s);
      } else {
        linear(h, n, d, L.fc1, 4 * d, mlp, 4 * d, 0, 1, nullptr, 0, 0, s);
      }
    }
    RUN(10) { OP(10); pend = residual_linear(e, mlp, n, 4 * d, L.fc2, x, s); }
  }
  e->pending = pend;  // folded in by the final LayerNorm (logits_for_rows)
}

// final LayerNorm on selected rows + tied-embedding logits -> e->logits[0 .. n_sel)
void logits_for_rows(b200w_engine* e, const int* sel_rows_dev, int n_sel, cudaStream_t s) {
  B200W_REQUIRE(n_sel <= e->logit_rows, "logits buffer too small");
  bf16* hsel = e->d_hsel.as<bf16>();
  RUN(14) { OP(14); layernorm(e->d_x.as<float>(), sel_rows_dev, e->dec_ln.g, e->dec_ln.b, hsel, n_sel, e->d, s, &e->pending); }
  if (!sel_rows_dev) e->pending = SplitReduce{};  // written back into x: must not be added twice
  Lin emb;
  emb.w = e->tok_emb;
  emb.b = nullptr;
  RUN(11) {
    OP(11);
    if (use_skinny() && e->step_pass && e->d % 128 == 0) {
      // decode steps: the 133 MB embedding matrix is streamed once (the many-row kernel would stream it once per
      // 128-row tile), 406 weight tiles of 128 vocabulary rows
      for (int r0 = 0; r0 < n_sel; r0 += 256) {
        SkinnyDesc g;
        g.X = hsel + (long long)r0 * e->d; g.ldx = e->d; g.M = std::min(256, n_sel - r0);
        g.W = emb.w; g.ldw = e->d; g.N = e->V; g.K = e->d;
        g.out = e->logits.as<float>() + (long long)r0 * e->ldv; g.out_fp32 = 1; g.ldo = e->ldv;
        g.cache = e->skinny_cache;
        skinny_gemm(g, s);
      }
    } else {
      linear(hsel, n_sel, e->d, emb, e->V, e->logits.p, e->ldv, 1, 0, nullptr, 0, 0, s);
    }
  }
}

// ----------------------------------------------------------------------------------- fused decode step
// Decode steps with the segments between the attentions fused into persistent program kernels (dstep.cu):
//   K0            [embed, LN1, QKV]                                  (layer 0)
//   per layer     qkv_attention (split-K reduce + cache write + self-attention)
//                 A [out, LN2, cq, finalize]
//                 cross_attention_tc
//                 B [cout, LN3, fc1, GELU finalize, fc2] + ([LN1, QKV] of the next layer | [final LN, logits])
// 4 launches per layer instead of 13; same split factors, k ranges and summation orders as the launch-per-op path.
static bool use_fused_step() {
  // Opt-in (B200W_DSTEP=1).  Measured on the bench workload (profiles/r02_dstep_fused.txt): identical tokens, but
  // 5.6 ms per decode step against 3.97 ms for the launch-per-op step -- a software grid barrier costs ~3.3 us
  // (store drain + atomic + poll) where a PDL-chained launch costs ~1.5-2 us, so fusing removes no latency.
  static const bool on = getenv("B200W_DSTEP") != nullptr && atoi(getenv("B200W_DSTEP")) != 0;
  return on;
}
static bool fused_step_ok(const b200w_engine* e, int n) {
  return use_fused_step() && use_skinny() && e->d % 128 == 0 && e->H % 4 == 0 && dstep_supported(n, e->d) &&
         (size_t)16 * n * e->d * 4 <= e->d_split.bytes && n <= e->logit_rows;
}

void decoder_step_fused(b200w_engine* e, const bf16* cross_kv, const PassMeta& m, const int* table, int max_kvlen,
                        cudaStream_t s) {
  const int n = m.n, d = e->d;
  float* x = e->d_x.as<float>();
  bf16* h = e->d_h.as<bf16>();
  bf16* attn = e->d_attn.as<bf16>();
  bf16* cq = e->d_cq.as<bf16>();
  bf16* mlp = e->d_mlp.as<bf16>();
  bf16* hsel = e->d_hsel.as<bf16>();
  float* S = e->d_split.as<float>();
  DsPhaseDesc ph[DS_MAX_PHASES];
  auto reset = [&]() {
    memset(static_cast<void*>(ph), 0, sizeof(ph));
    for (auto& p : ph) { p.rows = n; p.splits = 1; }
  };
  auto gemm = [&](DsPhaseDesc& p, const bf16* X, const Lin& l, int N, int K, int max_splits) {
    p.type = DS_GEMM; p.X = X; p.W = l.w; p.N = N; p.K = K;
    p.splits = std::max(2, skinny_suggest_splits(N, K, max_splits));
    p.slabs = S; p.ldo = N;
    return p.splits;
  };
  auto ln = [&](DsPhaseDesc& p, const Norm& nm, int n_slabs, const float* bias, bf16* y) {
    p.type = DS_LN; p.x = x; p.gamma = nm.g; p.beta = nm.b; p.y = y; p.d = d;
    p.n_slabs = n_slabs; p.slabs = n_slabs ? S : nullptr; p.bias = n_slabs ? bias : nullptr;
  };
  auto fin = [&](DsPhaseDesc& p, int n_slabs, const float* bias, int N, int act, bf16* out) {
    p.type = DS_FIN; p.slabs = S; p.n_slabs = n_slabs; p.bias = bias; p.N = N; p.act = act; p.out = out;
  };
  // K0: embedding, first LayerNorm, first QKV projection
  int qkv_splits;
  RUN(2) {
    OP(2);
    reset();
    ph[0].type = DS_EMBED; ph[0].tok = m.tok; ph[0].pos = m.pos; ph[0].tok_emb = e->tok_emb; ph[0].pos_emb = e->dec_pos;
    ph[0].x = x; ph[0].d = d;
    ln(ph[1], e->dec[0].ln1, 0, nullptr, h);
    qkv_splits = gemm(ph[2], h, e->dec[0].qkv, 3 * d, d, 4);
    dstep_run(e->dstep_cache, 128 * n, ph, 3, s);
  }
  qkv_splits = std::max(2, skinny_suggest_splits(3 * d, d, 4));
  for (int l = 0; l < e->Ld; ++l) {
    const DecLayer& L = e->dec[l];
    bf16* kp = e->kpool.as<bf16>() + (long long)l * e->total_slots * d;
    bf16* vp = e->vpool.as<bf16>() + (long long)l * e->total_slots * d;
    RUN(4) {
      OP(4);
      QkvAttnParams qa;
      qa.partial = S; qa.n_splits = qkv_splits; qa.split_stride = (long long)n * 3 * d; qa.bias = L.qkv.b; qa.d = d;
      qa.slot = m.slot; qa.kpool = kp; qa.vpool = vp;
      qa.table = table; qa.table_stride = e->Tctx; qa.trow = m.trow; qa.kvlen = m.kvlen;
      qa.out = attn; qa.out_stride = d; qa.scale = 0.125f; qa.max_len = max_kvlen;
      qkv_attention(qa, n, e->H, s);
    }
    RUN(6) {
      OP(6);
      reset();
      const int so = gemm(ph[0], attn, L.out, d, d, 7);
      ln(ph[1], L.ln2, so, L.out.b, h);
      const int sq = gemm(ph[2], h, L.cq, d, d, 7);
      fin(ph[3], sq, L.cq.b, d, 0, cq);
      dstep_run(e->dstep_cache, 128 * n + 1 + 2 * l, ph, 4, s);
    }
    RUN(7) {
      e->optimer.next_work = (double)m.n_live_host * e->H * 2.0 * e->Tenc * HEAD_DIM * 2.0;
      OpScope _op_scope_x(e->optimer, 7, s);
      const bool probe = l == e->Ld / 2 && !e->optimer.on;
      if (probe) probe_record(e->xev0, s);
      cross_attention_tc(cross_kv, e->cfg.max_batch, e->Ld, l, e->H, e->Tenc, cq, d, m.stream, m.g_start, m.g_count,
                         m.n_live, m.n_groups, attn, d, 0.125f, e->x_tc.p, s, cross_step_pieces());
      if (probe) { probe_record(e->xev1, s); e->xev_recorded = true; }
    }
    RUN(9) {
      OP(9);
      reset();
      const int sc = gemm(ph[0], attn, L.cout, d, d, 7);
      ln(ph[1], L.ln3, sc, L.cout.b, h);
      const int s1 = gemm(ph[2], h, L.fc1, 4 * d, d, 3);
      fin(ph[3], s1, L.fc1.b, 4 * d, 1, mlp);
      const int s2 = gemm(ph[4], mlp, L.fc2, d, 4 * d, 10);
      if (l + 1 < e->Ld) {
        ln(ph[5], e->dec[l + 1].ln1, s2, L.fc2.b, h);
        gemm(ph[6], h, e->dec[l + 1].qkv, 3 * d, d, 4);
      } else {  // final LayerNorm + tied-embedding logits (one slab = the logits buffer itself, row stride ldv)
        ln(ph[5], e->dec_ln, s2, L.fc2.b, hsel);
        Lin emb;
        emb.w = e->tok_emb;
        gemm(ph[6], hsel, emb, e->V, d, 1);
        ph[6].splits = 1;
        ph[6].slabs = e->logits.as<float>();
        ph[6].ldo = e->ldv;
      }
      dstep_run(e->dstep_cache, 128 * n + 2 + 2 * l, ph, 7, s);
    }
  }
  e->pending = SplitReduce{};
}

void upload_ints(int* dst, const std::vector<int>& v, cudaStream_t s) {
  if (v.empty()) return;
  B200W_CHECK_CUDA(cudaMemcpyAsync(dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice, s));
}

}  // namespace

extern "C" {

int b200w_engine_create(const b200w_model_config* cfg, b200w_engine** out) {
  B200W_API_BEGIN
  B200W_REQUIRE(cfg && out, "null argument");
  B200W_REQUIRE(cfg->d_model % 128 == 0 && cfg->d_model <= 1280 && cfg->n_heads * 64 == cfg->d_model,
                "d_model must be a multiple of 128 (<= 1280) with 64-wide heads");
  B200W_REQUIRE(cfg->n_mels == 80 || cfg->n_mels == 128, "n_mels must be 80 or 128");
  B200W_REQUIRE(cfg->n_audio_ctx == 1500 && cfg->n_text_ctx >= 8 && cfg->n_text_ctx <= 448, "context sizes");
  B200W_REQUIRE(cfg->max_batch >= 1 && cfg->max_beam >= 1 && 2 * cfg->max_beam <= MAX_TOPK, "batch / beam limits");
  int ndev = 0;
  B200W_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  B200W_REQUIRE(ndev > 0, "no CUDA device (there is no CPU fallback)");
  b200w_engine* e = new b200w_engine();
  e->cfg = *cfg;
  e->skinny_cache = skinny_cache_create();
  e->dstep_cache = dstep_cache_create();
  e->d = cfg->d_model; e->H = cfg->n_heads; e->Le = cfg->n_enc_layers; e->Ld = cfg->n_dec_layers;
  e->V = cfg->n_vocab; e->ldv = (cfg->n_vocab + 7) / 8 * 8; e->Tctx = cfg->n_text_ctx; e->Tenc = cfg->n_audio_ctx;
  *out = e;
  B200W_API_END
}

int b200w_engine_destroy(b200w_engine* e) {
  B200W_API_BEGIN
  if (e) {
    DevBuf* bufs[] = {&e->e_x, &e->e_h, &e->e_qkv, &e->e_attn, &e->e_mlp, &e->d_x, &e->d_h, &e->d_qkv,
                      &e->d_attn, &e->d_cq, &e->d_mlp, &e->d_hsel, &e->kpool, &e->vpool, &e->tables[0],
                      &e->tables[1], &e->logits, &e->meta, &e->cross_off, &e->const_len, &e->mask,
                      &e->topk_ids, &e->topk_lp, &e->probs, &e->head_export, &e->al_w, &e->al_mat,
                      &e->al_trace, &e->al_path, &e->al_plen, &e->scratch_i, &e->d_split, &e->x_split, &e->x_tc};
    for (DevBuf* b : bufs) b->release();
    e->h_meta.release();
    e->h_out.release();
    cudaDeviceSynchronize();  // nothing of this engine may still be in flight
    for (auto& g : e->step_graphs) cudaGraphExecDestroy(g.exec);
    skinny_cache_destroy(e->skinny_cache);
    dstep_cache_destroy(e->dstep_cache);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->join_event) cudaEventDestroy(e->join_event);
    if (e->xev0) cudaEventDestroy(e->xev0);
    if (e->xev1) cudaEventDestroy(e->xev1);
    delete e;
  }
  B200W_API_END
}

int b200w_engine_set_weight(b200w_engine* e, const char* name, const void* dev_ptr, int64_t numel, int dtype) {
  B200W_API_BEGIN
  B200W_REQUIRE(e && name && dev_ptr, "null argument");
  B200W_REQUIRE(!e->finalized, "engine already finalized");
  e->weights[name] = Weight{dev_ptr, numel, dtype};
  B200W_API_END
}

int b200w_engine_finalize(b200w_engine* e) {
  B200W_API_BEGIN
  B200W_REQUIRE(e, "null engine");
  const int64_t d = e->d;
  e->conv1 = need_lin(e, "enc.conv1", d, 3 * 128);
  e->conv2 = need_lin(e, "enc.conv2", d, 3 * d);
  e->enc_pos = static_cast<const float*>(need(e, "enc.pos", (int64_t)e->Tenc * d, 0));
  e->enc.resize(e->Le);
  for (int l = 0; l < e->Le; ++l) {
    const std::string p = "enc.L" + std::to_string(l);
    EncLayer& L = e->enc[l];
    L.ln1 = need_norm(e, p + ".ln1", d);
    L.qkv = need_lin(e, p + ".qkv", 3 * d, d);
    L.out = need_lin(e, p + ".out", d, d);
    L.ln2 = need_norm(e, p + ".ln2", d);
    L.fc1 = need_lin(e, p + ".fc1", 4 * d, d);
    L.fc2 = need_lin(e, p + ".fc2", d, 4 * d);
  }
  e->enc_ln_post = need_norm(e, "enc.ln_post", d);
  e->tok_emb = static_cast<const bf16*>(need(e, "dec.tok_emb", (int64_t)e->V * d, 1));
  e->dec_pos = static_cast<const float*>(need(e, "dec.pos", (int64_t)e->Tctx * d, 0));
  e->dec.resize(e->Ld);
  for (int l = 0; l < e->Ld; ++l) {
    const std::string p = "dec.L" + std::to_string(l);
    DecLayer& L = e->dec[l];
    L.ln1 = need_norm(e, p + ".ln1", d);
    L.qkv = need_lin(e, p + ".qkv", 3 * d, d);
    L.out = need_lin(e, p + ".out", d, d);
    L.ln2 = need_norm(e, p + ".ln2", d);
    L.cq = need_lin(e, p + ".cq", d, d);
    L.cout = need_lin(e, p + ".cout", d, d);
    L.ln3 = need_norm(e, p + ".ln3", d);
    L.fc1 = need_lin(e, p + ".fc1", 4 * d, d);
    L.fc2 = need_lin(e, p + ".fc2", d, 4 * d);
  }
  e->cross_kv = need_lin(e, "dec.cross_kv", (int64_t)e->Ld * 2 * d, d);
  e->dec_ln = need_norm(e, "dec.ln", d);

  const int64_t Bm = e->cfg.max_batch, T = e->Tenc;
  e->e_x.alloc(Bm * T * d * 4);
  e->e_h.alloc(Bm * T * d * 2);
  e->e_qkv.alloc(Bm * T * 3 * d * 2);
  e->e_attn.alloc(Bm * T * d * 2);
  e->e_mlp.alloc(Bm * T * 4 * d * 2);  // also holds conv1's output [3000][d]
  e->n_max = (int)(Bm * e->Tctx);
  const int64_t n = e->n_max;
  e->d_x.alloc(n * d * 4);
  e->d_h.alloc(n * d * 2);
  e->d_qkv.alloc(n * 3 * d * 2);
  e->d_attn.alloc(n * d * 2);
  B200W_CHECK_CUDA(cudaMemset(e->d_attn.p, 0, n * d * 2));  // padding rows are never written: keep them finite
  e->d_cq.alloc(n * d * 2);
  e->d_mlp.alloc(n * 4 * d * 2);
  {
    // split-K slabs: the old path (at most 16 slabs of [256 rows][d]) and the decode-step path
    // (14 slabs of [rows][d] or 4 slabs of [rows][3d], rows = max_batch * max_beam)
    const size_t step_rows = (size_t)Bm * e->cfg.max_beam;
    const size_t need = std::max<size_t>((size_t)(128 * 160 * 64) * 16 * 4, 16 * step_rows * (size_t)d * 4);
    e->d_split.alloc(need);
  }
  e->x_split.alloc(cross_attention_scratch_bytes(64, e->H));
  e->x_tc.alloc(cross_attention_tc_scratch_bytes());
  B200W_CHECK_CUDA(cudaMemset(e->x_tc.p, 0, cross_attention_tc_scratch_bytes()));  // arrival counters start at zero
  e->slots_per_stream = e->Tctx * e->cfg.max_beam;
  e->total_slots = (long long)e->slots_per_stream * Bm;
  e->kpool.alloc((size_t)e->Ld * e->total_slots * d * 2);
  e->vpool.alloc((size_t)e->Ld * e->total_slots * d * 2);
  const int64_t rows = Bm * e->cfg.max_beam;
  e->tables[0].alloc(rows * e->Tctx * 4);
  e->tables[1].alloc(rows * e->Tctx * 4);
  e->logit_rows = (int)std::max<int64_t>(2 * rows, 1024);
  e->logits.alloc((size_t)e->logit_rows * e->ldv * 4);
  e->d_hsel.alloc((size_t)e->logit_rows * d * 2);
  e->meta.alloc((size_t)n * 12 * 4 + (size_t)n * 16 + 64);
  e->scratch_i.alloc(((size_t)e->logit_rows + 4 * Bm + n) * 4);
  e->cross_off.alloc(Bm * 8);
  e->const_len.alloc(n * 4);
  e->mask.alloc(e->V + 16);
  e->topk_ids.alloc(rows * MAX_TOPK * 4);
  e->topk_lp.alloc(rows * MAX_TOPK * 4);
  e->probs.alloc(std::max<int64_t>(n, Bm * 128) * 4);
  e->head_export.alloc((size_t)e->Ld * e->H * 4);
  {
    std::vector<long long> off(Bm);
    for (int64_t b = 0; b < Bm; ++b) off[b] = b * T * e->Ld * 2 * d;
    B200W_CHECK_CUDA(cudaMemcpy(e->cross_off.p, off.data(), Bm * 8, cudaMemcpyHostToDevice));
    std::vector<int> cl(n, (int)T);
    B200W_CHECK_CUDA(cudaMemcpy(e->const_len.p, cl.data(), n * 4, cudaMemcpyHostToDevice));
  }
  // pinned staging buffers get their final size here: captured decode graphs hold their addresses
  e->h_meta.alloc((size_t)n * 12 * 4 + (size_t)n * 16 + 64);
  e->h_out.alloc(std::max<size_t>({(size_t)rows * MAX_TOPK * 8 + 64, (size_t)Bm * (e->Tctx + T + 4) * 8 + Bm * 8,
                                   (size_t)Bm * 128 * 4}));
  B200W_CHECK_CUDA(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  B200W_CHECK_CUDA(cudaEventCreateWithFlags(&e->join_event, cudaEventDisableTiming));
  B200W_CHECK_CUDA(cudaEventCreate(&e->xev0));
  B200W_CHECK_CUDA(cudaEventCreate(&e->xev1));
  e->use_graphs = getenv("B200W_NO_GRAPHS") == nullptr;
  e->finalized = true;
  B200W_API_END
}

int b200w_encode(b200w_engine* e, const void* windows_dev, int B, void* enc_out_dev, void* stream) {
  B200W_API_BEGIN
  B200W_REQUIRE(e && e->finalized, "engine not finalized");
  std::lock_guard<std::mutex> engine_lock(e->mu);
  B200W_REQUIRE(B >= 1 && B <= e->cfg.max_batch, "batch out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int d = e->d, T = e->Tenc, M = B * T;
  float* x = e->e_x.as<float>();
  bf16* h = e->e_h.as<bf16>();
  bf16* qkv = e->e_qkv.as<bf16>();
  bf16* attn = e->e_attn.as<bf16>();
  bf16* mlp = e->e_mlp.as<bf16>();
  bf16* x1 = mlp;  // conv1 output [B][3000][d], dead before the first MLP
  conv1d_k3(static_cast<const bf16*>(windows_dev), e->conv1.w, e->conv1.b, nullptr, x1, B, 2 * T, 128, d, 1, 0, s);
  conv1d_k3(x1, e->conv2.w, e->conv2.b, e->enc_pos, x, B, 2 * T, d, d, 2, 1, s);
  for (int l = 0; l < e->Le; ++l) {
    const EncLayer& L = e->enc[l];
    layernorm(x, nullptr, L.ln1.g, L.ln1.b, h, M, d, s);
    linear(h, M, d, L.qkv, 3 * d, qkv, 3 * d, 0, 0, nullptr, 0, 0, s);
    enc_attention(qkv, B, T, d, e->H, attn, s);
    linear(attn, M, d, L.out, d, x, d, 1, 0, x, 1, d, s);
    layernorm(x, nullptr, L.ln2.g, L.ln2.b, h, M, d, s);
    linear(h, M, d, L.fc1, 4 * d, mlp, 4 * d, 0, 1, nullptr, 0, 0, s);
    linear(mlp, M, 4 * d, L.fc2, d, x, d, 1, 0, x, 1, d, s);
  }
  layernorm(x, nullptr, e->enc_ln_post.g, e->enc_ln_post.b, static_cast<bf16*>(enc_out_dev), M, d, s);
  B200W_API_END
}

int b200w_cross_kv(b200w_engine* e, const void* enc_out_dev, int B, void* cross_kv_dev, void* stream) {
  B200W_API_BEGIN
  B200W_REQUIRE(e && e->finalized, "engine not finalized");
  std::lock_guard<std::mutex> engine_lock(e->mu);
  B200W_REQUIRE(B >= 1 && B <= e->cfg.max_batch, "batch out of range");
  const int N = e->Ld * 2 * e->d;
  GemmDesc g;
  g.A = static_cast<const bf16*>(enc_out_dev);
  g.a_rows = B * e->Tenc;
  g.a_cols = e->d;
  g.a_row_stride = e->d;
  g.W = e->cross_kv.w;
  g.N = N;
  g.K = e->d;
  g.w_row_stride = e->d;
  g.out_rows = B * e->Tenc;
  g.bias = e->cross_kv.b;
  g.out = cross_kv_dev;
  g.out_row_stride = N;
  g.head_major_T = e->Tenc;  // -> [B][L][2][H][1500][64]: each head's keys are one contiguous block
  g.head_major_d = e->d;
  gemm_bf16(g, static_cast<cudaStream_t>(stream));
  B200W_API_END
}

// -------------------------------------------------------------------------------------- decode
namespace {

struct Hyp {
  std::vector<int> tokens;  // sampled tokens (no prompt)
  double cum = 0.0;
  int table_row = 0;
};
struct Finished {
  std::vector<int> tokens;  // without EOT
  double cum = 0.0;
  double score = 0.0;
};
struct StreamState {
  int P = 0;            // prompt length
  int max_new = 0;
  bool done = false;
  std::vector<Hyp> alive;
  std::vector<Finished> finished;
};

int4 sampling_state(const std::vector<int>& toks, const b200w_decode_opts& o, int sample_stream = 0) {
  int flags = 0, ts_lo = 0;
  if (!o.without_timestamps) flags |= SF_RULES;
  const int n = (int)toks.size();
  if (n == 0) flags |= SF_FIRST;
  if (!o.without_timestamps && n > 0) {
    const bool last_ts = toks[n - 1] >= o.timestamp_begin;
    const bool pen_ts = n < 2 || toks[n - 2] >= o.timestamp_begin;
    if (last_ts) flags |= pen_ts ? SF_SUP_ALL_TS : SF_SUP_TEXT;
    int last_stamp = -1;
    for (int i = n - 1; i >= 0; --i)
      if (toks[i] >= o.timestamp_begin) { last_stamp = toks[i]; break; }
    if (last_stamp >= 0) ts_lo = (last_ts && !pen_ts) ? last_stamp : last_stamp + 1;
  }
  return make_int4(flags, ts_lo, (int)toks.size(), sample_stream);  // .z / .w: the sampler's draw index and stream
}

double norm_score(double cum, int len, float lp) {
  return cum / pow((double)std::max(len, 1), (double)lp);
}

}  // namespace

int b200w_decode(b200w_engine* e, const void* cross_kv_dev, int B, const int32_t* prompts_host,
                 const int32_t* prompt_lens_host, const b200w_decode_opts* opts, int32_t* tokens_out_host,
                 int32_t max_out, int32_t* n_tokens_out_host, float* sum_logprob_out_host,
                 float* no_speech_out_host, void* stream) {
  B200W_API_BEGIN
  B200W_REQUIRE(e && e->finalized, "engine not finalized");
  std::lock_guard<std::mutex> engine_lock(e->mu);
  B200W_REQUIRE(B >= 1 && B <= e->cfg.max_batch, "batch out of range");
  B200W_REQUIRE(opts && opts->beam_size >= 1 && opts->beam_size <= e->cfg.max_beam, "beam size out of range");
  const bool sampling = opts->temperature > 0.f;
  if (sampling) B200W_REQUIRE(opts->best_of >= 1 && opts->best_of <= e->cfg.max_beam, "best_of out of range");
  // everything below runs on the engine's own stream, ordered after the caller's stream
  cudaStream_t s = e->own_stream;
  B200W_CHECK_CUDA(cudaEventRecord(e->join_event, static_cast<cudaStream_t>(stream)));
  B200W_CHECK_CUDA(cudaStreamWaitEvent(s, e->join_event, 0));
  const b200w_decode_opts& o = *opts;
  // sampling: the K rows of a chunk are best_of independent samples (one token drawn per row per step)
  const int K = sampling ? o.best_of : o.beam_size, k_top = (K == 1 || sampling) ? 1 : 2 * K;
  const bf16* ckv = static_cast<const bf16*>(cross_kv_dev);
  const int Tctx = e->Tctx;
  const int max_cand = sampling ? K : std::max(1, (int)lround(K * (double)o.patience));

  // static masks: bit0 suppress list, bit1 first-position blanks
  {
    std::vector<unsigned char> mk(e->V, 0);
    for (int i = 0; i < o.n_suppress; ++i)
      if (o.suppress_ids[i] >= 0 && o.suppress_ids[i] < e->V) mk[o.suppress_ids[i]] |= 1;
    if (o.suppress_blank)
      for (int i = 0; i < o.n_blank; ++i)
        if (o.blank_ids[i] >= 0 && o.blank_ids[i] < e->V) mk[o.blank_ids[i]] |= 2;
    B200W_CHECK_CUDA(cudaMemcpyAsync(e->mask.p, mk.data(), e->V, cudaMemcpyHostToDevice, s));
    B200W_CHECK_CUDA(cudaStreamSynchronize(s));  // mk goes out of scope
  }

  e->xev_recorded = false;
  const bool timing = getenv("B200W_DECODE_TIMING") != nullptr;
  e->optimer.on = e->optimer.forced || (getenv("B200W_DECODE_PROFILE") != nullptr && !e->use_graphs);
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(now() - t0).count();
  };
  const auto t_begin = now();
  double t_fill = 0, t_gpu = 0, t_book = 0;
  int n_steps = 0;
  // ---- prefill: every prompt token of every stream in one pass
  std::vector<StreamState> st(B);
  HostMeta hm;
  std::vector<int> sel_rows, sot_rows(B, -1);
  std::vector<int> table0((size_t)B * K * Tctx, 0);
  int off = 0, max_P = 0;
  for (int b = 0; b < B; ++b) {
    const int P = prompt_lens_host[b];
    B200W_REQUIRE(P >= 1 && P < Tctx, "prompt length out of range");
    st[b].P = P;
    st[b].max_new = Tctx - P;
    if (o.max_new_tokens > 0) st[b].max_new = std::min(st[b].max_new, (int)o.max_new_tokens);
    max_P = std::max(max_P, P);
    const int slot0 = b * e->slots_per_stream;
    for (int i = 0; i < P; ++i) {
      const int t = prompts_host[off + i];
      B200W_REQUIRE(t >= 0 && t < e->V, "prompt token out of range");
      if (t == o.sot) sot_rows[b] = hm.n();
      hm.push(t, i, slot0 + i, b * K, i + 1, b);
      for (int k = 0; k < K; ++k) table0[((size_t)b * K + k) * Tctx + i] = slot0 + i;
    }
    sel_rows.push_back(hm.n() - 1);  // logits at the last prompt position
    off += P;
  }
  for (int b = 0; b < B; ++b) sel_rows.push_back(sot_rows[b] >= 0 ? sot_rows[b] : sel_rows[b]);
  int cur = 0;
  B200W_CHECK_CUDA(cudaMemcpyAsync(e->tables[cur].p, table0.data(), table0.size() * 4, cudaMemcpyHostToDevice, s));
  PassMeta pm = upload_meta(e, hm, s);
  decoder_forward(e, ckv, pm, e->tables[cur].as<int>(), max_P, nullptr, s);
  // selected rows: [last position x B | sot position x B]
  int* sel_dev = e->scratch_i.as<int>();
  upload_ints(sel_dev, sel_rows, s);
  logits_for_rows(e, sel_dev, 2 * B, s);
  // no-speech probability from the raw logits at <|startoftranscript|>
  std::vector<int> ns_tok(B, o.no_speech), ns_rows(B);
  for (int b = 0; b < B; ++b) ns_rows[b] = B + b;
  int* ns_dev = e->scratch_i.as<int>() + 2 * B;
  upload_ints(ns_dev, ns_rows, s);
  upload_ints(ns_dev + B, ns_tok, s);
  token_probs(e->logits.as<float>(), e->ldv, e->V, ns_dev, ns_dev + B, nullptr, 0, e->probs.as<float>(), B, s);
  e->h_out.alloc(std::max<size_t>((size_t)B * K * MAX_TOPK * 8, (size_t)B * 4) + 64);
  B200W_CHECK_CUDA(cudaMemcpyAsync(e->h_out.p, e->probs.p, B * 4, cudaMemcpyDeviceToHost, s));
  B200W_CHECK_CUDA(cudaStreamSynchronize(s));
  for (int b = 0; b < B; ++b) no_speech_out_host[b] = e->h_out.as<float>()[b];

  // ---- first sampled position: one logits row per stream (rows 0..B of e->logits)
  // (sampling: K independent draws from the same row, the k-th launch reads the states / writes the outputs of
  // entries [k B, (k + 1) B); the sampler's stream id of (chunk b, sample k) is b K + k, as in the steps)
  const int n_first = sampling ? K : 1;
  std::vector<int4> sstate((size_t)B * n_first);
  for (int k = 0; k < n_first; ++k)
    for (int b = 0; b < B; ++b) sstate[(size_t)k * B + b] = sampling_state({}, o, b * K + k);
  int4* state_dev = reinterpret_cast<int4*>(e->meta.p);  // reuse: the pass metadata is consumed
  B200W_CHECK_CUDA(cudaMemcpyAsync(state_dev, sstate.data(), sstate.size() * 16, cudaMemcpyHostToDevice, s));
  SampleParams sp;
  sp.logits = e->logits.as<float>(); sp.ld = e->ldv; sp.n_vocab = e->V;
  sp.mask = e->mask.as<unsigned char>(); sp.state = state_dev;
  sp.eot = o.eot; sp.no_timestamps = o.no_timestamps; sp.timestamp_begin = o.timestamp_begin;
  sp.max_initial_ts = o.max_initial_timestamp_index; sp.k = k_top;
  sp.out_ids = e->topk_ids.as<int>(); sp.out_logprobs = e->topk_lp.as<float>();
  sp.temperature = sampling ? o.temperature : 0.f;
  sp.seed = o.seed;
  for (int k = 0; k < n_first; ++k) {
    SampleParams spk = sp;
    spk.state = state_dev + (size_t)k * B;
    spk.out_ids = sp.out_ids + (size_t)k * B * k_top;
    spk.out_logprobs = sp.out_logprobs + (size_t)k * B * k_top;
    sample_topk(spk, B, s);
  }
  int* h_ids = e->h_out.as<int>();
  float* h_lp = reinterpret_cast<float*>(h_ids + (size_t)B * K * MAX_TOPK);
  B200W_CHECK_CUDA(cudaMemcpyAsync(h_ids, sp.out_ids, (size_t)B * n_first * k_top * 4, cudaMemcpyDeviceToHost, s));
  B200W_CHECK_CUDA(cudaMemcpyAsync(h_lp, sp.out_logprobs, (size_t)B * n_first * k_top * 4, cudaMemcpyDeviceToHost, s));
  B200W_CHECK_CUDA(cudaStreamSynchronize(s));

  auto finish_stream = [&](StreamState& S) {
    // out of budget: every live hypothesis ends here (no EOT was produced)
    for (Hyp& h : S.alive) {
      Finished f;
      f.tokens = h.tokens;
      f.cum = h.cum;
      f.score = norm_score(h.cum, (int)h.tokens.size(), o.length_penalty);
      S.finished.push_back(std::move(f));
    }
    S.alive.clear();
    S.done = true;
  };

  for (int b = 0; b < B; ++b) {
    StreamState& S = st[b];
    const int n_cand = sampling ? K : k_top;  // sampling: one draw per sample; search: the k_top best of the row
    for (int c = 0; c < n_cand && (int)S.alive.size() < K; ++c) {
      const int tok = sampling ? h_ids[c * B + b] : h_ids[b * k_top + c];
      const float lp = sampling ? h_lp[c * B + b] : h_lp[b * k_top + c];
      if (tok == 0x7fffffff || !(lp > -INFINITY)) continue;
      if (tok == o.eot) {
        if ((int)S.finished.size() < max_cand) {
          Finished f;
          f.cum = lp;
          f.score = norm_score(lp, 0, o.length_penalty);
          S.finished.push_back(f);
        }
        continue;
      }
      Hyp h;
      h.tokens.push_back(tok);
      h.cum = lp;
      h.table_row = b * K + (int)S.alive.size();
      S.alive.push_back(std::move(h));
    }
    if (S.alive.empty() || (int)S.finished.size() >= max_cand) S.done = true;
    if (!S.done && S.max_new <= 1) finish_stream(S);
  }

  // ---- autoregressive steps.  Fixed shape: row r = b*K + k for every (stream, beam slot), whether
  // it is live or not (idle rows decode a dummy token into a scratch slot), so that every step
  // issues the same launches on the same buffers and can be replayed as a CUDA graph.
  const int opts_hash = o.eot * 31 + o.timestamp_begin * 17 + o.max_initial_timestamp_index * 7 + k_top +
                        (o.without_timestamps ? 1000003 : 0) +
                        (sampling ? (int)(__builtin_bit_cast(unsigned, o.temperature) * 2654435761u + o.seed * 40503u) : 0);
  B200W_CHECK_CUDA(cudaMemsetAsync(e->tables[cur ^ 1].p, 0, (size_t)e->cfg.max_batch * e->cfg.max_beam * Tctx * 4, s));
  auto enqueue_step = [&](const PassMeta& m, size_t meta_bytes, int from) {
    B200W_CHECK_CUDA(cudaMemcpyAsync(e->meta.p, e->h_meta.p, meta_bytes, cudaMemcpyHostToDevice, s));
    // fork the page tables: new[row] = old[parent][0:pos] + [slot]
    gather_fork_tables(e->tables[from].as<int>(), m.parent, m.trow, m.pos, m.slot, e->tables[from ^ 1].as<int>(),
                       m.n, Tctx, s);
    e->step_pass = true;
    try {
      if (fused_step_ok(e, m.n) && m.max_group <= CROSS_GROUP && m.n_groups <= 64) {
        decoder_step_fused(e, ckv, m, e->tables[from ^ 1].as<int>(), Tctx, s);
      } else {
        decoder_forward(e, ckv, m, e->tables[from ^ 1].as<int>(), Tctx, nullptr, s);
        logits_for_rows(e, nullptr, m.n, s);
      }
    } catch (...) {
      e->step_pass = false;
      throw;
    }
    e->step_pass = false;
    SampleParams sps = sp;
    sps.state = m.state;
    RUN(12) { OP(12); sample_topk(sps, m.n, s); }
    B200W_CHECK_CUDA(cudaMemcpyAsync(h_ids, sp.out_ids, (size_t)m.n * k_top * 4, cudaMemcpyDeviceToHost, s));
    B200W_CHECK_CUDA(cudaMemcpyAsync(h_lp, sp.out_logprobs, (size_t)m.n * k_top * 4, cudaMemcpyDeviceToHost, s));
  };
  const double t_prefill = ms_since(t_begin);
  for (int step = 1;; ++step) {
    const auto t_s0 = now();
    hm.clear();
    bool any = false;
    for (int b = 0; b < B; ++b) {
      StreamState& S = st[b];
      for (int k = 0; k < K; ++k) {
        const int row = b * K + k;
        if (!S.done && k < (int)S.alive.size()) {
          Hyp& h = S.alive[k];
          const int pos = S.P + step - 1;
          const int slot = b * e->slots_per_stream + S.P + (step - 1) * K + k;
          hm.push(h.tokens.back(), pos, slot, row, pos + 1, b, sampling_state(h.tokens, o, row), h.table_row);
          h.table_row = row;
          any = true;
        } else {  // idle row: one dummy token at position 0 in a scratch slot of its own
          hm.push(o.eot, 0, b * e->slots_per_stream + e->slots_per_stream - 1 - k, row, 1, b,
                  make_int4(0, 0, 0, 0), row, true);
        }
      }
    }
    if (!any) break;
    size_t meta_bytes = 0;
    pm = fill_host_meta(e, hm, &meta_bytes);
    t_fill += ms_since(t_s0);
    const auto t_s1 = now();
    ++n_steps;
    cudaGraphExec_t exec = nullptr;
    if (e->h_meta.grew || e->h_out.grew) {  // staging buffers moved: captured addresses are stale
      for (auto& g : e->step_graphs) cudaGraphExecDestroy(g.exec);
      e->step_graphs.clear();
      e->h_meta.grew = e->h_out.grew = false;
    }
    if (e->use_graphs && step >= 2 && !e->optimer.on) {
      for (auto& g : e->step_graphs)
        if (g.ckv == ckv && g.B == B && g.K == K && g.parity == cur && g.opts_hash == opts_hash) exec = g.exec;
      if (!exec) {
        cudaGraph_t graph = nullptr;
        B200W_CHECK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        try {
          enqueue_step(pm, meta_bytes, cur);
        } catch (...) {
          cudaStreamEndCapture(s, &graph);
          if (graph) cudaGraphDestroy(graph);
          throw;
        }
        B200W_CHECK_CUDA(cudaStreamEndCapture(s, &graph));
        B200W_CHECK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        B200W_CHECK_CUDA(cudaGraphDestroy(graph));
        if (e->step_graphs.size() >= 16) {
          for (auto& g : e->step_graphs) cudaGraphExecDestroy(g.exec);
          e->step_graphs.clear();
        }
        e->step_graphs.push_back({ckv, B, K, cur, opts_hash, exec});
      }
      B200W_CHECK_CUDA(cudaGraphLaunch(exec, s));
      g_launch_count.fetch_add(e->launches_per_step, std::memory_order_relaxed);
    } else {
      const unsigned long long before = g_launch_count.load();
      enqueue_step(pm, meta_bytes, cur);
      e->launches_per_step = g_launch_count.load() - before;
    }
    cur ^= 1;
    B200W_CHECK_CUDA(cudaStreamSynchronize(s));
    t_gpu += ms_since(t_s1);
    if (e->xev_recorded) {  // the probe events were (re)recorded by this step: graph replay or eager launch
      float pms = 0.f;
      if (cudaEventElapsedTime(&pms, e->xev0, e->xev1) == cudaSuccess) {
        e->probe_ms += pms;
        e->probe_n += 1;
        e->probe_bytes += (double)pm.n_live_host * e->H * 2.0 * e->Tenc * HEAD_DIM * 2.0;
      } else {
        cudaGetLastError();
      }
    }
    if (ablate_mask() != 0)  // timing-only runs: keep every beam alive whatever the (garbage) logits say
      for (int r = 0; r < B * K; ++r)
        for (int c = 0; c < k_top; ++c) { h_ids[r * k_top + c] = 1000 + c; h_lp[r * k_top + c] = -1.f - c; }
    const auto t_s2 = now();

    // host-side beam bookkeeping
    for (int b = 0; b < B; ++b) {
      StreamState& S = st[b];
      if (S.done) continue;
      const int na = (int)S.alive.size();
      const int r0 = b * K;
      if (sampling) {  // every sample continues with its own draw
        std::vector<Hyp> next;
        for (int k = 0; k < na; ++k) {
          const int tok = h_ids[r0 + k];
          const float lp = h_lp[r0 + k];
          Hyp& par = S.alive[k];
          if (tok == 0x7fffffff || !(lp > -INFINITY) || tok == o.eot) {  // (no admissible token: the sample ends)
            Finished f;
            f.tokens = par.tokens;
            f.cum = par.cum + (lp > -INFINITY ? (double)lp : 0.0);
            f.score = norm_score(f.cum, (int)par.tokens.size(), o.length_penalty);
            S.finished.push_back(std::move(f));
            continue;
          }
          par.tokens.push_back(tok);
          par.cum += (double)lp;
          next.push_back(std::move(par));
        }
        S.alive.swap(next);
        if (S.alive.empty()) { S.done = true; continue; }
        if (step + 1 >= S.max_new) finish_stream(S);
        continue;
      }
      struct Cand { double score; int k, tok; float lp; };
      std::vector<Cand> cands;
      cands.reserve((size_t)na * k_top);
      for (int k = 0; k < na; ++k)
        for (int c = 0; c < k_top; ++c) {
          const int tok = h_ids[(r0 + k) * k_top + c];
          const float lp = h_lp[(r0 + k) * k_top + c];
          if (tok == 0x7fffffff || !(lp > -INFINITY)) continue;
          cands.push_back({S.alive[k].cum + (double)lp, k, tok, lp});
        }
      std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& c) {
        if (a.score != c.score) return a.score > c.score;
        if (a.k != c.k) return a.k < c.k;
        return a.tok < c.tok;
      });
      std::vector<Hyp> next;
      const int limit = std::min((int)cands.size(), k_top);
      for (int c = 0; c < limit && (int)next.size() < K; ++c) {
        const Cand& cd = cands[c];
        const Hyp& par = S.alive[cd.k];
        if (cd.tok == o.eot) {
          if ((int)S.finished.size() < max_cand) {
            Finished f;
            f.tokens = par.tokens;
            f.cum = cd.score;
            f.score = norm_score(cd.score, (int)par.tokens.size(), o.length_penalty);
            S.finished.push_back(std::move(f));
          }
          continue;
        }
        Hyp h;
        h.tokens = par.tokens;
        h.tokens.push_back(cd.tok);
        h.cum = cd.score;
        h.table_row = par.table_row;
        next.push_back(std::move(h));
      }
      S.alive.swap(next);
      if (S.alive.empty() || (int)S.finished.size() >= max_cand) { S.done = true; continue; }
      if (step + 1 >= S.max_new) finish_stream(S);
    }
    t_book += ms_since(t_s2);
  }
  {  // the single-pass cross attention raises a flag instead of returning numbers it cannot represent
    int flag = 0;
    char* fp = e->x_tc.as<char>() + cross_attention_tc_error_offset();
    B200W_CHECK_CUDA(cudaMemcpyAsync(&flag, fp, 4, cudaMemcpyDeviceToHost, s));
    B200W_CHECK_CUDA(cudaStreamSynchronize(s));
    if (flag) {
      B200W_CHECK_CUDA(cudaMemsetAsync(fp, 0, 4, s));
      throw Error("cross attention: a score exceeds the first key tile's maximum by more than 2^100");
    }
  }
  e->optimer.collect("decode (prefill + steps)");
  if (timing)
    fprintf(stderr, "[b200w_decode] B=%d K=%d steps=%d prefill+first=%.2f ms, per step: fill %.3f ms, launch+gpu %.3f ms, "
            "bookkeeping %.3f ms, total %.1f ms\n", B, K, n_steps, t_prefill, t_fill / std::max(n_steps, 1),
            t_gpu / std::max(n_steps, 1), t_book / std::max(n_steps, 1), ms_since(t_begin));

  for (int b = 0; b < B; ++b) {
    StreamState& S = st[b];
    const Finished* best = nullptr;
    for (const Finished& f : S.finished)
      if (!best || f.score > best->score) best = &f;
    int n = 0;
    float cum = 0.f;
    if (best) {
      n = std::min((int)best->tokens.size(), (int)max_out);
      for (int i = 0; i < n; ++i) tokens_out_host[(size_t)b * max_out + i] = best->tokens[i];
      cum = (float)best->cum;
    }
    n_tokens_out_host[b] = n;
    sum_logprob_out_host[b] = cum;
  }
  B200W_API_END
}

int b200w_decoder_logits(b200w_engine* e, const void* cross_kv_dev, int B, const int32_t* tokens_host,
                         const int32_t* token_lens_host, float* logits_out_dev, void* stream) {
  B200W_API_BEGIN
  B200W_REQUIRE(e && e->finalized, "engine not finalized");
  std::lock_guard<std::mutex> engine_lock(e->mu);
  B200W_REQUIRE(B >= 1 && B <= e->cfg.max_batch, "batch out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HostMeta hm;
  std::vector<int> table((size_t)B * e->Tctx, 0);
  int off = 0, maxlen = 0;
  for (int b = 0; b < B; ++b) {
    const int n = token_lens_host[b];
    B200W_REQUIRE(n >= 1 && n <= e->Tctx, "sequence length out of range");
    maxlen = std::max(maxlen, n);
    for (int i = 0; i < n; ++i) {
      const int slot = b * e->slots_per_stream + i;
      hm.push(tokens_host[off + i], i, slot, b, i + 1, b);
      table[(size_t)b * e->Tctx + i] = slot;
    }
    off += n;
  }
  B200W_CHECK_CUDA(cudaMemcpyAsync(e->tables[0].p, table.data(), table.size() * 4, cudaMemcpyHostToDevice, s));
  PassMeta pm = upload_meta(e, hm, s);
  decoder_forward(e, static_cast<const bf16*>(cross_kv_dev), pm, e->tables[0].as<int>(), maxlen, nullptr, s);
  B200W_CHECK_CUDA(cudaStreamSynchronize(s));  // host vectors above
  const int total = hm.n();
  for (int r0 = 0; r0 < total; r0 += e->logit_rows) {
    const int nr = std::min(e->logit_rows, total - r0);
    std::vector<int> rows(nr);
    for (int i = 0; i < nr; ++i) rows[i] = r0 + i;
    upload_ints(e->scratch_i.as<int>(), rows, s);
    logits_for_rows(e, e->scratch_i.as<int>(), nr, s);
    B200W_CHECK_CUDA(cudaMemcpyAsync(logits_out_dev + (size_t)r0 * e->ldv, e->logits.p,
                                     (size_t)nr * e->ldv * 4, cudaMemcpyDeviceToDevice, s));
  }
  B200W_CHECK_CUDA(cudaStreamSynchronize(s));
  B200W_API_END
}

int b200w_detect_language(b200w_engine* e, const void* cross_kv_dev, int B, int32_t sot,
                          const int32_t* lang_ids_host, int32_t n_lang, float* probs_out_host, void* stream) {
  B200W_API_BEGIN
  B200W_REQUIRE(e && e->finalized, "engine not finalized");
  std::lock_guard<std::mutex> engine_lock(e->mu);
  B200W_REQUIRE(B >= 1 && B <= e->cfg.max_batch && n_lang >= 1 && n_lang <= 128, "arguments out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HostMeta hm;
  std::vector<int> table((size_t)B * e->Tctx, 0);
  for (int b = 0; b < B; ++b) {
    const int slot = b * e->slots_per_stream;
    hm.push(sot, 0, slot, b, 1, b);
    table[(size_t)b * e->Tctx] = slot;
  }
  B200W_CHECK_CUDA(cudaMemcpyAsync(e->tables[0].p, table.data(), table.size() * 4, cudaMemcpyHostToDevice, s));
  PassMeta pm = upload_meta(e, hm, s);
  decoder_forward(e, static_cast<const bf16*>(cross_kv_dev), pm, e->tables[0].as<int>(), 1, nullptr, s);
  logits_for_rows(e, nullptr, B, s);
  std::vector<int> ids(lang_ids_host, lang_ids_host + n_lang);
  upload_ints(e->scratch_i.as<int>(), ids, s);
  token_probs(e->logits.as<float>(), e->ldv, e->V, nullptr, nullptr, e->scratch_i.as<int>(), n_lang,
              e->probs.as<float>(), B, s);
  e->h_out.alloc((size_t)B * n_lang * 4);
  B200W_CHECK_CUDA(cudaMemcpyAsync(e->h_out.p, e->probs.p, (size_t)B * n_lang * 4, cudaMemcpyDeviceToHost, s));
  B200W_CHECK_CUDA(cudaStreamSynchronize(s));
  memcpy(probs_out_host, e->h_out.p, (size_t)B * n_lang * 4);
  B200W_API_END
}

int b200w_align(b200w_engine* e, const void* cross_kv_dev, int B, const int32_t* tokens_host,
                const int32_t* token_lens_host, const int32_t* num_frames_host, int32_t sot_len,
                const int32_t* align_heads_host, int32_t n_align, int32_t median_width,
                int32_t* path_out_host, int32_t max_path, int32_t* path_len_out_host,
                float* tok_prob_out_host, int32_t max_tok, void* stream) {
  B200W_API_BEGIN
  B200W_REQUIRE(e && e->finalized, "engine not finalized");
  std::lock_guard<std::mutex> engine_lock(e->mu);
  B200W_REQUIRE(B >= 1 && B <= e->cfg.max_batch && n_align >= 1, "arguments out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Tenc = e->Tenc;
  HostMeta hm;
  std::vector<int> table((size_t)B * e->Tctx, 0);
  std::vector<int> q0(B), ntok(B), nf(B);
  int off = 0, maxlen = 0, max_nf = 1;
  for (int b = 0; b < B; ++b) {
    const int n = token_lens_host[b];
    B200W_REQUIRE(n >= sot_len + 2 && n <= e->Tctx, "align sequence length out of range");
    q0[b] = hm.n();
    ntok[b] = n;
    nf[b] = std::max(1, std::min(Tenc, num_frames_host[b] / 2));
    max_nf = std::max(max_nf, nf[b]);
    maxlen = std::max(maxlen, n);
    for (int i = 0; i < n; ++i) {
      const int slot = b * e->slots_per_stream + i;
      hm.push(tokens_host[off + i], i, slot, b, i + 1, b);
      table[(size_t)b * e->Tctx + i] = slot;
    }
    off += n;
  }
  for (int i = 0; i < n_align; ++i)
    B200W_REQUIRE(align_heads_host[2 * i] >= 0 && align_heads_host[2 * i] < e->Ld &&
                      align_heads_host[2 * i + 1] >= 0 && align_heads_host[2 * i + 1] < e->H,
                  "alignment head out of range");
  const int total = hm.n();
  B200W_REQUIRE(max_path >= maxlen + max_nf, "max_path too small");
  e->al_w.alloc((size_t)n_align * total * Tenc * 4);
  e->al_mat.alloc((size_t)total * Tenc * 4);
  const long long trace_stride = (long long)(maxlen + 1) * (max_nf + 1);
  e->al_trace.alloc((size_t)B * trace_stride);
  e->al_path.alloc((size_t)B * max_path * 8);
  e->al_plen.alloc((size_t)B * 4 + 12 * (size_t)B);
  B200W_CHECK_CUDA(cudaMemcpyAsync(e->tables[0].p, table.data(), table.size() * 4, cudaMemcpyHostToDevice, s));
  PassMeta pm = upload_meta(e, hm, s);
  AlignExport ax;
  ax.probs = e->al_w.as<float>();
  ax.layer_has.assign(e->Ld, 0);
  {
    std::vector<int> ht((size_t)e->Ld * e->H, -1);
    for (int i = 0; i < n_align; ++i) {
      ht[(size_t)align_heads_host[2 * i] * e->H + align_heads_host[2 * i + 1]] = i;
      ax.layer_has[align_heads_host[2 * i]] = 1;
    }
    upload_ints(e->head_export.as<int>(), ht, s);
  }
  ax.head_table = e->head_export.as<int>();
  decoder_forward(e, static_cast<const bf16*>(cross_kv_dev), pm, e->tables[0].as<int>(), maxlen, &ax, s);

  // seq metadata after path_len: [q0 | ntok | nf]
  int* plen = e->al_plen.as<int>();
  int* seq = plen + B;
  upload_ints(seq, q0, s);
  upload_ints(seq + B, ntok, s);
  upload_ints(seq + 2 * B, nf, s);
  align_postprocess(e->al_w.as<float>(), n_align, total, Tenc, seq, seq + B, seq + 2 * B, B, maxlen, max_nf,
                    median_width, e->al_mat.as<float>(), sot_len, 1, e->al_trace.as<signed char>(),
                    trace_stride, e->al_path.as<int>(), max_path, plen, s);

  // teacher-forced probability of every text token: row (sot_len + i) predicts token sot_len + 1 + i
  std::vector<int> prow, ptok, pcount(B);
  for (int b = 0; b < B; ++b) {
    const int n_text = ntok[b] - sot_len - 2;  // minus no_timestamps and eot
    pcount[b] = std::max(0, std::min(n_text, (int)max_tok));
    for (int i = 0; i < pcount[b]; ++i) {
      prow.push_back(q0[b] + sot_len + i);
      ptok.push_back(hm.tok[q0[b] + sot_len + 1 + i]);
    }
  }
  std::vector<float> probs_host(prow.size());
  for (size_t r0 = 0; r0 < prow.size(); r0 += e->logit_rows) {
    const int nr = (int)std::min<size_t>(e->logit_rows, prow.size() - r0);
    std::vector<int> rr(prow.begin() + r0, prow.begin() + r0 + nr), tt(ptok.begin() + r0, ptok.begin() + r0 + nr);
    int* scratch = e->scratch_i.as<int>();
    upload_ints(scratch, rr, s);
    logits_for_rows(e, scratch, nr, s);
    upload_ints(scratch, tt, s);
    token_probs(e->logits.as<float>(), e->ldv, e->V, nullptr, scratch, nullptr, 0, e->probs.as<float>(), nr, s);
    B200W_CHECK_CUDA(cudaMemcpyAsync(probs_host.data() + r0, e->probs.p, (size_t)nr * 4, cudaMemcpyDeviceToHost, s));
    B200W_CHECK_CUDA(cudaStreamSynchronize(s));
  }
  e->h_out.alloc((size_t)B * max_path * 8 + (size_t)B * 4);
  B200W_CHECK_CUDA(cudaMemcpyAsync(e->h_out.p, e->al_path.p, (size_t)B * max_path * 8, cudaMemcpyDeviceToHost, s));
  B200W_CHECK_CUDA(cudaMemcpyAsync(e->h_out.as<char>() + (size_t)B * max_path * 8, plen, (size_t)B * 4,
                                   cudaMemcpyDeviceToHost, s));
  B200W_CHECK_CUDA(cudaStreamSynchronize(s));
  memcpy(path_out_host, e->h_out.p, (size_t)B * max_path * 8);
  memcpy(path_len_out_host, e->h_out.as<char>() + (size_t)B * max_path * 8, (size_t)B * 4);
  size_t k = 0;
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < (int)max_tok; ++i) tok_prob_out_host[(size_t)b * max_tok + i] = 0.f;
    for (int i = 0; i < pcount[b]; ++i) tok_prob_out_host[(size_t)b * max_tok + i] = probs_host[k++];
  }
  B200W_API_END
}

}  // extern "C"

extern "C" int b200w_op_profile_enable(b200w_engine* e, int on) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lock(e->mu);
  OpTimer& t = e->optimer;
  t.forced = on != 0;
  if (on) {
    for (int i = 0; i < 16; ++i) { t.tot_ms[i] = 0; t.tot_work[i] = 0; t.tot_n[i] = 0; }
  }
  return 0;
}
extern "C" int b200w_op_profile_read(b200w_engine* e, int op, const char** name, double* total_ms, int64_t* launches,
                                     double* work) {
  if (!e || op < 0 || op >= 16) return -1;
  std::lock_guard<std::mutex> lock(e->mu);
  const OpTimer& t = e->optimer;
  if (name) *name = t.names[op];
  if (total_ms) *total_ms = t.tot_ms[op];
  if (launches) *launches = t.tot_n[op];
  if (work) *work = t.tot_work[op];
  return 0;
}

extern "C" int b200w_step_probe_reset(b200w_engine* e) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lock(e->mu);
  e->probe_ms = 0;
  e->probe_bytes = 0;
  e->probe_n = 0;
  return 0;
}
extern "C" int b200w_step_probe_read(b200w_engine* e, double* total_ms, int64_t* launches, double* bytes) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lock(e->mu);
  if (total_ms) *total_ms = e->probe_ms;
  if (launches) *launches = e->probe_n;
  if (bytes) *bytes = e->probe_bytes;
  return 0;
}

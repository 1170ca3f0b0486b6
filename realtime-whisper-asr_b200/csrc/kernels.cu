// Small memory-bound kernels around the GEMMs: LayerNorm, embeddings, KV-cache scatter, Whisper
// logit rules + log-softmax + top-k, token probabilities, and the word-timing chain
// (normalise -> median filter -> head mean -> DTW).
#include "kernels.h"

#include <float.h>

namespace b200w {

// exact-erf GELU (Abramowitz-Stegun 7.1.26), same arithmetic as the GEMM epilogue (ptx.cuh)
__device__ __forceinline__ float gelu_erf_k(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;  // 1 / (1 + p z): MUFU reciprocal (1 ulp) instead of the IEEE-rounded sequence
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// ------------------------------------------------------------------------------- LayerNorm
// one warp per row; x fp32 (optionally gathered through row_idx) -> y bf16.  d % 128 == 0, d <= 1280.
// Optional fused split-K reduce: x[row] += bias + sum_s partial[s][row] (slabs added in order) before
// the normalisation; the sum is written back to x unless rows are gathered.
__global__ void __launch_bounds__(128) layernorm_kernel(float* __restrict__ x, const int* __restrict__ row_idx,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ y,
                                                        int rows, int d, const float* __restrict__ partial,
                                                        int n_partial, long long partial_stride,
                                                        const float* __restrict__ pbias) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long long src = row_idx ? row_idx[row] : row;
  float4* xp = reinterpret_cast<float4*>(x + src * d);
  const int nv = d >> 7;  // float4 per lane
  float4 v[10];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 10; ++i)
    if (i < nv) {
      v[i] = xp[lane + 32 * i];
      if (n_partial > 0) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(pbias) + lane + 32 * i);
        v[i].x += bb.x; v[i].y += bb.y; v[i].z += bb.z; v[i].w += bb.w;
        for (int s = 0; s < n_partial; ++s) {
          const float4 pp = reinterpret_cast<const float4*>(partial + s * partial_stride + src * d)[lane + 32 * i];
          v[i].x += pp.x; v[i].y += pp.y; v[i].z += pp.z; v[i].w += pp.w;
        }
        if (!row_idx) xp[lane + 32 * i] = v[i];
      }
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 10; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      sq += a * a + b * b + c * c + e * e;
    }
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)d + 1e-5f);
  const float4* gp = reinterpret_cast<const float4*>(gamma);
  const float4* bp = reinterpret_cast<const float4*>(beta);
  uint2* yp = reinterpret_cast<uint2*>(y + (long long)row * d);
#pragma unroll
  for (int i = 0; i < 10; ++i)
    if (i < nv) {
      const float4 g = __ldg(gp + lane + 32 * i), b = __ldg(bp + lane + 32 * i);
      __nv_bfloat162 lo = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + b.x,
                                                (v[i].y - mean) * rstd * g.y + b.y);
      __nv_bfloat162 hi = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + b.z,
                                                (v[i].w - mean) * rstd * g.w + b.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      yp[lane + 32 * i] = u;
    }
}

// Few rows (decode steps) with pending split-K partial sums: one CTA per row, one float4 per thread, so
// the 1 + 1 + n_partial loads of a thread are all independent and in flight together.
__global__ void __launch_bounds__(320) layernorm_reduce_kernel(float* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, bf16* __restrict__ y,
                                                               int d, const float* __restrict__ partial, int n_partial,
                                                               long long partial_stride,
                                                               const float* __restrict__ pbias) {
  __shared__ float s_red[2][10];
  pdl_trigger();
  const int row = blockIdx.x, t = threadIdx.x;
  const float4 bb = __ldg(reinterpret_cast<const float4*>(pbias) + t);  // weights: safe ahead of the wait
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + t), be = __ldg(reinterpret_cast<const float4*>(beta) + t);
  pdl_wait();
  float4* xp = reinterpret_cast<float4*>(x + (long long)row * d);
  float4 v = xp[t];
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* pp = reinterpret_cast<const float4*>(partial + (long long)row * d) + t;
  const long long st4 = partial_stride / 4;
  int sidx = 0;
  for (; sidx + 4 <= n_partial; sidx += 4) {  // slabs are summed in a fixed order -> deterministic
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = pp[(sidx + u) * st4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc[u].x += q[u].x; acc[u].y += q[u].y; acc[u].z += q[u].z; acc[u].w += q[u].w; }
  }
  for (; sidx < n_partial; ++sidx) {
    const float4 q = pp[sidx * st4];
    acc[0].x += q.x; acc[0].y += q.y; acc[0].z += q.z; acc[0].w += q.w;
  }
  v.x += bb.x + ((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x));
  v.y += bb.y + ((acc[0].y + acc[1].y) + (acc[2].y + acc[3].y));
  v.z += bb.z + ((acc[0].z + acc[1].z) + (acc[2].z + acc[3].z));
  v.w += bb.w + ((acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
  xp[t] = v;
  const int warp = t >> 5, lane = t & 31, nw = blockDim.x >> 5;
  float sum = v.x + v.y + v.z + v.w;
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_red[0][warp] = sum;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < nw; ++w) tot += s_red[0][w];
  const float mean = tot / (float)d;
  const float a = v.x - mean, b = v.y - mean, c = v.z - mean, e = v.w - mean;
  float sq = a * a + b * b + c * c + e * e;
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) s_red[1][warp] = sq;
  __syncthreads();
  float tsq = 0.f;
  for (int w = 0; w < nw; ++w) tsq += s_red[1][w];
  const float rstd = rsqrtf(tsq / (float)d + 1e-5f);
  __nv_bfloat162 lo = __floats2bfloat162_rn(a * rstd * g.x + be.x, b * rstd * g.y + be.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(c * rstd * g.z + be.z, e * rstd * g.w + be.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  reinterpret_cast<uint2*>(y + (long long)row * d)[t] = u;
}

void layernorm(float* x, const int* row_idx, const float* gamma, const float* beta, bf16* y, int rows, int d,
               cudaStream_t stream, const SplitReduce* sr) {
  if (rows <= 0) return;
  B200W_REQUIRE(d % 128 == 0 && d <= 1280, "LayerNorm width must be a multiple of 128, <= 1280");
  const bool red = sr && sr->n > 0;
  if (red && !row_idx) {
    launch_pdl(layernorm_reduce_kernel, dim3(rows), dim3(d / 4), 0, stream, x, gamma, beta, y, d, sr->partial, sr->n,
               sr->stride, sr->bias);
    B200W_COUNT_LAUNCH();
    B200W_CHECK_CUDA(cudaGetLastError());
    return;
  }
  launch_pdl(layernorm_kernel, dim3(cdiv(rows, 4)), dim3(128), 0, stream, x, row_idx, gamma, beta, y, rows, d,
             red ? sr->partial : (const float*)nullptr, red ? sr->n : 0, red ? sr->stride : 0LL,
             red ? sr->bias : (const float*)nullptr);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------- embeddings
__global__ void embed_kernel(const int* __restrict__ tok, const int* __restrict__ pos,
                             const bf16* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                             float* __restrict__ x, int d) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x;
  const bf16* te = tok_emb + (long long)tok[n] * d;
  const float* pe = pos_emb + (long long)pos[n] * d;
  for (int c = threadIdx.x * 2; c < d; c += blockDim.x * 2) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(te + c));
    const float2 p = *reinterpret_cast<const float2*>(pe + c);
    *reinterpret_cast<float2*>(x + (long long)n * d + c) = make_float2(t.x + p.x, t.y + p.y);
  }
}

void embed_tokens(const int* tok, const int* pos, const bf16* tok_emb, const float* pos_emb, float* x,
                  int n, int d, cudaStream_t stream) {
  if (n <= 0) return;
  launch_pdl(embed_kernel, dim3(n), dim3(128), 0, stream, tok, pos, tok_emb, pos_emb, x, d);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------- KV scatter
// qkv bf16 [n][3d] -> kpool[slot[n]][d], vpool[slot[n]][d]
__global__ void kv_scatter_kernel(const bf16* __restrict__ qkv, const int* __restrict__ slot,
                                  bf16* __restrict__ kpool, bf16* __restrict__ vpool, int d) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x;
  const long long s = slot[n];
  const uint4* src = reinterpret_cast<const uint4*>(qkv + (long long)n * 3 * d + d);
  uint4* kd = reinterpret_cast<uint4*>(kpool + s * d);
  uint4* vd = reinterpret_cast<uint4*>(vpool + s * d);
  const int nv = d >> 3;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    kd[i] = src[i];
    vd[i] = src[nv + i];
  }
}

void kv_scatter(const bf16* qkv, const int* slot, bf16* kpool, bf16* vpool, int n, int d,
                cudaStream_t stream) {
  if (n <= 0) return;
  launch_pdl(kv_scatter_kernel, dim3(n), dim3(128), 0, stream, qkv, slot, kpool, vpool, d);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------- split-K finalize
// out[row][c] = bf16(act(bias[c] + sum_s partial[s][row][c])) for GEMMs whose consumer is not a LayerNorm
// (slabs added in order: deterministic).  With slot != null the row is a fused QKV projection: columns
// [0, d) (Q) go to out, [d, 2d) to kpool[slot[row]], [2d, 3d) to vpool[slot[row]] -- the KV-cache scatter
// of the decode step folded into the reduction.
__global__ void __launch_bounds__(256) split_finalize_kernel(const float* __restrict__ partial, int n_splits,
                                                             long long stride, const float* __restrict__ bias, int N,
                                                             int act, bf16* __restrict__ out, long long ldo,
                                                             const int* __restrict__ slot, bf16* __restrict__ kpool,
                                                             bf16* __restrict__ vpool, int d) {
  pdl_trigger();
  const int row = blockIdx.x;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < N) a = __ldg(reinterpret_cast<const float4*>(bias + c));  // weights: safe ahead of the wait
  long long sl = 0;
  if (slot) sl = slot[row];  // pass metadata: written long before the kernel ahead of this one started
  pdl_wait();
  if (c >= N) return;
  const float* pp = partial + (long long)row * N + c;
#pragma unroll 4
  for (int s = 0; s < n_splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(pp + s * stride);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  if (act == 1) { a.x = gelu_erf_k(a.x); a.y = gelu_erf_k(a.y); a.z = gelu_erf_k(a.z); a.w = gelu_erf_k(a.w); }
  __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  bf16* dst = out + (long long)row * ldo + c;
  if (slot && c >= d) {
    dst = c < 2 * d ? kpool + sl * d + (c - d) : vpool + sl * d + (c - 2 * d);
  }
  *reinterpret_cast<uint2*>(dst) = u;
}

void split_finalize(const float* partial, int n_splits, long long stride, const float* bias, int rows, int N, int act,
                    bf16* out, long long ldo, const int* slot, bf16* kpool, bf16* vpool, int d, cudaStream_t stream) {
  if (rows <= 0) return;
  B200W_REQUIRE(N % 4 == 0 && ldo % 4 == 0 && (!slot || d % 4 == 0), "split_finalize: widths must be multiples of 4");
  launch_pdl(split_finalize_kernel, dim3(rows, cdiv(N / 4, 256)), dim3(256), 0, stream, partial, n_splits, stride, bias, N,
             act, out, ldo, slot, kpool, vpool, d);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------- sampling
__device__ __forceinline__ float block_reduce_max(float v, float* s) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, s[w]);
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* s) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += s[w];
  return r;
}

__device__ __forceinline__ bool token_masked_m(int v, unsigned mbyte, const SampleParams& p, const int4 st) {
  // st.x flags, st.y ts_lo.  Written as straight-line predicate logic (no branches: it runs per logit).
  const bool first = (st.x & SF_FIRST) != 0, rules = (st.x & SF_RULES) != 0;
  const bool is_ts = v >= p.timestamp_begin;
  bool barred = (mbyte & 1) != 0;                          // static suppress list
  barred |= first & ((mbyte & 2) != 0);                    // suppress_blank at the first position
  bool r = v == p.no_timestamps;
  r |= is_ts & (((st.x & SF_SUP_ALL_TS) != 0)              // after a timestamp pair: no timestamp
                | (v < st.y)                               // timestamps must not decrease
                | (first & (v > p.timestamp_begin + p.max_initial_ts)));
  r |= !is_ts & (first                                     // first token must be a timestamp
                 | (((st.x & SF_SUP_TEXT) != 0) & (v < p.eot)));  // after a lone timestamp: ts or EOT
  return barred | (rules & r);
}
__device__ __forceinline__ bool token_masked(int v, const SampleParams& p, const int4 st,
                                             const unsigned char* mask) {
  return token_masked_m(v, mask[v], p, st);
}

// One CTA (512 threads) per row: masked log-softmax with Whisper's timestamp rules, then top-k.
// Pass 1 applies the rules once and writes the masked logits back (-inf where barred), so the later
// passes are plain streaming reads of an L2-resident row.
static constexpr int SAMPLE_THREADS = 512;
static constexpr int SAMPLE_WARPS = SAMPLE_THREADS / 32;

__device__ __forceinline__ float block_max_1024(float v, float* s) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x & 31) < SAMPLE_WARPS ? s[threadIdx.x & 31] : -INFINITY;
  for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}
__device__ __forceinline__ float block_sum_1024(float v, float* s) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x & 31) < SAMPLE_WARPS ? s[threadIdx.x & 31] : 0.f;
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

__global__ void __launch_bounds__(SAMPLE_THREADS) sample_kernel(const SampleParams p) {
  __shared__ float s_red[32];
  __shared__ float s_cv[32];
  __shared__ int s_ci[32];
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x, tid = threadIdx.x;
  float* lg = const_cast<float*>(p.logits) + (long long)row * p.ld;
  const int4 st = p.state[row];
  const int V = p.n_vocab;

  // rows are 16-byte aligned (ld % 8 == 0): stream them as float4, the ragged tail as scalars
  const int nv4 = V >> 2;
  float4* lg4 = reinterpret_cast<float4*>(lg);
  const uchar4* mk4 = reinterpret_cast<const uchar4*>(p.mask);
  const int tb = p.timestamp_begin;
  float mx = -INFINITY, mx_text = -INFINITY;
  for (int i = tid; i < nv4; i += SAMPLE_THREADS) {
    float4 x = lg4[i];
    const uchar4 m = mk4[i];
    const int v = 4 * i;
    const bool k0 = token_masked_m(v, m.x, p, st), k1 = token_masked_m(v + 1, m.y, p, st),
               k2 = token_masked_m(v + 2, m.z, p, st), k3 = token_masked_m(v + 3, m.w, p, st);
    if (k0 | k1 | k2 | k3) {
      if (k0) x.x = -INFINITY;
      if (k1) x.y = -INFINITY;
      if (k2) x.z = -INFINITY;
      if (k3) x.w = -INFINITY;
      lg4[i] = x;
    }
    mx = fmaxf(fmaxf(mx, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
    if (v < tb) mx_text = fmaxf(mx_text, x.x);
    if (v + 1 < tb) mx_text = fmaxf(mx_text, x.y);
    if (v + 2 < tb) mx_text = fmaxf(mx_text, x.z);
    if (v + 3 < tb) mx_text = fmaxf(mx_text, x.w);
  }
  for (int v = 4 * nv4 + tid; v < V; v += SAMPLE_THREADS) {
    float x = lg[v];
    if (token_masked(v, p, st, p.mask)) { x = -INFINITY; lg[v] = x; }
    mx = fmaxf(mx, x);
    if (v < tb) mx_text = fmaxf(mx_text, x);
  }
  mx = block_max_1024(mx, s_red);
  mx_text = block_max_1024(mx_text, s_red);
  float sum_all = 0.f, sum_ts = 0.f;
  for (int i = tid; i < nv4; i += SAMPLE_THREADS) {
    const float4 x = lg4[i];
    const int v = 4 * i;
    const float e0 = __expf(x.x - mx), e1 = __expf(x.y - mx), e2 = __expf(x.z - mx), e3 = __expf(x.w - mx);
    sum_all += (e0 + e1) + (e2 + e3);
    if (v + 3 >= tb) sum_ts += (v >= tb ? e0 : 0.f) + (v + 1 >= tb ? e1 : 0.f) + (v + 2 >= tb ? e2 : 0.f) + e3;
  }
  for (int v = 4 * nv4 + tid; v < V; v += SAMPLE_THREADS) {
    const float e = __expf(lg[v] - mx);
    sum_all += e;
    if (v >= tb) sum_ts += e;
  }
  sum_all = block_sum_1024(sum_all, s_red);
  sum_ts = block_sum_1024(sum_ts, s_red);
  float lse = mx + logf(sum_all);
  int v_lo = 0;
  if ((st.x & SF_RULES) && sum_ts > 0.f) {
    // if the total timestamp mass beats every single text token, force a timestamp
    const float ts_logprob = mx + logf(sum_ts) - lse;
    const float text_logprob = mx_text - lse;
    if (ts_logprob > text_logprob) {
      v_lo = tb;
      lse = mx + logf(sum_ts);
    }
  }

  // per-thread top-k (k <= MAX_TOPK), kept sorted descending; ties -> lower token id first.
  // The lists live in registers: an insert is a fully unrolled shift chain (no dynamic indexing, which
  // would put the arrays in local memory); after the first few elements almost nothing passes `thresh`.
  float bv[MAX_TOPK];
  int bi[MAX_TOPK];
#pragma unroll
  for (int j = 0; j < MAX_TOPK; ++j) { bv[j] = -INFINITY; bi[j] = 0x7fffffff; }
  const int k = p.k;
  float thresh = -INFINITY;  // == bv[k-1]
  auto consider = [&](float xr, int v) {
    const float x = xr - lse;
    if (v >= v_lo && x > thresh) {
      // tokens are visited in increasing id order per thread, so a strict '>' keeps the lower id first on ties
#pragma unroll
      for (int j = MAX_TOPK - 1; j > 0; --j) {
        const bool shift = j < k && x > bv[j - 1];      // element j-1 moves down to j
        const bool here = j < k && !shift && x > bv[j];  // x lands at j
        if (shift) { bv[j] = bv[j - 1]; bi[j] = bi[j - 1]; }
        else if (here) { bv[j] = x; bi[j] = v; }
      }
      if (x > bv[0]) { bv[0] = x; bi[0] = v; }
#pragma unroll
      for (int j = 0; j < MAX_TOPK; ++j)
        if (j == k - 1) thresh = bv[j];
    }
  };
  for (int i = tid; i < nv4; i += SAMPLE_THREADS) {
    const float4 x = lg4[i];
    const int v = 4 * i;
    if (fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)) - lse > thresh) {
      consider(x.x, v); consider(x.y, v + 1); consider(x.z, v + 2); consider(x.w, v + 3);
    }
  }
  for (int v = 4 * nv4 + tid; v < V; v += SAMPLE_THREADS) consider(lg[v], v);
  int head = 0;
  for (int r = 0; r < k; ++r) {
    // block argmax over every thread's current head; ties -> lower id
    float cv = -INFINITY;
    int ci = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < MAX_TOPK; ++j)
      if (j == head && j < k) { cv = bv[j]; ci = bi[j]; }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, cv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ci, o);
      if (ov > cv || (ov == cv && oi < ci)) { cv = ov; ci = oi; }
    }
    __syncthreads();
    if ((tid & 31) == 0) { s_cv[tid >> 5] = cv; s_ci[tid >> 5] = ci; }
    __syncthreads();
    cv = (tid & 31) < SAMPLE_WARPS ? s_cv[tid & 31] : -INFINITY;
    ci = (tid & 31) < SAMPLE_WARPS ? s_ci[tid & 31] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, cv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ci, o);
      if (ov > cv || (ov == cv && oi < ci)) { cv = ov; ci = oi; }
    }
    if (tid == 0) {
      p.out_ids[row * k + r] = ci;
      p.out_logprobs[row * k + r] = cv;
    }
    int my = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < MAX_TOPK; ++j)
      if (j == head && j < k) my = bi[j];
    if (my == ci && ci != 0x7fffffff) ++head;
  }
}

// The same computation with the row kept on chip: the CTA streams its row from global memory ONCE (masking and the
// maxima on the way in, 8 loads in flight per thread), the exponential sums and the top-k selection then read shared
// memory.  The streaming kernel above keeps one 16-byte load in flight per thread and pays a round trip per loop
// iteration, three passes over (ncu: 0.9 % of DRAM throughput, 9.7 long-scoreboard stalls per issue).  What bounds
// this kernel is the instruction count (it runs once per logit), so:
//  * logits below EOT (97 % of the row) take a short masking path: the rule outcome for text tokens is per row;
//  * top-k: one sorted list per WARP (entry j in lane j).  A candidate enters by ballot / shuffle only when it
//    beats the warp's k-th value AND a floor known before the scan starts: every thread keeps the maximum of its
//    own elements, a warp takes the k-th largest of its 32 thread maxima (at least k elements are >= it) and the
//    CTA the largest such value over its warps.  About k + a few dozen of the 51 866 elements pass.
// Reductions have the shape of the streaming kernel's (same threads, same element-to-thread map, same trees), so
// maxima, sums and therefore every log-probability are bit-identical to it.
template <bool SAMPLING>
__global__ void __launch_bounds__(SAMPLE_THREADS, 1) sample_row_kernel(const SampleParams p) {
  extern __shared__ __align__(16) float s_row[];  // [ceil(V / 4) * 4]
  __shared__ float s_red[32];
  __shared__ float s_lv[SAMPLE_WARPS][MAX_TOPK];
  __shared__ int s_li[SAMPLE_WARPS][MAX_TOPK];
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* lg = p.logits + (long long)row * p.ld;
  const int4 st = p.state[row];
  const int V = p.n_vocab;
  const int nv4 = V >> 2;
  const float4* lg4 = reinterpret_cast<const float4*>(lg);
  float4* s4 = reinterpret_cast<float4*>(s_row);
  const uchar4* mk4 = reinterpret_cast<const uchar4*>(p.mask);
  const int tb = p.timestamp_begin;
  // text tokens (v < eot <= no_timestamps, timestamp_begin): barred by the rules as a class, or by their mask bits
  const bool rules = (st.x & SF_RULES) != 0, first = (st.x & SF_FIRST) != 0;
  const bool text_barred = rules && (first || (st.x & SF_SUP_TEXT) != 0);
  const unsigned mbits = first ? 3u : 1u;
  const int text_end = min(min(p.eot, p.no_timestamps), tb);  // [0, text_end): the short path is exact
  float mx = -INFINITY, mx_text = -INFINITY, mx_ts = -INFINITY;   // thread maxima: all / text / timestamps
  constexpr int UB = 8;  // loads in flight per thread (the compiler does not batch them by itself: ~26 serial round trips)
  for (int i0 = tid; i0 < nv4; i0 += SAMPLE_THREADS * UB) {
    float4 xs[UB];
    uchar4 ms[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int i = i0 + u * SAMPLE_THREADS;
      if (i < nv4) { xs[u] = __ldg(lg4 + i); ms[u] = __ldg(mk4 + i); }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int i = i0 + u * SAMPLE_THREADS;
      if (i < nv4) {
        float4 x = xs[u];
        const uchar4 m = ms[u];
        const int v = 4 * i;
        if (v + 3 < text_end) {
          if (text_barred || (m.x & mbits)) x.x = -INFINITY;
          if (text_barred || (m.y & mbits)) x.y = -INFINITY;
          if (text_barred || (m.z & mbits)) x.z = -INFINITY;
          if (text_barred || (m.w & mbits)) x.w = -INFINITY;
          const float m4 = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
          mx = fmaxf(mx, m4);
          mx_text = fmaxf(mx_text, m4);
        } else {
          if (token_masked_m(v, m.x, p, st)) x.x = -INFINITY;
          if (token_masked_m(v + 1, m.y, p, st)) x.y = -INFINITY;
          if (token_masked_m(v + 2, m.z, p, st)) x.z = -INFINITY;
          if (token_masked_m(v + 3, m.w, p, st)) x.w = -INFINITY;
          mx = fmaxf(fmaxf(mx, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
          if (v < tb) mx_text = fmaxf(mx_text, x.x); else mx_ts = fmaxf(mx_ts, x.x);
          if (v + 1 < tb) mx_text = fmaxf(mx_text, x.y); else mx_ts = fmaxf(mx_ts, x.y);
          if (v + 2 < tb) mx_text = fmaxf(mx_text, x.z); else mx_ts = fmaxf(mx_ts, x.z);
          if (v + 3 < tb) mx_text = fmaxf(mx_text, x.w); else mx_ts = fmaxf(mx_ts, x.w);
        }
        s4[i] = x;
      }
    }
  }
  for (int v = 4 * nv4 + tid; v < V; v += SAMPLE_THREADS) {
    float x = lg[v];
    if (token_masked(v, p, st, p.mask)) x = -INFINITY;
    s_row[v] = x;
    mx = fmaxf(mx, x);
    if (v < tb) mx_text = fmaxf(mx_text, x); else mx_ts = fmaxf(mx_ts, x);
  }
  const float my_all = mx, my_ts = mx_ts;
  mx = block_max_1024(mx, s_red);  // (the barriers inside also publish the row)
  mx_text = block_max_1024(mx_text, s_red);
  // sampling: the distribution is softmax(x / T); everything below works on y = (x - max) / T
  const float inv_t = SAMPLING ? 1.0f / p.temperature : 1.0f;
  float sum_all = 0.f, sum_ts = 0.f;
  for (int i = tid; i < nv4; i += SAMPLE_THREADS) {
    const float4 x = s4[i];
    const int v = 4 * i;
    float e0, e1, e2, e3;
    if (SAMPLING) {
      e0 = __expf((x.x - mx) * inv_t); e1 = __expf((x.y - mx) * inv_t);
      e2 = __expf((x.z - mx) * inv_t); e3 = __expf((x.w - mx) * inv_t);
    } else {
      e0 = __expf(x.x - mx); e1 = __expf(x.y - mx); e2 = __expf(x.z - mx); e3 = __expf(x.w - mx);
    }
    sum_all += (e0 + e1) + (e2 + e3);
    if (v + 3 >= tb) sum_ts += (v >= tb ? e0 : 0.f) + (v + 1 >= tb ? e1 : 0.f) + (v + 2 >= tb ? e2 : 0.f) + e3;
  }
  for (int v = 4 * nv4 + tid; v < V; v += SAMPLE_THREADS) {
    const float e = __expf((s_row[v] - mx) * inv_t);
    sum_all += e;
    if (v >= tb) sum_ts += e;
  }
  const float my_sum_all = sum_all, my_sum_ts = sum_ts;  // this thread's part (the sampler's inverse CDF)
  sum_all = block_sum_1024(sum_all, s_red);
  sum_ts = block_sum_1024(sum_ts, s_red);
  float lse = mx + logf(sum_all);
  int v_lo = 0;
  if (rules && sum_ts > 0.f) {
    const float ts_logprob = SAMPLING ? logf(sum_ts) - logf(sum_all) : mx + logf(sum_ts) - lse;
    const float text_logprob = SAMPLING ? (mx_text - mx) * inv_t - logf(sum_all) : mx_text - lse;
    if (ts_logprob > text_logprob) {
      v_lo = tb;
      lse = mx + logf(sum_ts);
    }
  }

  if (SAMPLING) {
    // inverse CDF in thread-major order (thread, then its elements in the order of the passes above): the thread whose
    // interval of the running sum contains u * total walks its elements again
    __shared__ float s_scan[SAMPLE_WARPS];
    __shared__ int s_pick[2];
    const float mine = v_lo == 0 ? my_sum_all : my_sum_ts, total = v_lo == 0 ? sum_all : sum_ts;
    float incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const float t2 = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t2;
    }
    if (lane == 31) s_scan[warp] = incl;
    if (tid == 0) { s_pick[0] = 0x7fffffff; s_pick[1] = -1; }
    __syncthreads();
    float base = 0.f;
    for (int w = 0; w < warp; ++w) base += s_scan[w];
    const float hi = base + incl, lo = hi - mine;
    // u in [0, 1): splitmix64 of (seed, sample stream, draw index)
    unsigned long long z = ((unsigned long long)p.seed << 32) ^ ((unsigned long long)(unsigned)st.w << 20) ^ (unsigned)st.z;
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
    const float target = u * total;
    auto walk = [&](float from, bool take_last) {  // this thread's elements; returns through s_pick
      float run = from;
      int last = -1;
      for (int i = tid; i < nv4; i += SAMPLE_THREADS) {
        const float4 x = s4[i];
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int v = 4 * i + c;
          if (v >= v_lo && xs[c] > -INFINITY) {
            run += __expf((xs[c] - mx) * inv_t);
            last = v;
            if (!take_last && run > target) return v;
          }
        }
      }
      for (int v = 4 * nv4 + tid; v < V; v += SAMPLE_THREADS) {
        const float x = s_row[v];
        if (v >= v_lo && x > -INFINITY) {
          run += __expf((x - mx) * inv_t);
          last = v;
          if (!take_last && run > target) return v;
        }
      }
      return last;
    };
    int pick = -1;
    if (mine > 0.f && target >= lo && target < hi) pick = walk(lo, false);
    if (pick >= 0) atomicMin(&s_pick[0], pick);                 // (rounding: more than one interval may claim it)
    if (mine > 0.f) atomicMax(&s_pick[1], tid);                  // fallback: the last thread holding any mass
    __syncthreads();
    if (s_pick[0] == 0x7fffffff && tid == s_pick[1]) {           // target fell on the rounding gap at the very end
      const int v = walk(0.f, true);
      if (v >= 0) s_pick[0] = v;
    }
    __syncthreads();
    if (tid == 0) {
      const int v = s_pick[0];
      p.out_ids[row] = v;
      p.out_logprobs[row] = v == 0x7fffffff ? -INFINITY : (s_row[v] - mx) * inv_t - logf(total);
    }
    return;
  }

  const int k = p.k;
  // floor of the k-th best log-probability (see the header)
  float floor0;
  {
    const float c = (v_lo == 0 ? my_all : my_ts) - lse;  // this thread's best admissible candidate
    int rank = 0;
    for (int o = 1; o < 32; ++o) {
      const int src = (lane + o) & 31;
      const float ov = __shfl_sync(0xffffffffu, c, src);
      rank += (ov > c || (ov == c && src < lane)) ? 1 : 0;
    }
    const unsigned at = __ballot_sync(0xffffffffu, rank == k - 1);  // exactly one lane (k <= 32)
    floor0 = block_max_1024(__shfl_sync(0xffffffffu, c, __ffs(at) - 1), s_red);
  }

  // warp-level top-k: lane j holds the warp's j-th best (value, id); ids are visited in increasing order by a warp
  // (iterations, then lanes, then the 4 elements), so a strict '>' keeps the lower id first on ties
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  float thresh = -INFINITY;  // value of entry k-1
  auto insert = [&](float x, int v) {  // warp-uniform call
    const unsigned ge = __ballot_sync(0xffffffffu, lane < k && bv >= x);
    const int pos = __popc(ge);  // entries that stay ahead of x (they are a prefix: the list is sorted)
    const float uv = __shfl_up_sync(0xffffffffu, bv, 1);
    const int ui = __shfl_up_sync(0xffffffffu, bi, 1);
    if (lane == pos) { bv = x; bi = v; }
    else if (lane > pos) { bv = uv; bi = ui; }
    thresh = __shfl_sync(0xffffffffu, bv, k - 1);
  };
  const int i_lo = v_lo >> 2;  // float4 groups below it hold no admissible token
  for (int i0 = warp * 32; i0 < nv4; i0 += SAMPLE_THREADS) {  // i = i0 + lane: the element map of the passes above
    const int i = i0 + lane;
    float c0 = -INFINITY, c1 = -INFINITY, c2 = -INFINITY, c3 = -INFINITY;
    if (i < nv4 && i >= i_lo) {
      const float4 x = s4[i];
      const int v = 4 * i;
      c0 = x.x - lse; c1 = x.y - lse; c2 = x.z - lse; c3 = x.w - lse;
      if (v < v_lo) {  // the group straddling v_lo
        c0 = -INFINITY;
        if (v + 1 < v_lo) c1 = -INFINITY;
        if (v + 2 < v_lo) c2 = -INFINITY;
      }
    }
    const float c4 = fmaxf(fmaxf(c0, c1), fmaxf(c2, c3));
    unsigned m = __ballot_sync(0xffffffffu, c4 > thresh && c4 >= floor0);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float y0 = __shfl_sync(0xffffffffu, c0, src), y1 = __shfl_sync(0xffffffffu, c1, src),
                  y2 = __shfl_sync(0xffffffffu, c2, src), y3 = __shfl_sync(0xffffffffu, c3, src);
      const int vs = 4 * (i0 + src);
      if (y0 > thresh && y0 >= floor0) insert(y0, vs);
      if (y1 > thresh && y1 >= floor0) insert(y1, vs + 1);
      if (y2 > thresh && y2 >= floor0) insert(y2, vs + 2);
      if (y3 > thresh && y3 >= floor0) insert(y3, vs + 3);
    }
  }
  if (warp == 0) {  // ragged tail (V % 4 elements)
    const int v = 4 * nv4 + lane;
    const float c = (v < V && v >= v_lo) ? s_row[v] - lse : -INFINITY;
    unsigned m = __ballot_sync(0xffffffffu, c > thresh);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float y = __shfl_sync(0xffffffffu, c, src);
      if (y > thresh) insert(y, 4 * nv4 + src);
    }
  }
  if (lane < MAX_TOPK) { s_lv[warp][lane] = bv; s_li[warp][lane] = bi; }
  __syncthreads();
  if (warp == 0) {  // merge: lane l walks warp l's sorted list
    int head = 0;
    for (int r = 0; r < k; ++r) {
      const bool has = lane < SAMPLE_WARPS && head < k;
      float cv = has ? s_lv[lane][head] : -INFINITY;
      int ci = has ? s_li[lane][head] : 0x7fffffff;
      const int mine = ci;
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, cv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ci, o);
        if (ov > cv || (ov == cv && oi < ci)) { cv = ov; ci = oi; }
      }
      if (lane == 0) {
        p.out_ids[row * k + r] = ci;
        p.out_logprobs[row * k + r] = cv;
      }
      if (has && mine == ci && ci != 0x7fffffff) ++head;
    }
  }
}

void sample_topk(const SampleParams& p, int rows, cudaStream_t stream, bool streaming) {
  if (rows <= 0) return;
  B200W_REQUIRE(p.k >= 1 && p.k <= MAX_TOPK, "top-k out of range");
  const size_t row_bytes = (size_t)((p.n_vocab + 3) / 4) * 16;
  static const bool streaming_only = getenv("B200W_SAMPLE_STREAMING") != nullptr;
  if (p.temperature > 0.f) {
    B200W_REQUIRE(p.k == 1, "sampling draws one token per row");
    B200W_REQUIRE(row_bytes <= 220 * 1024, "sampling needs the row in shared memory (vocabulary too large)");
    ensure_dynamic_smem(sample_row_kernel<true>, row_bytes);
    launch_pdl(sample_row_kernel<true>, dim3(rows), dim3(SAMPLE_THREADS), row_bytes, stream, p);
  } else if (!streaming && !streaming_only && row_bytes <= 220 * 1024) {
    ensure_dynamic_smem(sample_row_kernel<false>, row_bytes);
    launch_pdl(sample_row_kernel<false>, dim3(rows), dim3(SAMPLE_THREADS), row_bytes, stream, p);
  } else {
    launch_pdl(sample_kernel, dim3(rows), dim3(SAMPLE_THREADS), 0, stream, p);
  }
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// softmax probability of given tokens over raw logits rows.  One CTA per (row).
// mode 0: out[row] = softmax(logits[row])[tok[row]]
// mode 1: out[row][j] = softmax restricted to ids[0..n_ids) (language detection)
__global__ void __launch_bounds__(256) token_prob_kernel(const float* __restrict__ logits, int ld,
                                                         int n_vocab, const int* __restrict__ rows,
                                                         const int* __restrict__ tok,
                                                         const int* __restrict__ ids, int n_ids,
                                                         float* __restrict__ out) {
  __shared__ float s_red[8];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* lg = logits + (long long)(rows ? rows[r] : r) * ld;
  if (ids == nullptr) {
    float mx = -INFINITY;
    for (int v = tid; v < n_vocab; v += 256) mx = fmaxf(mx, lg[v]);
    mx = block_reduce_max(mx, s_red);
    float sum = 0.f;
    for (int v = tid; v < n_vocab; v += 256) sum += __expf(lg[v] - mx);
    sum = block_reduce_sum(sum, s_red);
    if (tid == 0) out[r] = __expf(lg[tok[r]] - mx) / sum;
  } else {
    float mx = -INFINITY;
    for (int j = tid; j < n_ids; j += 256) mx = fmaxf(mx, lg[ids[j]]);
    mx = block_reduce_max(mx, s_red);
    float sum = 0.f;
    for (int j = tid; j < n_ids; j += 256) sum += __expf(lg[ids[j]] - mx);
    sum = block_reduce_sum(sum, s_red);
    for (int j = tid; j < n_ids; j += 256) out[(long long)r * n_ids + j] = __expf(lg[ids[j]] - mx) / sum;
  }
}

void token_probs(const float* logits, int ld, int n_vocab, const int* rows, const int* tok,
                 const int* ids, int n_ids, float* out, int n_rows, cudaStream_t stream) {
  if (n_rows <= 0) return;
  token_prob_kernel<<<n_rows, 256, 0, stream>>>(logits, ld, n_vocab, rows, tok, ids, n_ids, out);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------- word timing
// w: [n_align][n_q][stride] attention probabilities; sequence b owns query rows [q0, q0 + n_tok)
// and frames [0, nf).  Standardise every (head, frame) column over the tokens (population std).
__global__ void align_normalize_kernel(float* __restrict__ w, int n_q, int stride,
                                       const int* __restrict__ seq_q0, const int* __restrict__ seq_ntok,
                                       const int* __restrict__ seq_nf) {
  const int b = blockIdx.z, a = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int nf = seq_nf[b], nt = seq_ntok[b];
  if (f >= nf) return;
  float* col = w + ((long long)a * n_q + seq_q0[b]) * stride + f;
  float mean = 0.f;
  for (int t = 0; t < nt; ++t) mean += col[(long long)t * stride];
  mean /= (float)nt;
  float var = 0.f;
  for (int t = 0; t < nt; ++t) {
    const float dlt = col[(long long)t * stride] - mean;
    var += dlt * dlt;
  }
  const float stdv = sqrtf(var / (float)nt);
  for (int t = 0; t < nt; ++t) col[(long long)t * stride] = (col[(long long)t * stride] - mean) / stdv;
}

// matrix[b][t][f] = mean over heads of median_{width}(w[a][t][f-h..f+h]) (reflect padding)
__global__ void align_median_mean_kernel(const float* __restrict__ w, int n_align, int n_q, int stride,
                                         const int* __restrict__ seq_q0,
                                         const int* __restrict__ seq_ntok,
                                         const int* __restrict__ seq_nf, int width,
                                         float* __restrict__ matrix, int mat_stride) {
  const int b = blockIdx.z, t = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int nf = seq_nf[b];
  if (t >= seq_ntok[b] || f >= nf) return;
  const int half = width / 2;
  float acc = 0.f;
  for (int a = 0; a < n_align; ++a) {
    const float* row = w + ((long long)a * n_q + seq_q0[b] + t) * stride;
    float val;
    if (nf <= half || width <= 1) {
      val = row[f];  // too short to filter (OpenAI/CT2 skip the filter)
    } else {
      float win[9];
      for (int j = 0; j < width; ++j) {
        int g = f - half + j;
        if (g < 0) g = -g;
        if (g >= nf) g = 2 * (nf - 1) - g;
        win[j] = row[g];
      }
      for (int i = 1; i < width; ++i) {  // insertion sort
        const float x = win[i];
        int j = i - 1;
        while (j >= 0 && win[j] > x) { win[j + 1] = win[j]; --j; }
        win[j + 1] = x;
      }
      val = win[half];
    }
    acc += val;
  }
  matrix[((long long)seq_q0[b] + t) * mat_stride + f] = acc / (float)n_align;
}

// DTW over cost = -matrix[rows r0 .. r0+N) x [0, M), one CTA per sequence (anti-diagonal wavefront).
// trace: int8 [(N+1)][(M+1)] scratch per sequence.  path: (text, time) pairs in forward order.
__global__ void __launch_bounds__(512) dtw_kernel(const float* __restrict__ matrix, int mat_stride,
                                                  const int* __restrict__ seq_q0,
                                                  const int* __restrict__ seq_ntok,
                                                  const int* __restrict__ seq_nf, int row_lo,
                                                  int row_hi_trim, signed char* __restrict__ trace_all,
                                                  long long trace_stride, int* __restrict__ path,
                                                  int max_path, int* __restrict__ path_len) {
  extern __shared__ float s_diag[];  // 3 x (Nmax + 1)
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = seq_ntok[b] - row_lo - row_hi_trim;  // text rows
  const int M = seq_nf[b];
  int* out = path + (long long)b * max_path * 2;
  if (N <= 0 || M <= 0) {
    if (tid == 0) path_len[b] = 0;
    return;
  }
  const float* x = matrix + ((long long)seq_q0[b] + row_lo) * mat_stride;
  signed char* trace = trace_all + (long long)b * trace_stride;
  const int W = M + 1;
  const int NP = N + 1;
  float* d0 = s_diag;            // diagonal d-2
  float* d1 = s_diag + NP;       // diagonal d-1
  float* d2 = s_diag + 2 * NP;   // diagonal d   (indexed by i)
  for (int i = tid; i < NP; i += blockDim.x) { d0[i] = INFINITY; d1[i] = INFINITY; d2[i] = INFINITY; }
  __syncthreads();
  if (tid == 0) d0[0] = 0.f;  // cost[0][0] lives on diagonal 0
  __syncthreads();
  // diagonal 1 holds cost[0][1] and cost[1][0] = inf (already)
  for (int d = 2; d <= N + M; ++d) {
    const int ilo = max(1, d - M), ihi = min(N, d - 1);
    for (int i = ilo + tid; i <= ihi; i += blockDim.x) {
      const int j = d - i;
      const float c0 = d0[i - 1];  // cost[i-1][j-1]
      const float c1 = d1[i - 1];  // cost[i-1][j]
      const float c2 = d1[i];      // cost[i][j-1]
      float c;
      signed char t;
      if (c0 < c1 && c0 < c2) { c = c0; t = 0; }
      else if (c1 < c0 && c1 < c2) { c = c1; t = 1; }
      else { c = c2; t = 2; }
      d2[i] = -x[(long long)(i - 1) * mat_stride + (j - 1)] + c;
      trace[(long long)i * W + j] = t;
    }
    __syncthreads();
    // rotate: entries outside the active range of the new d1/d0 must read as inf
    float* tmp = d0; d0 = d1; d1 = d2; d2 = tmp;
    for (int i = tid; i < NP; i += blockDim.x) d2[i] = INFINITY;
    if (tid == 0 && d == 2) { /* cost[0][*] = inf for * > 0 */ d0[0] = INFINITY; }
    __syncthreads();
  }
  if (tid == 0) {
    int i = N, j = M, n = 0;
    // backtrace; written reversed into the tail of the buffer, then shifted to the front
    int* tmpbuf = out;
    while ((i > 0 || j > 0) && n < max_path) {
      tmpbuf[2 * n] = i - 1;
      tmpbuf[2 * n + 1] = j - 1;
      ++n;
      signed char t;
      if (i == 0) t = 2; else if (j == 0) t = 1; else t = trace[(long long)i * W + j];
      if (t == 0) { --i; --j; } else if (t == 1) { --i; } else { --j; }
    }
    for (int a = 0, z = n - 1; a < z; ++a, --z) {
      const int t0 = tmpbuf[2 * a], t1 = tmpbuf[2 * a + 1];
      tmpbuf[2 * a] = tmpbuf[2 * z]; tmpbuf[2 * a + 1] = tmpbuf[2 * z + 1];
      tmpbuf[2 * z] = t0; tmpbuf[2 * z + 1] = t1;
    }
    path_len[b] = n;
  }
}

void align_postprocess(float* weights, int n_align, int n_q, int stride, const int* seq_q0,
                       const int* seq_ntok, const int* seq_nf, int n_seq, int max_ntok, int max_nf,
                       int median_width, float* matrix, int row_lo, int row_hi_trim,
                       signed char* trace, long long trace_stride, int* path, int max_path,
                       int* path_len, cudaStream_t stream) {
  if (n_seq <= 0) return;
  B200W_REQUIRE(median_width <= 9 && (median_width & 1) == 1, "median width must be odd and <= 9");
  align_normalize_kernel<<<dim3(cdiv(max_nf, 128), n_align, n_seq), 128, 0, stream>>>(
      weights, n_q, stride, seq_q0, seq_ntok, seq_nf);
  B200W_COUNT_LAUNCH();
  align_median_mean_kernel<<<dim3(cdiv(max_nf, 128), max_ntok, n_seq), 128, 0, stream>>>(
      weights, n_align, n_q, stride, seq_q0, seq_ntok, seq_nf, median_width, matrix, stride);
  B200W_COUNT_LAUNCH();
  const size_t smem = 3 * (size_t)(max_ntok + 1) * sizeof(float);
  dtw_kernel<<<n_seq, 512, smem, stream>>>(matrix, stride, seq_q0, seq_ntok, seq_nf, row_lo,
                                           row_hi_trim, trace, trace_stride, path, max_path, path_len);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------- page tables
// Beam fork: new[row[i]][0:pos[i]] = old[parent[i]][0:pos[i]]; new[row[i]][pos[i]] = slot[i].
__global__ void fork_tables_kernel(const int* __restrict__ old_t, const int* __restrict__ parent,
                                   const int* __restrict__ row, const int* __restrict__ pos,
                                   const int* __restrict__ slot, int* __restrict__ new_t, int stride) {
  const int i = blockIdx.x;
  const int* s = old_t + (long long)parent[i] * stride;
  int* d = new_t + (long long)row[i] * stride;
  const int len = pos[i];
  for (int j = threadIdx.x; j < len; j += blockDim.x) d[j] = s[j];
  if (threadIdx.x == 0) d[len] = slot[i];
}

void gather_fork_tables(const int* old_t, const int* parent, const int* row, const int* pos,
                        const int* slot, int* new_t, int n, int stride, cudaStream_t stream) {
  if (n <= 0) return;
  fork_tables_kernel<<<n, 128, 0, stream>>>(old_t, parent, row, pos, slot, new_t, stride);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

}  // namespace b200w

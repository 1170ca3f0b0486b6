// Skinny GEMM for decode steps (M <= 256 activation rows): out[M][N] = X[M][K] . W[N][K]^T.
//
// The weight matrix is the big operand (3-13 MB, streamed from HBM once per step), the activations
// are tiny, so the roles of the tcgen05 operands are swapped with respect to gemm.cu: 128 WEIGHT rows
// are the MMA's M dimension and all activation rows (padded to a multiple of 16) are its N dimension:
//     D[128 n][Mpad m] += W_tile[128 n][64 k] . X_tile[Mpad m][64 k]^T          (one MMA per 16 k)
// Consequences: no zero-padded 128-row activation tiles (a k-block costs 16 KB of weights + Mpad*128 B
// of activations instead of 8 KB + 32 KB), and the accumulator comes out transposed, so for a fixed
// m the 32 lanes of an epilogue warp hold 32 CONSECUTIVE n: every global store is a contiguous
// 128-byte (fp32) / 64-byte (bf16) row segment.
//
// One tile per CTA, grid = ceil(N/128) x splits.  split-K partial sums go to fp32 slabs
// [split][M][ldo] that the consumer reduces in slab order (LayerNorm, qkv_finalize, the cross-attention
// Q staging); without splits the epilogue applies bias (+GELU) and stores bf16 or fp32.
// Roles: warp 0 TMA producer (weights are requested ahead of the PDL wait), warp 1 MMA issuer,
// warps 2-9 epilogue (TMEM lane quarter = warp % 4, alternate 32-column chunks per half).
#include "skinny.h"

#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>

#include "gemm.h"
#include "ptx.cuh"

namespace b200w {

static constexpr int SK_BK = 64;
static constexpr int SK_THREADS = 320;
static constexpr int SK_MAX_STAGES = 12;

struct SkinnyParams {
  int M, N, Mpad;
  int total_kb, splits, n_tiles, stages;
  int tile_rows;  // weight rows per CTA: 128, or 64 (rows 64..127 of the MMA's A tile are then don't-care)
  const float* bias;
  void* out;
  long long ldo, split_stride;
  long long* trace;  // development aid: 16 stamps per CTA, or null
};
#define SK_TRACE(slot) do { if (p.trace) p.trace[blockIdx.x * 16 + (slot)] = clock64(); } while (0)
__device__ __forceinline__ long long sk_gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// Everything a launch needs, in GLOBAL memory: the kernel gets one pointer, fetches the parameter
// block with a single coalesced load and lets the TMA unit read the descriptors from the same
// record (measured 1.9 us faster per launch than a 350-byte kernel-parameter block, which lives in
// the constant bank and is cold in every freshly launched CTA).  Records are created once per
// distinct call (a decode step repeats the same calls on the same buffers) and are indexed by a SkinnyCache that
// its owner (an engine) frees; the records themselves come from alloc_descriptor_record(): their addresses are
// never reused (the TMA unit caches descriptors by address).
struct __align__(128) SkinnyLaunch {
  CUtensorMap tmW, tmX;
  SkinnyParams p;
};

// Out-of-line spin with a cheap spin-count watchdog: this kernel is launch-latency critical and every instruction of code
// has to be fetched from L2 by every freshly launched CTA (measured: ~1 us per 16 KB of executed
// straight-line code), so everything here is kept small and un-unrolled where it can be.
__device__ __noinline__ void sk_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins == (1u << 28)) __trap();  // a protocol bug must fault (launch error), never hang the GPU
  }
}

// ACT: 0 none, 1 exact-erf GELU.  OUT_FP32: fp32 (split partial sums / logits) or bf16.
template <int ACT, int OUT_FP32>
__global__ void __launch_bounds__(SK_THREADS, 1) skinny_gemm_kernel(const SkinnyLaunch* __restrict__ L) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ SkinnyParams p_sh;
  __shared__ __align__(8) uint64_t bars[2 * SK_MAX_STAGES + 1];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = clock64(), g_start = sk_gtime();
  // Everything that does not need the parameters starts before they arrive: the parameter fetch,
  // the descriptor prefetches, the barrier initialisation and the TMEM allocation overlap.
  if (threadIdx.x < sizeof(SkinnyParams) / 4)
    reinterpret_cast<uint32_t*>(&p_sh)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(&L->p) + threadIdx.x);
  const CUtensorMap* tmW = &L->tmW;
  const CUtensorMap* tmX = &L->tmX;
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * SK_MAX_STAGES, tfull = empty0 + 8 * SK_MAX_STAGES;
  if (threadIdx.x == 64) {
    tma_prefetch_desc(tmW);
    tma_prefetch_desc(tmX);
#pragma unroll 1
    for (int s = 0; s < 2 * SK_MAX_STAGES + 1; ++s) mbar_init(full0 + 8 * s, 1);
    fence_barrier_init();
  }
  pdl_trigger();
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const SkinnyParams p = p_sh;  // registers: later global stores must not force re-reads of the shared copy
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int w_bytes = p.tile_rows * 128;
  const int stage_bytes = w_bytes + p.Mpad * 128;
  const uint32_t tmem_base = tmem_slot;
  if (p.trace && threadIdx.x == 0) { p.trace[blockIdx.x * 16] = g_start; p.trace[blockIdx.x * 16 + 1] = t_start; SK_TRACE(2); }
  const int ks = blockIdx.x / p.n_tiles, nt = blockIdx.x - ks * p.n_tiles;
  const int n0 = nt * p.tile_rows;
  const int kb_lo = (int)((long long)p.total_kb * ks / p.splits);
  const int n_kb = (int)((long long)p.total_kb * (ks + 1) / p.splits) - kb_lo;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      // weights first: they do not depend on the kernel ahead of this one
      const int pre = n_kb < p.stages ? n_kb : p.stages;
#pragma unroll 1
      for (int i = 0; i < pre; ++i) {
        const uint32_t fb = full0 + 8 * i;
        mbar_arrive_expect_tx(fb, stage_bytes);
        tma_load_3d(smem_u32(smem + i * stage_bytes), tmW, fb, (kb_lo + i) * SK_BK, n0, 0);
      }
      pdl_wait();
      SK_TRACE(3);
      uint32_t stage = 0, phase = 0;
#pragma unroll 1
      for (int kb = 0; kb < n_kb; ++kb) {
        const uint32_t fb = full0 + 8 * stage;
        const uint32_t st = smem_u32(smem + stage * stage_bytes);
        if (kb >= pre) {
          sk_wait(empty0 + 8 * stage, phase ^ 1);
          mbar_arrive_expect_tx(fb, stage_bytes);
          tma_load_3d(st, tmW, fb, (kb_lo + kb) * SK_BK, n0, 0);
        }
        tma_load_3d(st + w_bytes, tmX, fb, (kb_lo + kb) * SK_BK, 0, 0);
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      const uint32_t idesc = make_idesc_bf16(128, p.Mpad, 0, 0);
      uint32_t stage = 0, phase = 0;
#pragma unroll 1
      for (int kb = 0; kb < n_kb; ++kb) {
        sk_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        if (kb == 0) SK_TRACE(4);
        const uint64_t wd = make_sw128_desc(smem_u32(smem + stage * stage_bytes));
        const uint64_t xd = make_sw128_desc(smem_u32(smem + stage * stage_bytes + w_bytes));
#pragma unroll
        for (int k = 0; k < SK_BK / 16; ++k) tc_mma_bf16(tmem_base, wd + 2 * k, xd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        tc_commit(empty0 + 8 * stage);
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
      }
      tc_commit(tfull);
      SK_TRACE(5);
    }
  } else {  // ---------------- epilogue: thread = weight row n, columns = activation rows m
    pdl_wait();
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int n = n0 + q * 32 + lane;
    const bool n_ok = n < p.N && q * 32 < p.tile_rows;
    const float bias = (p.bias && n_ok) ? __ldg(p.bias + n) : 0.f;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const long long ldo = p.ldo;
    const int M = p.M;
    sk_wait(tfull, 0);
    tc_fence_after();
    if (warp == 2 && lane == 0) SK_TRACE(6);
#pragma unroll 1
    for (int c0 = half * 32; c0 < p.Mpad; c0 += 64) {
      uint32_t v[32];
      tmem_ld32(taddr + c0, v);
      tmem_ld_wait();
      if (warp == 2 && lane == 0 && c0 < 192) SK_TRACE(10 + (c0 >> 6) * 2);
      if (!n_ok) continue;
      // Full chunks store unconditionally with a running pointer (4 instructions per element); the
      // per-element predicate of the ragged last chunk costs a branch + reconvergence per element.
      const int full = c0 + 32 <= M;
      if (OUT_FP32) {
        float* op = reinterpret_cast<float*>(p.out) + (long long)ks * p.split_stride + n + (long long)c0 * ldo;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float y = __uint_as_float(v[j]) + bias;
            if (ACT == 1) y = gelu_erf(y);
            *op = y;
            op += ldo;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float y = __uint_as_float(v[j]) + bias;
            if (ACT == 1) y = gelu_erf(y);
            if (c0 + j < M) *op = y;
            op += ldo;
          }
        }
      } else {
        bf16* op = reinterpret_cast<bf16*>(p.out) + n + (long long)c0 * ldo;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float y = __uint_as_float(v[j]) + bias;
            if (ACT == 1) y = gelu_erf(y);
            *op = __float2bfloat16(y);
            op += ldo;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float y = __uint_as_float(v[j]) + bias;
            if (ACT == 1) y = gelu_erf(y);
            if (c0 + j < M) *op = __float2bfloat16(y);
            op += ldo;
          }
        }
      }
      if (warp == 2 && lane == 0 && c0 < 192) SK_TRACE(11 + (c0 >> 6) * 2);
    }
  }
  if (warp == 2 && lane == 0) SK_TRACE(7);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
  if (p.trace && threadIdx.x == 0) { SK_TRACE(8); p.trace[blockIdx.x * 16 + 9] = sk_gtime(); }
}

struct SkinnyKey {
  const void *x, *w, *bias, *out;
  long long ldx, ldw, ldo, ss;
  int M, N, K, act, of, sp, tr;
  const void* trace;
  bool operator<(const SkinnyKey& o) const {
    return std::tie(x, w, bias, out, ldx, ldw, ldo, ss, M, N, K, act, of, sp, tr, trace) <
           std::tie(o.x, o.w, o.bias, o.out, o.ldx, o.ldw, o.ldo, o.ss, o.M, o.N, o.K, o.act, o.of, o.sp, o.tr, o.trace);
  }
};
struct SkinnyCache {
  std::mutex mu;  // engines may be driven from several host threads
  std::map<SkinnyKey, SkinnyLaunch*> recs;
  int device = -1;
};
SkinnyCache* skinny_cache_create() { return new SkinnyCache(); }
void skinny_cache_destroy(SkinnyCache* c) {
  if (!c) return;
  delete c;  // (the records stay: see alloc_descriptor_record)
}
static SkinnyCache* default_cache() {
  static SkinnyCache* per_dev[64] = {nullptr};
  static std::mutex mu;
  int dev = 0;
  B200W_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (!per_dev[dev & 63]) per_dev[dev & 63] = new SkinnyCache();
  return per_dev[dev & 63];
}

static long long* g_sk_trace = nullptr;
void skinny_set_trace(void* buf) { g_sk_trace = static_cast<long long*>(buf); }

int skinny_suggest_splits(int N, int K, int max_splits, int tile_rows) {
  const int n_tiles = cdiv(N, tile_rows), kb = K / SK_BK;
  int s = num_sms() / n_tiles;       // one wave
  if (s > kb / 2) s = kb / 2;        // at least two k-blocks per CTA
  if (s > max_splits) s = max_splits;
  return s < 1 ? 1 : s;
}

void skinny_gemm(const SkinnyDesc& d, cudaStream_t stream) {
  B200W_REQUIRE(d.X && d.W && d.out, "null operand");
  B200W_REQUIRE(d.M >= 1 && d.M <= 256, "skinny_gemm handles 1..256 activation rows");
  B200W_REQUIRE(d.K % SK_BK == 0 && d.K > 0, "K must be a positive multiple of 64");
  const int splits = d.splits > 1 ? d.splits : 1;
  B200W_REQUIRE(splits == 1 || (d.out_fp32 && !d.bias && d.act == 0 && d.split_stride > 0),
                "split-K needs a plain fp32 partial-sum output");
  B200W_REQUIRE(d.K / SK_BK >= splits, "more splits than k-blocks");
  SkinnyParams p{};
  p.M = d.M;
  p.N = d.N;
  p.Mpad = (d.M + 15) & ~15;
  p.total_kb = d.K / SK_BK;
  p.splits = splits;
  p.tile_rows = d.tile_rows == 64 ? 64 : 128;
  p.n_tiles = cdiv(d.N, p.tile_rows);
  // the MMA always reads a 128-row A tile: with 64-row tiles its second half overlaps the activation tile
  B200W_REQUIRE(p.tile_rows == 128 || p.Mpad * 128 >= 8192, "64-row tiles need at least 64 activation rows");
  const int stage_bytes = p.tile_rows * 128 + p.Mpad * 128;
  p.stages = (200 * 1024) / stage_bytes;
  const int max_kb = cdiv(p.total_kb, splits);
  if (p.stages > max_kb) p.stages = max_kb;
  if (p.stages > SK_MAX_STAGES) p.stages = SK_MAX_STAGES;
  p.bias = d.bias;
  p.out = d.out;
  p.ldo = d.ldo;
  p.split_stride = d.split_stride;
  p.trace = g_sk_trace;
  CUtensorMap tmW, tmX;
  make_tmap_bf16_3d(&tmW, d.W, d.K, d.N, 1, d.ldw, (uint64_t)d.N * d.ldw, SK_BK, p.tile_rows);
  make_tmap_bf16_3d(&tmX, d.X, d.K, d.M, 1, d.ldx, (uint64_t)d.M * d.ldx, SK_BK, p.Mpad);
  // launch record: one per distinct (operands, shape, epilogue), reused by every later identical call
  const SkinnyKey key{d.X, d.W, d.bias, d.out, d.ldx, d.ldw, d.ldo, d.split_stride, d.M, d.N, d.K, d.act, d.out_fp32, splits, p.tile_rows, g_sk_trace};
  SkinnyCache* cache = d.cache ? d.cache : default_cache();
  std::lock_guard<std::mutex> lock(cache->mu);
  int dev = 0;
  B200W_CHECK_CUDA(cudaGetDevice(&dev));
  if (cache->device < 0) cache->device = dev;
  B200W_REQUIRE(cache->device == dev, "skinny_gemm: launch-record cache used from two devices");
  auto it = cache->recs.find(key);
  if (it == cache->recs.end()) {
    if (!d.cache && cache->recs.size() >= 512) {  // the shared default is bounded: drop everything once idle
      cache->recs.clear();  // (the records themselves are never freed nor reused: see alloc_descriptor_record)
    }
    SkinnyLaunch host;
    host.tmW = tmW;
    host.tmX = tmX;
    host.p = p;
    SkinnyLaunch* devrec = static_cast<SkinnyLaunch*>(alloc_descriptor_record(sizeof(SkinnyLaunch)));
    B200W_CHECK_CUDA(cudaMemcpy(devrec, &host, sizeof(host), cudaMemcpyHostToDevice));
    it = cache->recs.emplace(key, devrec).first;
  }
  const SkinnyLaunch* rec = it->second;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  auto launch = [&](auto kernel) {
    ensure_dynamic_smem(kernel, smem);
    launch_pdl(kernel, dim3(p.n_tiles * splits), dim3(SK_THREADS), smem, stream, rec);
  };
  if (d.out_fp32) {
    B200W_REQUIRE(d.act == 0, "GELU is only built for bf16 outputs");
    launch(skinny_gemm_kernel<0, 1>);
  } else if (d.act == 1) {
    launch(skinny_gemm_kernel<1, 0>);
  } else {
    launch(skinny_gemm_kernel<0, 0>);
  }
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

}  // namespace b200w

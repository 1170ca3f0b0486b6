// Attention kernels.
//  * enc_attn_kernel: non-causal flash-style attention (head 64) on tcgen05: S = Q K^T and O += P V with
//    fp32 accumulators in TMEM; Q/K/V tiles arrive by TMA (128B swizzle); the online softmax runs in
//    registers (one query row per thread, exp2 domain) and hands P back through shared memory; three
//    CTAs per SM overlap each other's MMA and softmax phases.  Two modes: encoder self-attention
//    (T = 1500, operands = column ranges of the fused QKV activation) and cross attention for passes
//    with many query rows per stream (prompts, the alignment pass): Q = activation rows of a segment,
//    K/V = the per-chunk cross K/V blocks.
//  * cross_attn_sp_kernel: decode-step cross attention, HBM-bound: K/V of every live (stream, head)
//    stream once through a TMA ring, tcgen05 for both products, single-pass softmax.
//  * cross_attn_kernel / cross_attn_split_kernel: CUDA-core cross attention; the first exports the
//    probabilities of the alignment heads, the second is the fallback of the tensor-core kernel.
//  * dec_attn_kernel: decoder self-attention over the token-granular paged KV cache: one CTA per
//    (query, 4 heads), page-table row staged in shared memory, value rows prefetched to L2 during the
//    key pass.  Latency-bound CUDA-core work.
#include "attention.h"
#include <atomic>
#include <deque>
#include <mutex>

#include "gemm.h"
#include "ptx.cuh"

namespace b200w {

static constexpr int QT = 128;   // queries per CTA
static constexpr int HD = 64;    // head dim
static constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB: a [128][64] bf16 tile

struct EncAttnParams {
  int T, d;
  float scale_log2e;
  bf16* out;  // [B][T][d]
  // cross-attention mode (seg_start != null): grid.z walks segments = runs of query rows of one stream;
  // Q rows come from a [rows][d] activation, K/V from the per-chunk cross K/V blocks [block][T][64]
  const int* seg_start;
  const int* seg_count;
  const int* row_stream;
  int blocks_per_stream, layer_block0, n_heads;
  int dbg = 0;  // B200W_ENC_ATTN_DBG: bit1 one CTA records a phase trace
};

// KT = keys per inner block (64 or 128).  Smaller blocks mean less shared memory and TMEM per CTA,
// so more CTAs share an SM and hide each other's MMA / barrier latencies.
template <int KT, int CTAS>
__global__ void __launch_bounds__(128, CTAS)
enc_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const EncAttnParams p) {
  constexpr int KV_BYTES = KT * HD * 2;          // one K or V block
  constexpr int P_BYTES = 128 * KT * 2;          // P tile: KT/64 halves of [128][64]
  constexpr uint32_t TMEM_COLS = KT == 64 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sP = sV + KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const uint32_t barQ = smem_u32(bars), barK = smem_u32(bars + 1), barV = smem_u32(bars + 2),
                 barS = smem_u32(bars + 3), barO = smem_u32(bars + 4);
  // cross mode: this CTA's query rows are [row0 + q0, row0 + n_rows) of the activation
  const bool cross = p.seg_start != nullptr;
  int row0 = 0, n_rows = p.T;
  if (cross) {
    pdl_trigger();
    pdl_wait();
    row0 = p.seg_start[b];
    n_rows = p.seg_count[b];
    if (q0 >= n_rows) return;  // before any barrier / TMEM use: whole CTA leaves
  }

  if (tid == 0) {
    mbar_init(barQ, 1);
    mbar_init(barK, 1);
    mbar_init(barV, 1);
    mbar_init(barS, 1);
    mbar_init(barO, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + KT;
  const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;

  const int nblk = (p.T + KT - 1) / KT;
  // self (encoder): Q, K, V are column ranges of the fused QKV activation of window b.
  // cross: Q = activation rows, K and V = [T][64] blocks (stream, layer, k|v, head) of the cross K/V buffer.
  const int colQ = h * HD;
  const int colK = cross ? 0 : p.d + h * HD, colV = cross ? 0 : 2 * p.d + h * HD;
  const int rowQ = row0 + q0, zQ = cross ? 0 : b;
  int zK = b, zV = b;
  if (cross) {
    zK = p.row_stream[row0] * p.blocks_per_stream + p.layer_block0 + h;
    zV = zK + p.n_heads;
  }

  if (tid == 0) {
    mbar_arrive_expect_tx(barQ, TILE_BYTES);
    tma_load_3d(smem_u32(sQ), &tmQ, barQ, colQ, rowQ, zQ);
    mbar_arrive_expect_tx(barK, KV_BYTES);
    tma_load_3d(smem_u32(sK), &tmKV, barK, colK, 0, zK);
    mbar_arrive_expect_tx(barV, KV_BYTES);
    tma_load_3d(smem_u32(sV), &tmKV, barV, colV, 0, zV);
  }

  constexpr uint32_t idescS = make_idesc_bf16(128, KT, 0, 0);  // Q (K-major) x K (K-major)
  constexpr uint32_t idescO = make_idesc_bf16(128, 64, 0, 1);  // P (K-major) x V (MN-major)

  float o[HD];
#pragma unroll
  for (int i = 0; i < HD; ++i) o[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  const int r = tid;  // query row inside the tile == TMEM lane

  for (int j = 0; j < nblk; ++j) {
    const uint32_t ph = j & 1;
    if (tid == 0 && j == 0) {  // S of block 0; S of block j+1 is issued behind the P.V of block j (below)
      mbar_wait(barQ, 0);
      mbar_wait(barK, 0);
      tc_fence_after();
      const uint64_t qd = make_sw128_desc(smem_u32(sQ));
      const uint64_t kd = make_sw128_desc(smem_u32(sK));
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tmem_S, qd + 2 * k, kd + 2 * k, idescS, k != 0);
      tc_commit(barS);
    }
    mbar_wait(barS, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < nblk) {  // K tile is free again: prefetch the next one
      mbar_arrive_expect_tx(barK, KV_BYTES);
      tma_load_3d(smem_u32(sK), &tmKV, barK, colK, (j + 1) * KT, zK);
    }
    const int kv_valid = min(KT, p.T - j * KT);
    // one pass over the 64 scores of this block, held in registers:
    // max -> exp2 (ex2.approx) -> bf16 pack into the swizzled K-major P tile
    static_assert(KT == 64, "the single-pass softmax keeps one 64-key block in registers");
    float alpha, lsum = 0.f;
    {
      uint32_t v[64];
      tmem_ld32(tmem_S + lane_sel, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(tmem_S + lane_sel + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      if (kv_valid < KT) {  // ragged tail of the last block (uniform branch)
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kv_valid) v[i] = 0xff800000u;  // -inf
      }
      float mx = __uint_as_float(v[0]);
#pragma unroll
      for (int i = 1; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      const float m_new = fmaxf(m_run, mx * p.scale_log2e);
      alpha = ex2_approx(m_run - m_new);  // first block: exp2(-inf) = 0
      m_run = m_new;
      uint32_t packed[32];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2e, -m_new));
        const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2e, -m_new));
        lsum += p0 + p1;
        __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
        packed[i >> 1] = *reinterpret_cast<uint32_t*>(&pb);
      }
      uint8_t* row = sP + r * 128;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc)  // eight 16-byte chunks (8 keys each)
        *reinterpret_cast<uint4*>(row + ((cc ^ (r & 7)) << 4)) =
            make_uint4(packed[4 * cc], packed[4 * cc + 1], packed[4 * cc + 2], packed[4 * cc + 3]);
    }
    l_run = l_run * alpha + lsum;
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] *= alpha;
    fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the MMA's async proxy
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(barV, ph);
      tc_fence_after();
      const uint64_t vd = make_sw128_desc(smem_u32(sV));
#pragma unroll
      for (int k = 0; k < KT / 16; ++k) {
        const uint64_t pd = make_sw128_desc(smem_u32(sP + (k >> 2) * TILE_BYTES)) + 2 * (k & 3);
        tc_mma_bf16(tmem_O, pd, vd + 128 * k, idescO, k != 0);  // V: +16 key rows = +2048 B
      }
      tc_commit(barO);
      if (j + 1 < nblk) {
        // software pipeline: every thread is done with S of this block (barrier above), so the next block's
        // S = Q K^T runs on the tensor pipe while the threads wait for / accumulate this block's O
        mbar_wait(barK, ph ^ 1);
        tc_fence_after();
        const uint64_t qd = make_sw128_desc(smem_u32(sQ));
        const uint64_t kd = make_sw128_desc(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tmem_S, qd + 2 * k, kd + 2 * k, idescS, k != 0);
        tc_commit(barS);
      }
    }
    mbar_wait(barO, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < nblk) {
      mbar_arrive_expect_tx(barV, KV_BYTES);
      tma_load_3d(smem_u32(sV), &tmKV, barV, colV, (j + 1) * KT, zV);
    }
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_O + lane_sel + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c0 + i] += __uint_as_float(v[i]);
    }
    tc_fence_before();  // next iteration's MMAs overwrite S / O after the CTA-wide sync below
    __syncthreads();
  }

  const int t = q0 + r;
  if (t < n_rows) {
    const float inv = 1.0f / l_run;
    bf16* dst = p.out + (cross ? (long long)(row0 + t) : (long long)b * p.T + t) * p.d + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) {
      uint4 u;
      __nv_bfloat162* u2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) u2[i] = __floats2bfloat162_rn(o[c + 2 * i] * inv, o[c + 2 * i + 1] * inv);
      *reinterpret_cast<uint4*>(dst + c) = u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- Second flash kernel (encoder self-attention and the flash cross attention), round 2.
// The kernel above is one warpgroup doing everything in lock step: per 64-key block two CTA-wide barriers, the O
// tile read back from TMEM and rescaled in registers (a third of its instruction stream), MMAs issued by a softmax
// thread.  ncu: tensor pipe 21.5 % active, issue slots 47.5 % busy; 783 us at the bench shape.  This one:
//  * warp-specialised: warps 0-3 softmax (thread = query row = TMEM lane), warp 4 TMA producer (4-stage K and V
//    rings), warp 5 MMA issuer; everything hands over through mbarriers, no CTA-wide barrier inside the loop.
//    The issuer runs its loop with the WHOLE warp and elects a lane only around the tcgen05 instructions, so that
//    descriptors live in uniform registers (a single divergent lane spent ~70 cycles per MMA on R2UR chains);
//  * S double-buffered in TMEM: S(j+1) = Q K_{j+1}^T runs on the tensor pipe while the softmax of block j runs;
//    issue order S(0) S(1) PV(0) S(2) PV(1) ...;
//  * O stays in TMEM for the whole key loop (the P.V MMAs accumulate) and is rescaled LAZILY: probabilities are
//    taken relative to a reference maximum that only moves when the running maximum exceeds it by more than 2^8
//    (exact arithmetic: l and O use the same reference; the result is O / l whatever the reference was);
//  * the A operands of both products live in TENSOR MEMORY.  With every operand in shared memory (first versions of
//    this kernel: 660-700 us whatever the softmax did, phase trace: ~70 cycles per 128x64x16 MMA instead of 32) the
//    kernel is shared-memory-bandwidth bound: per 64-key block the MMAs fetch Q (16 KB, again every block) + K (8 KB)
//    and P (16 KB) + V (8 KB), the softmax stores P (16 KB), TMA writes K and V (16 KB): 80 KB at 128 B / clock =
//    640 cycles per block against 262 cycles of tensor work.  Q is copied once into TMEM (bf16 pairs, 32 columns)
//    and P is written by tcgen05.st over the first 32 columns of the S buffer it was computed from; the MMAs read
//    only K and V from shared memory (16 KB per block);
//  * row maximum and row sum as four independent chains (FMNMX3, FADD2), scale-and-shift as FFMA2;
//  * what is left is the exponential: 64 ex2 per row per block = 512 MUFU cycles per CTA-block against 262 tensor
//    cycles (head dim 64), so one exponential in PE is evaluated on the FMA pipe (ex2_poly) instead.
// 2 CTAs per SM (256 TMEM columns, 82 KB shared memory each).  Measured (tools/enc_attn_bench.py, 32 x 1500 x 20
// heads): 783 us -> 495 us (744 TFLOP/s).  Also tried and dropped: two query tiles per CTA sharing the K / V stream
// (lock-step softmax phases, 770 us with operands in shared memory) and two threads per query row (no gain).
// Phase trace of one CTA (B200W_ENC_ATTN_DBG=2): clock64 stamps, [role][block][stamp]
__device__ long long g_enc_trace[4 * 32 * 8];
template <int PE>  // every PE-th exponential on the FMA pipe (0 = none)
__global__ void __launch_bounds__(192, 2)
enc_attn4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const EncAttnParams p) {
  constexpr int KT = 64, NS = 4;
  constexpr int KV_BYTES = KT * HD * 2;  // one K or V block (8 KB)
  constexpr uint32_t TMEM_COLS = 256;    // S0 / P0 | S1 / P1 | O | Q (32 columns of bf16 pairs)
  constexpr float TAU = 8.0f;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;        // [NS]
  uint8_t* sV = sK + NS * KV_BYTES;     // [NS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NS * KV_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 4 * NS + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const uint32_t barQ = smem_u32(bars), barQT = smem_u32(bars + 1);         // Q in shared memory / in tensor memory
  auto barK = [&](int s) { return smem_u32(bars + 2 + s); };                // K block landed
  auto barV = [&](int s) { return smem_u32(bars + 2 + NS + s); };           // V block landed
  auto barKfree = [&](int s) { return smem_u32(bars + 2 + 2 * NS + s); };   // S MMAs of the block done
  auto barVfree = [&](int s) { return smem_u32(bars + 2 + 3 * NS + s); };   // P.V MMAs of the block done
  auto barS = [&](int s) { return smem_u32(bars + 2 + 4 * NS + s); };       // S = Q K^T complete
  auto barP = [&](int s) { return smem_u32(bars + 2 + 4 * NS + 2 + s); };   // P written by all 128 rows
  auto barO = [&](int s) { return smem_u32(bars + 2 + 4 * NS + 4 + s); };   // O += P V complete
  const bool cross = p.seg_start != nullptr;
  int row0 = 0, n_rows = p.T;
  if (cross) {
    pdl_trigger();
    pdl_wait();
    row0 = p.seg_start[b];
    n_rows = p.seg_count[b];
    if (q0 >= n_rows) return;  // before any barrier / TMEM use: whole CTA leaves
  }
  if (tid == 0) {
    mbar_init(barQ, 1);
    mbar_init(barQT, 128);
    for (int s = 0; s < NS; ++s) {
      mbar_init(barK(s), 1);
      mbar_init(barV(s), 1);
      mbar_init(barKfree(s), 1);
      mbar_init(barVfree(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(barS(s), 1);
      mbar_init(barP(s), 128);
      mbar_init(barO(s), 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 2 * KT, tmem_Q = tmem_base + 3 * KT;
  const int nblk = (p.T + KT - 1) / KT;
  const bool tracing = (p.dbg & 2) && blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 1;
  auto stamp = [&](int role, int j, int k) {
    if (tracing && j < 32 && (role < 2 || lane == 0)) g_enc_trace[(role * 32 + j) * 8 + k] = clock64();
  };

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int colQ = h * HD;
      const int colK = cross ? 0 : p.d + h * HD, colV = cross ? 0 : 2 * p.d + h * HD;
      const int rowQ = row0 + q0, zQ = cross ? 0 : b;
      int zK = b, zV = b;
      if (cross) {
        zK = p.row_stream[row0] * p.blocks_per_stream + p.layer_block0 + h;
        zV = zK + p.n_heads;
      }
      mbar_arrive_expect_tx(barQ, TILE_BYTES);
      tma_load_3d(smem_u32(sQ), &tmQ, barQ, colQ, rowQ, zQ);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % NS, n = j / NS;
        if (j >= NS) mbar_wait_backoff(barKfree(s), (n - 1) & 1);
        mbar_arrive_expect_tx(barK(s), KV_BYTES);
        tma_load_3d(smem_u32(sK + s * KV_BYTES), &tmKV, barK(s), colK, j * KT, zK);
        if (j >= NS) mbar_wait_backoff(barVfree(s), (n - 1) & 1);
        mbar_arrive_expect_tx(barV(s), KV_BYTES);
        tma_load_3d(smem_u32(sV + s * KV_BYTES), &tmKV, barV(s), colV, j * KT, zV);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, elected lane issues)
    constexpr uint32_t idescS = make_idesc_bf16(128, KT, 0, 0);  // Q (TMEM) x K (K-major)
    constexpr uint32_t idescO = make_idesc_bf16(128, 64, 0, 1);  // P (TMEM) x V (MN-major)
    auto pv = [&](int i) {  // O += P(i) V(i)
      const int sb = i & 1, nb = i >> 1, s = i % NS, n = i / NS;
      mbar_wait_backoff(barV(s), n & 1);
      stamp(2, i, 3);
      mbar_wait_backoff(barP(sb), nb & 1);
      stamp(2, i, 4);
      tc_fence_after();
      const uint64_t vd = make_sw128_desc(smem_u32(sV + s * KV_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KT / 16; ++k)  // P: 16 keys = 8 columns; V: +16 key rows = +2048 B
          tc_mma_bf16_ta(tmem_O, tmem_base + sb * KT + 8 * k, vd + 128 * k, idescO, (i > 0 || k > 0) ? 1u : 0u);
        tc_commit(barO(sb));
        tc_commit(barVfree(s));
      }
      __syncwarp();
      stamp(2, i, 6);
    };
    mbar_wait_backoff(barQT, 0);
    tc_fence_after();
    for (int j = 0; j < nblk; ++j) {
      const int sb = j & 1, s = j % NS, n = j / NS;
      stamp(2, j, 0);
      mbar_wait_backoff(barK(s), n & 1);
      stamp(2, j, 1);
      tc_fence_after();
      // (S buffer sb holds P(j-2) until PV(j-2) has read it: that MMA was issued before this one, and MMAs execute
      // in issue order)
      const uint64_t kd = make_sw128_desc(smem_u32(sK + s * KV_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16_ta(tmem_base + sb * KT, tmem_Q + 8 * k, kd + 2 * k, idescS, k != 0);
        tc_commit(barS(sb));
        tc_commit(barKfree(s));
      }
      __syncwarp();
      stamp(2, j, 2);
      if (j > 0) pv(j - 1);
    }
    pv(nblk - 1);
  } else {
    // ------------------------------------------------------------------ softmax: thread = query row = TMEM lane
    const int r = tid;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    {  // this row of Q: swizzled shared-memory tile -> 32 columns of bf16 pairs in tensor memory
      mbar_wait(barQ, 0);
      uint32_t qv[32];
      const uint8_t* row = sQ + r * 128;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + ((cc ^ (r & 7)) << 4));
        qv[4 * cc] = u.x; qv[4 * cc + 1] = u.y; qv[4 * cc + 2] = u.z; qv[4 * cc + 3] = u.w;
      }
      tmem_st32(tmem_Q + lane_sel, qv);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(barQT);
    }
    float m_ref = -INFINITY, l_run = 0.f;  // reference maximum (log2 domain) of every probability so far
    for (int j = 0; j < nblk; ++j) {
      const int s = j & 1, n = j >> 1;
      if (r == 0) stamp(0, j, 0);
      mbar_wait(barS(s), n & 1);
      if (r == 0) stamp(0, j, 1);
      tc_fence_after();
      uint32_t v[64];
      tmem_ld32(tmem_base + s * KT + lane_sel, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(tmem_base + s * KT + lane_sel + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      if (r == 0) stamp(0, j, 2);
      const int kv_valid = min(KT, p.T - j * KT);
      if (kv_valid < KT) {  // ragged tail of the last block (uniform branch)
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kv_valid) v[i] = 0xff800000u;  // -inf
      }
      float mxc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int o = 16 * c;
        float m = fmax3(__uint_as_float(v[o]), __uint_as_float(v[o + 1]), __uint_as_float(v[o + 2]));
#pragma unroll
        for (int i = 3; i + 1 < 16; i += 2) m = fmax3(m, __uint_as_float(v[o + i]), __uint_as_float(v[o + i + 1]));
        mxc[c] = fmaxf(m, __uint_as_float(v[o + 15]));
      }
      const float mxs = fmaxf(fmax3(mxc[0], mxc[1], mxc[2]), mxc[3]) * p.scale_log2e;
      if (j == 0) {
        m_ref = mxs;
      } else {
        const bool need = mxs > m_ref + TAU;
        if (__any_sync(0xffffffffu, need)) {  // rare after the first blocks: rescale this warp's rows of O
          mbar_wait(barO((j - 1) & 1), ((j - 1) >> 1) & 1);  // every P.V issued so far is complete
          tc_fence_after();
          const float alpha = need ? ex2_approx(m_ref - mxs) : 1.0f;
#pragma unroll
          for (int c0 = 0; c0 < HD; c0 += 32) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_sel + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem_O + lane_sel + c0, o);
          }
          tmem_st_wait();
          l_run *= alpha;
          if (need) m_ref = mxs;
        }
      }
      uint32_t packed[32];
      {
        const unsigned long long sc2 = pack_f32x2(p.scale_log2e, p.scale_log2e), nm2 = pack_f32x2(-m_ref, -m_ref);
        unsigned long long ls2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          float x0, x1;
          unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2), x0, x1);
          const float p0 = (PE > 0 && (i % (PE > 0 ? PE : 1)) == 0) ? ex2_poly(x0) : ex2_approx(x0);
          const float p1 = ex2_approx(x1);
          ls2[(i >> 1) & 3] = fadd2(ls2[(i >> 1) & 3], pack_f32x2(p0, p1));
          __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
          packed[i >> 1] = *reinterpret_cast<uint32_t*>(&pb);
        }
        float a, b2;
        unpack_f32x2(fadd2(fadd2(ls2[0], ls2[1]), fadd2(ls2[2], ls2[3])), a, b2);
        l_run += a + b2;
      }
      if (r == 0) stamp(0, j, 3);
      // P over the first 32 columns of this block's S buffer (every S value of the row is in registers by now)
      tmem_st32(tmem_base + s * KT + lane_sel, packed);
      tmem_st_wait();
      if (r == 0) stamp(0, j, 4);
      tc_fence_before();
      mbar_arrive(barP(s));
      if (r == 0) stamp(0, j, 5);
    }
    mbar_wait(barO((nblk - 1) & 1), ((nblk - 1) >> 1) & 1);
    tc_fence_after();
    const int t = q0 + r;
    const float inv = 1.0f / l_run;
    bf16* dst = p.out + (cross ? (long long)(row0 + t) : (long long)b * p.T + t) * p.d + h * HD;
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_sel + c0, o);
      tmem_ld_wait();
      if (t < n_rows) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 u;
          __nv_bfloat162* u2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            u2[i] = __floats2bfloat162_rn(__uint_as_float(o[c + 2 * i]) * inv, __uint_as_float(o[c + 2 * i + 1]) * inv);
          *reinterpret_cast<uint4*>(dst + c0 + c) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}
static constexpr int ENC4_SMEM = TILE_BYTES + 2 * 4 * (64 * HD * 2) + (2 + 4 * 4 + 6) * 8 + 16 + 1024;

static int enc_attn_variant() {  // 0: first kernel (B200W_ENC_ATTN_V1=1), 4: the kernel above
  static const int v = getenv("B200W_ENC_ATTN_V1") ? 0 : 4;
  return v;
}
void enc_attention_trace_read(long long* host_out, int n) {
  B200W_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_enc_trace, sizeof(long long) * std::min(n, 4 * 32 * 8)));
}
static int enc_attn_dbg() {
  static const int v = getenv("B200W_ENC_ATTN_DBG") ? atoi(getenv("B200W_ENC_ATTN_DBG")) : 0;
  return v;
}
static int enc_attn_pe() {
  static const int v = getenv("B200W_ENC_ATTN_PE") ? atoi(getenv("B200W_ENC_ATTN_PE")) : 0;
  return v;
}
template <typename... Args>
static void launch_enc_attn4(bool pdl, int q_rows, int gy, int gz, cudaStream_t stream, Args&&... args) {
  const dim3 grid(cdiv(q_rows, QT), gy, gz);
  auto go = [&](auto kernel) {
    ensure_dynamic_smem(kernel, ENC4_SMEM);
    if (pdl) launch_pdl(kernel, grid, dim3(192), (size_t)ENC4_SMEM, stream, args...);
    else kernel<<<grid, 192, ENC4_SMEM, stream>>>(args...);
  };
  switch (enc_attn_pe()) {
    case 4: go(enc_attn4_kernel<4>); break;
    case 8: go(enc_attn4_kernel<8>); break;
    default: go(enc_attn4_kernel<0>); break;
  }
}

// ---- optional per-launch timing of the encoder attention (bench.py's second tensor roofline): CUDA events around
// every launch on the launching stream, read back after a synchronize.
struct AttnProfSlot {
  cudaEvent_t beg, end;
  double flops;
};
static std::atomic<int> g_attn_prof_on{0};
static std::deque<AttnProfSlot> g_attn_prof_slots;  // guarded by g_attn_prof_mu
static size_t g_attn_prof_used = 0;
static std::mutex g_attn_prof_mu;
void enc_attention_profile_begin() {
  std::lock_guard<std::mutex> lock(g_attn_prof_mu);
  g_attn_prof_on = 1;
  g_attn_prof_used = 0;
}
void enc_attention_profile_end(double* total_ms, double* total_flops, long long* launches) {
  std::lock_guard<std::mutex> lock(g_attn_prof_mu);
  g_attn_prof_on = 0;
  double ms = 0, fl = 0;
  for (size_t i = 0; i < g_attn_prof_used; ++i) {
    float t = 0.f;
    B200W_CHECK_CUDA(cudaEventSynchronize(g_attn_prof_slots[i].end));
    B200W_CHECK_CUDA(cudaEventElapsedTime(&t, g_attn_prof_slots[i].beg, g_attn_prof_slots[i].end));
    ms += t;
    fl += g_attn_prof_slots[i].flops;
  }
  *total_ms = ms;
  *total_flops = fl;
  *launches = (long long)g_attn_prof_used;
}
static AttnProfSlot* attn_prof_slot(double flops) {
  std::lock_guard<std::mutex> lock(g_attn_prof_mu);
  if (g_attn_prof_used == g_attn_prof_slots.size()) {
    AttnProfSlot s;
    B200W_CHECK_CUDA(cudaEventCreate(&s.beg));
    B200W_CHECK_CUDA(cudaEventCreate(&s.end));
    g_attn_prof_slots.push_back(s);
  }
  AttnProfSlot* s = &g_attn_prof_slots[g_attn_prof_used++];
  s->flops = flops;
  return s;
}

template <int KT, int CTAS>
static void launch_enc_attn(const bf16* qkv, int B, int T, int d, int n_heads, bf16* out, cudaStream_t stream) {
  CUtensorMap tmQ, tmKV;
  make_tmap_bf16_3d(&tmQ, qkv, 3 * (uint64_t)d, T, B, 3 * (uint64_t)d, (uint64_t)T * 3 * d, HD, 128);
  make_tmap_bf16_3d(&tmKV, qkv, 3 * (uint64_t)d, T, B, 3 * (uint64_t)d, (uint64_t)T * 3 * d, HD, KT);
  constexpr int smem_bytes = TILE_BYTES + 2 * KT * HD * 2 + 128 * KT * 2 + 5 * 8 + 16 + 1024;
  ensure_dynamic_smem(enc_attn_kernel<KT, CTAS>, smem_bytes);
  EncAttnParams p{T, d, 0.125f * 1.4426950408889634f, out, nullptr, nullptr, nullptr, 0, 0, n_heads};
  AttnProfSlot* ps = nullptr;
  if (g_attn_prof_on.load(std::memory_order_relaxed)) {
    ps = attn_prof_slot(4.0 * (double)T * (double)T * (double)d * (double)B);  // S = Q K^T and O = P V: 2 x 2 T^2 d
    B200W_CHECK_CUDA(cudaEventRecord(ps->beg, stream));
  }
  if (enc_attn_variant() >= 2) {
    make_tmap_bf16_3d(&tmKV, qkv, 3 * (uint64_t)d, T, B, 3 * (uint64_t)d, (uint64_t)T * 3 * d, HD, 64);
    p.dbg = enc_attn_dbg();
    launch_enc_attn4(false, T, n_heads, B, stream, tmQ, tmKV, p);
  } else {
    enc_attn_kernel<KT, CTAS><<<dim3(cdiv(T, QT), n_heads, B), 128, smem_bytes, stream>>>(tmQ, tmKV, p);
  }
  if (ps) B200W_CHECK_CUDA(cudaEventRecord(ps->end, stream));
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

void enc_attention(const bf16* qkv, int B, int T, int d, int n_heads, bf16* out, cudaStream_t stream) {
  B200W_REQUIRE(d == n_heads * HD, "head dim must be 64");
  if (B <= 0) return;
  launch_enc_attn<64, 3>(qkv, B, T, d, n_heads, out, stream);
}

// Cross attention for passes with many query rows per stream (prompts, the teacher-forced alignment pass):
// the same flash kernel, one CTA per (128 query rows of a segment, head): every segment streams its K/V
// once instead of once per group of 8 rows.
void cross_attention_flash(const bf16* cross_kv, int n_streams, int n_layers, int layer, int n_heads, int len,
                           const bf16* q, int n_q_rows, const int* row_stream, const int* seg_start,
                           const int* seg_count, int n_segs, int max_seg_rows, bf16* out, cudaStream_t stream) {
  if (n_segs <= 0) return;
  const int d = n_heads * HD;
  CUtensorMap tmQ, tmKV;
  make_tmap_bf16_3d(&tmQ, q, d, n_q_rows, 1, d, (uint64_t)n_q_rows * d, HD, 128);
  make_tmap_bf16_3d(&tmKV, cross_kv, HD, len, (uint64_t)n_streams * n_layers * 2 * n_heads, HD, (uint64_t)len * HD, HD, 64);
  constexpr int smem_bytes = TILE_BYTES + 2 * 64 * HD * 2 + 128 * 64 * 2 + 5 * 8 + 16 + 1024;
  ensure_dynamic_smem(enc_attn_kernel<64, 3>, smem_bytes);
  EncAttnParams p{len, d, 0.125f * 1.4426950408889634f, out, seg_start, seg_count, row_stream,
                  n_layers * 2 * n_heads, 2 * layer * n_heads, n_heads};
  if (enc_attn_variant() >= 2) {
    p.dbg = enc_attn_dbg();
    launch_enc_attn4(true, max_seg_rows, n_heads, n_segs, stream, tmQ, tmKV, p);
  } else
    launch_pdl(enc_attn_kernel<64, 3>, dim3(cdiv(max_seg_rows, QT), n_heads, n_segs), dim3(128), (size_t)smem_bytes, stream,
               tmQ, tmKV, p);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------ decoder attention
// One CTA per (query, group of NW heads), one warp per head.  The query's page-table row is staged
// in shared memory once (all heads of a query share it), so the K and V row loads do not wait on a
// table load each: a warp keeps one 128-byte K row per lane, then 16 V rows, in flight at a time (two K rows per lane /
// 32 V rows measured slower: more registers, fewer resident CTAs).
template <int NW>
__global__ void __launch_bounds__(NW * 32) dec_attn_kernel(const DecAttnParams p) {
  extern __shared__ float s_dyn[];  // [max_len] table (int) | [NW][max_len] scores
  pdl_trigger();
  pdl_wait();
  const int qi = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y * NW + warp;
  int* s_tab = reinterpret_cast<int*>(s_dyn);
  float* sc = s_dyn + p.max_len + warp * p.max_len;
  const int len = p.kvlen[qi];
  const int tr = p.trow[qi];
  const long long base = p.kv_base_off ? p.kv_base_off[tr] : 0;
  const bf16* kb = p.kbase + base + h * HD;
  const bf16* vb = p.vbase + base + h * HD;
  {
    const int* tab = p.table ? p.table + (long long)tr * p.table_stride : nullptr;
    for (int i = threadIdx.x; i < len; i += NW * 32) s_tab[i] = tab ? tab[i] : i;
  }
  // q in registers (every lane holds the whole 64-vector, pre-scaled)
  float q[HD];
  {
    const bf16* qp = p.q + (long long)qi * p.q_stride + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(qp + c);
      const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(u2[i]);
        q[c + 2 * i] = f.x * p.scale;
        q[c + 2 * i + 1] = f.y * p.scale;
      }
    }
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int i = lane; i < len; i += 32) {
    const long long roff = (long long)s_tab[i] * p.kv_row_stride;
    const bf16* kp = kb + roff;
    uint4 kr[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) kr[c] = __ldg(reinterpret_cast<const uint4*>(kp) + c);
    // the value row of the same key (one 128-byte line) is needed right after the softmax: start
    // pulling it into L2 now, so that the P.V loop does not pay HBM latency per batch of rows
    asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + roff));
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&kr[c]);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __bfloat1622float2(u2[t]);
        acc = fmaf(q[8 * c + 2 * t], f.x, acc);
        acc = fmaf(q[8 * c + 2 * t + 1], f.y, acc);
      }
    }
    sc[i] = acc;
    mx = fmaxf(mx, acc);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int i = lane; i < len; i += 32) {
    const float e = __expf(sc[i] - mx);
    sc[i] = e;
    sum += e;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  __syncwarp();

  const int ex = p.head_export ? p.head_export[h] : -1;
  if (ex >= 0) {
    float* dst = p.probs_out + ((long long)ex * p.n_queries + qi) * p.probs_stride;
    for (int i = lane; i < len; i += 32) dst[i] = sc[i] * inv;
  }

  float2 acc = make_float2(0.f, 0.f);
  const bf16* vl = vb + 2 * lane;
  int i = 0;
  for (; i + 16 <= len; i += 16) {  // 16 independent 128-byte row reads in flight per warp
    __nv_bfloat162 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u)
      v[u] = *reinterpret_cast<const __nv_bfloat162*>(vl + (long long)s_tab[i + u] * p.kv_row_stride);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float pw = sc[i + u];
      const float2 f = __bfloat1622float2(v[u]);
      acc.x = fmaf(pw, f.x, acc.x);
      acc.y = fmaf(pw, f.y, acc.y);
    }
  }
  {  // ragged tail: still all loads first
    __nv_bfloat162 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u)
      if (i + u < len) v[u] = *reinterpret_cast<const __nv_bfloat162*>(vl + (long long)s_tab[i + u] * p.kv_row_stride);
#pragma unroll
    for (int u = 0; u < 16; ++u)
      if (i + u < len) {
        const float2 f = __bfloat1622float2(v[u]);
        acc.x = fmaf(sc[i + u], f.x, acc.x);
        acc.y = fmaf(sc[i + u], f.y, acc.y);
      }
  }
  bf16* op = p.out + (long long)qi * p.out_stride + h * HD + 2 * lane;
  *reinterpret_cast<__nv_bfloat162*>(op) = __floats2bfloat162_rn(acc.x * inv, acc.y * inv);
}

// (Round 2 tried to regroup this kernel per (stream, head) -- the beams of a stream in one CTA, key / value rows of
// the shared history loaded once for all beams, scores by one thread per (beam, position), every load of a phase in
// flight at once.  It passed the block tests and lost: 30.4 us per launch against 25.3 us in the ncu launch list of
// a full bench step and 4.33 ms against 4.10 ms per decode step in an A/B on one box (tools/gpu_ab.sh) -- ~11 M
// warp instructions per launch, issue-bound.  This kernel, one warp per (row, head), stays.)
// Decode steps: the QKV projection arrives as split-K partial sums.  This kernel reduces the slabs of ITS
// row and head group (bias, bf16 rounding: the same arithmetic as split_finalize), writes the new key /
// value into the cache, and attends: the row's history comes from the cache through the page table, its
// own position (the last one) from shared memory.  One launch instead of reduction + attention, and q
// never goes through global memory.  A row only needs its own new key / value, so CTAs are independent.
__global__ void __launch_bounds__(128) qkv_attn_kernel(const QkvAttnParams p) {
  extern __shared__ float s_dyn[];  // [max_len] table (int) | [4][max_len] scores
  __shared__ float s_q[256], s_k[256], s_v[256];
  pdl_trigger();
  const int qi = blockIdx.x, hg = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = hg * 256 + 2 * tid;  // this thread's two columns inside each of the q / k / v sections
  const int d = p.d, N = 3 * d;
  float2 bq = __ldg(reinterpret_cast<const float2*>(p.bias + c));
  float2 bk = __ldg(reinterpret_cast<const float2*>(p.bias + d + c));
  float2 bv = __ldg(reinterpret_cast<const float2*>(p.bias + 2 * d + c));
  pdl_wait();
  {
    const float* pp = p.partial + (long long)qi * N + c;
#pragma unroll 4
    for (int s2 = 0; s2 < p.n_splits; ++s2) {
      const float2 a = *reinterpret_cast<const float2*>(pp + s2 * p.split_stride);
      const float2 b = *reinterpret_cast<const float2*>(pp + s2 * p.split_stride + d);
      const float2 e = *reinterpret_cast<const float2*>(pp + s2 * p.split_stride + 2 * d);
      bq.x += a.x; bq.y += a.y; bk.x += b.x; bk.y += b.y; bv.x += e.x; bv.y += e.y;
    }
  }
  const __nv_bfloat162 q2 = __floats2bfloat162_rn(bq.x, bq.y), k2 = __floats2bfloat162_rn(bk.x, bk.y),
                       v2 = __floats2bfloat162_rn(bv.x, bv.y);
  const long long slot = p.slot[qi];
  *reinterpret_cast<__nv_bfloat162*>(p.kpool + slot * d + c) = k2;
  *reinterpret_cast<__nv_bfloat162*>(p.vpool + slot * d + c) = v2;
  {
    const float2 qf = __bfloat1622float2(q2), kf = __bfloat1622float2(k2), vf = __bfloat1622float2(v2);
    s_q[2 * tid] = qf.x * p.scale; s_q[2 * tid + 1] = qf.y * p.scale;
    s_k[2 * tid] = kf.x; s_k[2 * tid + 1] = kf.y;
    s_v[2 * tid] = vf.x; s_v[2 * tid + 1] = vf.y;
  }
  int* s_tab = reinterpret_cast<int*>(s_dyn);
  float* sc = s_dyn + p.max_len + warp * p.max_len;
  const int len = p.kvlen[qi];
  const int hist = len - 1;  // positions [0, hist) come from the cache, position hist is this row's own token
  const int tr = p.trow[qi];
  const int h = hg * 4 + warp;
  const bf16* kb = p.kpool + h * HD;
  const bf16* vb = p.vpool + h * HD;
  {
    const int* tab = p.table + (long long)tr * p.table_stride;
    for (int i = tid; i < hist; i += 128) s_tab[i] = tab[i];
  }
  __syncthreads();
  float q[HD];
#pragma unroll
  for (int i = 0; i < HD; ++i) q[i] = s_q[warp * HD + i];
  float mx = -INFINITY;
  for (int i = lane; i < hist; i += 32) {
    const long long roff = (long long)s_tab[i] * d;
    const bf16* kp = kb + roff;
    uint4 kr[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) kr[cc] = __ldg(reinterpret_cast<const uint4*>(kp) + cc);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + roff));
    float acc = 0.f;
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&kr[cc]);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __bfloat1622float2(u2[t]);
        acc = fmaf(q[8 * cc + 2 * t], f.x, acc);
        acc = fmaf(q[8 * cc + 2 * t + 1], f.y, acc);
      }
    }
    sc[i] = acc;
    mx = fmaxf(mx, acc);
  }
  float own = 0.f;  // score of the row's own position (every lane computes it: smem broadcast reads)
#pragma unroll
  for (int i = 0; i < HD; ++i) own = fmaf(q[i], s_k[warp * HD + i], own);
  mx = fmaxf(mx, own);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int i = lane; i < hist; i += 32) {
    const float e = __expf(sc[i] - mx);
    sc[i] = e;
    sum += e;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float p_own = __expf(own - mx);
  sum += p_own;
  const float inv = 1.0f / sum;
  __syncwarp();
  float2 acc = make_float2(p_own * s_v[warp * HD + 2 * lane], p_own * s_v[warp * HD + 2 * lane + 1]);
  const bf16* vl = vb + 2 * lane;
  int i = 0;
  for (; i + 16 <= hist; i += 16) {
    __nv_bfloat162 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = *reinterpret_cast<const __nv_bfloat162*>(vl + (long long)s_tab[i + u] * d);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float pw = sc[i + u];
      const float2 f = __bfloat1622float2(v[u]);
      acc.x = fmaf(pw, f.x, acc.x);
      acc.y = fmaf(pw, f.y, acc.y);
    }
  }
  {
    __nv_bfloat162 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u)
      if (i + u < hist) v[u] = *reinterpret_cast<const __nv_bfloat162*>(vl + (long long)s_tab[i + u] * d);
#pragma unroll
    for (int u = 0; u < 16; ++u)
      if (i + u < hist) {
        const float2 f = __bfloat1622float2(v[u]);
        acc.x = fmaf(sc[i + u], f.x, acc.x);
        acc.y = fmaf(sc[i + u], f.y, acc.y);
      }
  }
  bf16* op = p.out + (long long)qi * p.out_stride + h * HD + 2 * lane;
  *reinterpret_cast<__nv_bfloat162*>(op) = __floats2bfloat162_rn(acc.x * inv, acc.y * inv);
}

void qkv_attention(const QkvAttnParams& p, int n_rows, int n_heads, cudaStream_t stream) {
  if (n_rows <= 0) return;
  B200W_REQUIRE(n_heads % 4 == 0 && p.d == n_heads * HD, "qkv_attention: heads must be a multiple of 4, head dim 64");
  const size_t smem = (size_t)(1 + 4) * p.max_len * sizeof(float);
  launch_pdl(qkv_attn_kernel, dim3(n_rows, n_heads / 4), dim3(128), smem, stream, p);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

void dec_attention(const DecAttnParams& p, int n_heads, cudaStream_t stream) {
  if (p.n_queries <= 0) return;
  B200W_REQUIRE(n_heads % 2 == 0, "head count must be even");
  if (n_heads % 4 == 0) {
    const size_t smem = (size_t)(1 + 4) * p.max_len * sizeof(float);
    launch_pdl(dec_attn_kernel<4>, dim3(p.n_queries, n_heads / 4), dim3(128), smem, stream, p);
  } else {
    const size_t smem = (size_t)(1 + 2) * p.max_len * sizeof(float);
    launch_pdl(dec_attn_kernel<2>, dim3(p.n_queries, n_heads / 2), dim3(64), smem, stream, p);
  }
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}


// ------------------------------------------------------------------- cross attention (decode)
// Every hypothesis of a stream attends to the same per-chunk cross K/V, so one CTA serves up to
// XQ consecutive query rows of one stream for one head and reads that head's K and V exactly once:
// 16 warps split the 1500 keys, scores for all queries live in shared memory, one warp per query
// normalises, and the P.V product is accumulated per warp and reduced through shared memory.
static constexpr int XW = 16;   // warps per CTA (queries per CTA: template XQ <= 8)

template <int XQ>
__global__ void __launch_bounds__(XW * 32) cross_attn_kernel(const CrossAttnParams p) {
  extern __shared__ float xs[];
  const int len = p.len;
  const int len_pad = (len + 31) & ~31;
  float* s_q = xs;                          // [XQ][64]
  float* s_inv = s_q + XQ * HD;             // [8] (padded so that what follows stays 16-byte aligned)
  float* s_sc = s_inv + 8;                  // [XQ][len_pad]
  float* s_acc = s_sc + XQ * len_pad;       // [XW][XQ][64]
  const int g = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  pdl_wait();
  const int q0 = p.g_start[g], nq = p.g_count[g];
  if (nq == 0) return;
  if (p.export_only && p.head_export[h] < 0) return;
  const long long base = p.kv_base_off[p.stream[q0]];
  const bf16* kb = p.kbase + base + h * p.head_stride;
  const bf16* vb = p.vbase + base + h * p.head_stride;

  for (int i = tid; i < XQ * HD; i += XW * 32) {
    const int q = i >> 6, c = i & 63;
    s_q[i] = q < nq ? __bfloat162float(p.q[(long long)(q0 + q) * p.q_stride + h * HD + c]) * p.scale : 0.f;
  }
  __syncthreads();

  // scores: one key per lane per pass; the next key's row is loaded while this one is multiplied
  {
    uint4 cur[8];
    int key = tid;
    const bool have0 = key < len;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      cur[c] = have0 ? __ldg(reinterpret_cast<const uint4*>(kb + (long long)key * p.kv_row_stride) + c)
                     : make_uint4(0u, 0u, 0u, 0u);
    for (; key < len_pad; key += XW * 32) {
      uint4 nxt[8];
      const int nkey = key + XW * 32;
      const bool haven = nkey < len;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        nxt[c] = haven ? __ldg(reinterpret_cast<const uint4*>(kb + (long long)nkey * p.kv_row_stride) + c)
                       : make_uint4(0u, 0u, 0u, 0u);
      float acc[XQ];
#pragma unroll
      for (int q = 0; q < XQ; ++q) acc[q] = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&cur[c]);
        float kf[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(u2[t]);
          kf[2 * t] = f.x;
          kf[2 * t + 1] = f.y;
        }
#pragma unroll
        for (int q = 0; q < XQ; ++q) {
          const float4 qa = *reinterpret_cast<const float4*>(s_q + q * HD + 8 * c);
          const float4 qb = *reinterpret_cast<const float4*>(s_q + q * HD + 8 * c + 4);
          acc[q] = fmaf(qa.x, kf[0], acc[q]); acc[q] = fmaf(qa.y, kf[1], acc[q]);
          acc[q] = fmaf(qa.z, kf[2], acc[q]); acc[q] = fmaf(qa.w, kf[3], acc[q]);
          acc[q] = fmaf(qb.x, kf[4], acc[q]); acc[q] = fmaf(qb.y, kf[5], acc[q]);
          acc[q] = fmaf(qb.z, kf[6], acc[q]); acc[q] = fmaf(qb.w, kf[7], acc[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < XQ; ++q) s_sc[q * len_pad + key] = key < len ? acc[q] : -INFINITY;
#pragma unroll
      for (int c = 0; c < 8; ++c) cur[c] = nxt[c];
    }
  }
  __syncthreads();

  // softmax: warp q owns query q
  if (warp < nq) {
    float* sc = s_sc + warp * len_pad;
    float mx = -INFINITY;
    for (int i = lane; i < len; i += 32) mx = fmaxf(mx, sc[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = lane; i < len_pad; i += 32) {
      const float e = i < len ? __expf(sc[i] - mx) : 0.f;
      sc[i] = e;
      sum += e;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    if (lane == 0) s_inv[warp] = inv;
    const int ex = p.head_export ? p.head_export[h] : -1;
    if (ex >= 0) {
      float* dst = p.probs_out + ((long long)ex * p.n_queries + q0 + warp) * p.probs_stride;
      for (int i = lane; i < len; i += 32) dst[i] = sc[i] * inv;
    }
  }
  __syncthreads();

  // P.V: half-warps take alternate keys, 16 lanes x 4 dims cover the head
  {
    const int half = lane >> 4, l16 = lane & 15;
    float acc[XQ][4];
#pragma unroll
    for (int q = 0; q < XQ; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
    const bf16* vl = vb + 4 * l16;
    for (int k0 = warp * 2 + half; k0 < len_pad; k0 += 12 * XW * 2) {  // 12 rows in flight per lane
      uint2 v[12];
      int kk[12];
#pragma unroll
      for (int u = 0; u < 12; ++u) {
        kk[u] = k0 + u * XW * 2;
        v[u] = kk[u] < len ? __ldg(reinterpret_cast<const uint2*>(vl + (long long)kk[u] * p.kv_row_stride))
                           : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 12; ++u) {
        if (kk[u] >= len) continue;
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[u].x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[u].y));
#pragma unroll
        for (int q = 0; q < XQ; ++q) {
          const float pw = s_sc[q * len_pad + kk[u]];
          acc[q][0] = fmaf(pw, a.x, acc[q][0]); acc[q][1] = fmaf(pw, a.y, acc[q][1]);
          acc[q][2] = fmaf(pw, b.x, acc[q][2]); acc[q][3] = fmaf(pw, b.y, acc[q][3]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < XQ; ++q)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[q][c] += __shfl_xor_sync(0xffffffffu, acc[q][c], 16);
    if (half == 0) {
#pragma unroll
      for (int q = 0; q < XQ; ++q)
        *reinterpret_cast<float4*>(s_acc + ((warp * XQ + q) * HD) + 4 * l16) =
            make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
    }
  }
  __syncthreads();
  {
    const int q = tid >> 6, c = tid & 63;  // 512 threads = 8 x 64 outputs
    if (q < nq && q < XQ) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < XW; ++w) sum += s_acc[(w * XQ + q) * HD + c];
      p.out[(long long)(q0 + q) * p.out_stride + h * HD + c] = __float2bfloat16(sum * s_inv[q]);
    }
  }
}

// ---- split-KV variant for decode steps (no probability export): grid (group, head, key split).
// Each CTA (4 warps) handles a quarter of the keys and writes an un-normalised partial result
// (max, sum, weighted values); cross_attn_merge_kernel combines the splits.  Many small CTAs per SM
// overlap each other's load and compute phases, which one big CTA per (stream, head) cannot.
static constexpr int SW = 4;       // warps per CTA
static constexpr int KSPLIT = 4;   // key splits

template <int XQ>
__global__ void __launch_bounds__(SW * 32) cross_attn_split_kernel(const CrossAttnParams p, float* __restrict__ part) {
  extern __shared__ float xs3[];
  const int keys_per = ((p.len + KSPLIT - 1) / KSPLIT + 31) & ~31;   // 384 for 1500
  float* s_q = xs3;                          // [XQ][64]
  float* s_sc = s_q + XQ * HD;               // [XQ][keys_per]
  float* s_acc = s_sc + XQ * keys_per;       // [SW][XQ][64]
  float* s_ml = s_acc + SW * XQ * HD;        // [2][8]
  const int g = blockIdx.x, h = blockIdx.y, sp = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  pdl_wait();
  const int q0 = p.g_start[g], nq = p.g_count[g];
  if (nq == 0) return;  // padding rows of a finished stream: nothing to attend, no K/V traffic
  const long long base = p.kv_base_off[p.stream[q0]];
  const int k_lo = sp * keys_per;
  const int n_keys = max(0, min(p.len - k_lo, keys_per));
  const bf16* kb = p.kbase + base + h * p.head_stride + (long long)k_lo * p.kv_row_stride;
  const bf16* vb = p.vbase + base + h * p.head_stride + (long long)k_lo * p.kv_row_stride;

  for (int i = tid; i < XQ * HD; i += SW * 32) {
    const int q = i >> 6, c = i & 63;
    s_q[i] = q < nq ? __bfloat162float(p.q[(long long)(q0 + q) * p.q_stride + h * HD + c]) * p.scale : 0.f;
  }
  __syncthreads();
  for (int key = tid; key < keys_per; key += SW * 32) {
    float acc[XQ];
#pragma unroll
    for (int q = 0; q < XQ; ++q) acc[q] = 0.f;
    if (key < n_keys) {
      const bf16* kp = kb + (long long)key * p.kv_row_stride;
      uint4 kr[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) kr[c] = __ldg(reinterpret_cast<const uint4*>(kp) + c);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&kr[c]);
        float kf[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(u2[t]);
          kf[2 * t] = f.x;
          kf[2 * t + 1] = f.y;
        }
#pragma unroll
        for (int q = 0; q < XQ; ++q) {
          const float4 qa = *reinterpret_cast<const float4*>(s_q + q * HD + 8 * c);
          const float4 qb = *reinterpret_cast<const float4*>(s_q + q * HD + 8 * c + 4);
          acc[q] = fmaf(qa.x, kf[0], acc[q]); acc[q] = fmaf(qa.y, kf[1], acc[q]);
          acc[q] = fmaf(qa.z, kf[2], acc[q]); acc[q] = fmaf(qa.w, kf[3], acc[q]);
          acc[q] = fmaf(qb.x, kf[4], acc[q]); acc[q] = fmaf(qb.y, kf[5], acc[q]);
          acc[q] = fmaf(qb.z, kf[6], acc[q]); acc[q] = fmaf(qb.w, kf[7], acc[q]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < XQ; ++q) s_sc[q * keys_per + key] = key < n_keys ? acc[q] : -INFINITY;
  }
  __syncthreads();
  for (int q = warp; q < XQ; q += SW) {  // local softmax of the split (max / sum kept for the merge)
    float* sc = s_sc + q * keys_per;
    float mx = -INFINITY;
    for (int i = lane; i < keys_per; i += 32) mx = fmaxf(mx, sc[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = lane; i < keys_per; i += 32) {
      const float e = mx > -INFINITY ? __expf(sc[i] - mx) : 0.f;
      sc[i] = e;
      sum += e;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) { s_ml[q] = mx; s_ml[8 + q] = sum; }
  }
  __syncthreads();
  {
    const int half = lane >> 4, l16 = lane & 15;
    float acc[XQ][4];
#pragma unroll
    for (int q = 0; q < XQ; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
    const bf16* vl = vb + 4 * l16;
    for (int k0 = warp * 2 + half; k0 < keys_per; k0 += 6 * SW * 2) {  // 6 rows in flight per lane
      uint2 v[6];
      int kk[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        kk[u] = k0 + u * SW * 2;
        v[u] = kk[u] < n_keys ? __ldg(reinterpret_cast<const uint2*>(vl + (long long)kk[u] * p.kv_row_stride))
                              : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        if (kk[u] >= n_keys) continue;
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[u].x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[u].y));
#pragma unroll
        for (int q = 0; q < XQ; ++q) {
          const float pw = s_sc[q * keys_per + kk[u]];
          acc[q][0] = fmaf(pw, a.x, acc[q][0]); acc[q][1] = fmaf(pw, a.y, acc[q][1]);
          acc[q][2] = fmaf(pw, b.x, acc[q][2]); acc[q][3] = fmaf(pw, b.y, acc[q][3]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < XQ; ++q)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[q][c] += __shfl_xor_sync(0xffffffffu, acc[q][c], 16);
    if (half == 0) {
#pragma unroll
      for (int q = 0; q < XQ; ++q)
        *reinterpret_cast<float4*>(s_acc + ((warp * XQ + q) * HD) + 4 * l16) =
            make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
    }
  }
  __syncthreads();
  // partial record of this (group, head, split): [XQ][66] = 64 values, max, sum
  float* rec = part + (((long long)g * gridDim.y + h) * KSPLIT + sp) * (8 * 66);
  for (int i = tid; i < XQ * HD; i += SW * 32) {
    const int q = i >> 6, c = i & 63;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < SW; ++w) sum += s_acc[(w * XQ + q) * HD + c];
    rec[q * 66 + c] = sum;
  }
  if (tid < XQ) { rec[tid * 66 + 64] = s_ml[tid]; rec[tid * 66 + 65] = s_ml[8 + tid]; }
}

__global__ void cross_attn_merge_kernel(const CrossAttnParams p, const float* __restrict__ part) {
  const int g = blockIdx.x, h = blockIdx.y;
  const int q = threadIdx.x >> 6, c = threadIdx.x & 63;
  pdl_trigger();
  pdl_wait();
  if (q >= p.g_count[g]) return;
  const float* rec = part + (((long long)g * gridDim.y + h) * KSPLIT) * (8 * 66) + q * 66;
  float M = -INFINITY;
#pragma unroll
  for (int s = 0; s < KSPLIT; ++s) M = fmaxf(M, rec[s * 8 * 66 + 64]);
  float L = 0.f, O = 0.f;
#pragma unroll
  for (int s = 0; s < KSPLIT; ++s) {
    const float m = rec[s * 8 * 66 + 64];
    const float w = m > -INFINITY ? __expf(m - M) : 0.f;
    L = fmaf(rec[s * 8 * 66 + 65], w, L);
    O = fmaf(rec[s * 8 * 66 + c], w, O);
  }
  p.out[(long long)(p.g_start[g] + q) * p.out_stride + h * HD + c] = __float2bfloat16(O / L);
}

template <int XQ>
static void launch_cross_split(const CrossAttnParams& p, int n_groups, int n_heads, float* part, cudaStream_t stream) {
  const int keys_per = ((p.len + KSPLIT - 1) / KSPLIT + 31) & ~31;
  const size_t smem = sizeof(float) * (XQ * HD + (size_t)XQ * keys_per + (size_t)SW * XQ * HD + 16);
  launch_pdl(cross_attn_split_kernel<XQ>, dim3(n_groups, n_heads, KSPLIT), dim3(SW * 32), smem, stream, p, part);
  B200W_COUNT_LAUNCH();
  launch_pdl(cross_attn_merge_kernel, dim3(n_groups, n_heads), dim3(XQ * 64), 0, stream, p, (const float*)part);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// ---- tensor-core variant for decode steps (no probability export): cross_attn_sp_kernel.
// Persistent CTAs (two per SM) walk the (live group, head) work items.  Per item the head's K and V ([1500][64]
// bf16 each, 384 KB together) stream ONCE through a TMA ring, and the arithmetic runs on tcgen05:
//   S^T[keys x 16] = K_tile[128 keys x 64] . Q^T        (keys = TMEM lanes: the softmax is spread over 128 threads)
//   O[128 x 64]   += P[128 x 128 keys] . V_tile[128 keys x 64]   (rows >= 8 of P are don't-care: a
//   row of A only feeds the same row of D, so the operand tile may overlap unrelated shared memory)
// Warps 0-3: probabilities and output.  Warp 4: TMA producer (lane 0) and Q staging (lanes 1-31).  Warp 5: MMA issuer.
// Load balance and batch invariance: every item is ALWAYS cut into the same `pieces` key ranges (a divisor of
// the tile count, chosen by the caller from the shapes alone), the units (item, piece) are numbered item-major
// and dealt to the CTAs in contiguous runs.  A piece leaves an un-normalised partial result in global scratch
// and the last piece of an item to finish (atomic counter, self-resetting) merges them in piece order, so an
// item's arithmetic never depends on how many other items share the launch.  pieces = 1: items are computed whole.
static constexpr int XT_MAX_TILES = 12;  // keys <= 1536
static constexpr int XT_Q_BYTES = 2048;
static constexpr int XT_MAX_PIECES = 6;
static constexpr int XT_CTAS_PER_SM = 2;
static constexpr int XT_PART_FLOATS = 8 * 64 + 16;  // O[8][64], m[8], l[8]

struct CrossTcParams {
  const bf16* q;
  int q_stride;
  const int* stream;
  const int* g_start;
  const int* g_count;
  const int* n_live_groups;  // device: groups [0, n_live) have at least one live row
  int n_heads;
  int blocks_per_stream;  // L * 2 * H  ([1500][64] blocks per stream in the cross K/V buffer)
  int layer_block0;       // 2 * layer * H
  int len;
  bf16* out;
  int out_stride;
  float scale_log2e;
  int pieces;      // key ranges per item (divides the tile count; 1 = whole items)
  float* partial;  // [item][XT_MAX_PIECES][XT_PART_FLOATS]
  int* counters;   // [item], zero between launches
  int* err;        // set to 1 when a score exceeds the kernel's exponent range (checked by the engine)
};

struct XtUnit {
  int item, t0, t1, piece, n_pieces, tail_idx;
};
// u-th unit of this CTA (every role evaluates the same sequence): CTA c owns the units [n c / G, n (c+1) / G)
__device__ __forceinline__ bool xt_unit(int u, int n_items, int n_tiles, XtUnit& un, int pieces) {
  const long long n_units = (long long)n_items * pieces;
  const int lo = (int)(n_units * blockIdx.x / gridDim.x), hi = (int)(n_units * (blockIdx.x + 1) / gridDim.x);
  const int g = lo + u;
  if (g >= hi) return false;
  un.item = g / pieces;
  un.piece = g - un.item * pieces;
  un.n_pieces = pieces;
  un.tail_idx = un.item;
  const int tpp = n_tiles / pieces;
  un.t0 = un.piece * tpp;
  un.t1 = un.t0 + tpp;
  return true;
}

// ---- the kernel.  Single pass over the keys: a classic softmax needs the row maximum before the first probability
// can be formed, i.e. all scores of a unit before any P.V product (round 1's kernel did that in two passes over the
// TMEM-resident scores and kept 12 score tiles alive).  Here the reference value of the exponentials is the maximum
// of the unit's FIRST tile: P = exp2(s c - m_0) is exact arithmetic for any reference
// (bf16 / fp32 keep their relative precision over 2^+-126), and a later score exceeding m_0 by more than 2^100
// -- which no attention row does -- raises an error flag that the engine turns into an exception (never silent).
// Every tile then flows  K_j -> S_j = K_j Q^T -> P_j -> O += P_j V_j  with no unit-wide synchronisation:
//   * K and V tiles interleaved in one TMA ring (K runs XS_LA tiles ahead of V), 6 slots;
//   * 8 score slots in TMEM and 4 probability slots in shared memory, recycled through mbarriers;
//   * O double-buffered in TMEM, so the output of unit i is read while unit i+1 accumulates.
static constexpr int XS_RING = 6;
static constexpr int XS_NS = 8;   // score slots (16 TMEM columns each)
static constexpr int XS_NP = 4;   // probability slots (2 KB each)
static constexpr int XS_LA = 2;   // the QK product runs this many tiles ahead of the PV product
static constexpr int XS_P_BYTES = XS_NP * 2048;
static constexpr int XS_N_BARS = 2 * XS_RING + 2 * XS_NS + 2 * XS_NP + 8;
static constexpr int XS_SMEM = XS_P_BYTES + 2 * XT_Q_BYTES + XS_RING * TILE_BYTES + XS_N_BARS * 8 + 512 + 16 + 1024;

__global__ void __launch_bounds__(192, 2)
cross_attn_sp_kernel(const __grid_constant__ CUtensorMap tmKV, const CrossTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sP = smem;                        // XS_NP slots of 2 KB (two 64-key halves of [8 queries][64] bf16)
  uint8_t* sQ = sP + XS_P_BYTES;             // two [16][64] Q tiles (double-buffered across units)
  uint8_t* ring = sQ + 2 * XT_Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + XS_RING * TILE_BYTES);
  float* s_red = reinterpret_cast<float*>(bars + XS_N_BARS);  // [3][4][8]: tile-0 max, sums, running max
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_red + 96 + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * XS_RING;
  const uint32_t bar_s0 = empty0 + 8 * XS_RING;        // [NS] scores of the slot are in TMEM
  const uint32_t bar_sf0 = bar_s0 + 8 * XS_NS;         // [NS] the softmax has read the slot
  const uint32_t bar_p0 = bar_sf0 + 8 * XS_NS;         // [NP] probabilities of the slot are in shared memory
  const uint32_t bar_pf0 = bar_p0 + 8 * XS_NP;         // [NP] the PV product has consumed the slot
  const uint32_t bar_q0 = bar_pf0 + 8 * XS_NP;         // [2] Q buffer is staged
  const uint32_t bar_qf0 = bar_q0 + 16;                // [2] the unit's QK products are done with the Q buffer
  const uint32_t bar_o0 = bar_qf0 + 16;                // [2] O buffer is complete
  const uint32_t bar_of0 = bar_o0 + 16;                // [2] O buffer has been read out
  pdl_trigger();
  if (tid == 0) {
    for (int i = 0; i < 2 * XS_RING + XS_NS; ++i) mbar_init(full0 + 8 * i, 1);  // ring + bar_s
    for (int i = 0; i < XS_NS; ++i) mbar_init(bar_sf0 + 8 * i, 128);
    for (int i = 0; i < XS_NP; ++i) mbar_init(bar_p0 + 8 * i, 128);
    for (int i = 0; i < XS_NP; ++i) mbar_init(bar_pf0 + 8 * i, 1);
    mbar_init(bar_q0, 31);
    mbar_init(bar_q0 + 8, 31);
    for (int i = 0; i < 6; ++i) mbar_init(bar_qf0 + 8 * i, 1);  // qfree[2], o[2], ofree[2]
    fence_barrier_init();
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_S = *tmem_slot, tmem_O = tmem_S + 16 * XS_NS;  // S slots [0,128), O buffers at 128 and 192
  const int n_tiles = (p.len + 127) >> 7;
  const int n_items = *p.n_live_groups * p.n_heads;
  XtUnit un;

  if (warp == 4) {
    if (lane == 0) {  // ---------------- TMA producer: K_j and V_(j-LA) interleaved, in the order the MMA warp consumes them
      uint32_t slot = 0, phase = 0;
      for (int u = 0; xt_unit(u, n_items, n_tiles, un, p.pieces); ++u) {
        const int g = un.item / p.n_heads, h = un.item - g * p.n_heads;
        const int blk = p.stream[p.g_start[g]] * p.blocks_per_stream + p.layer_block0 + h;
        const int nt = un.t1 - un.t0;
        for (int j = 0; j < nt + XS_LA; ++j) {
          for (int kv = 0; kv < 2; ++kv) {
            const int tile = kv == 0 ? j : j - XS_LA;
            if (tile < 0 || tile >= nt) continue;
            mbar_wait(empty0 + 8 * slot, phase ^ 1);
            const uint32_t fb = full0 + 8 * slot;
            mbar_arrive_expect_tx(fb, TILE_BYTES);
            tma_load_3d(smem_u32(ring + slot * TILE_BYTES), &tmKV, fb, 0, (un.t0 + tile) * 128, blk + kv * p.n_heads);
            if (++slot == XS_RING) { slot = 0; phase ^= 1; }
          }
        }
      }
    } else {  // ---------------- lanes 1..31: stage the Q tile of unit `it` one unit ahead of its use
      pdl_wait();
      for (int it = 0; xt_unit(it, n_items, n_tiles, un, p.pieces); ++it) {
        const int g = un.item / p.n_heads, h = un.item - g * p.n_heads;
        const int q0 = p.g_start[g], nq = p.g_count[g];
        uint4 u[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {  // 128 16-byte pieces over 31 lanes
          const int i = (lane - 1) + 31 * r;
          const int qr = i >> 3, ch = i & 7;
          u[r] = make_uint4(0u, 0u, 0u, 0u);
          if (i < 128 && qr < nq)
            u[r] = *reinterpret_cast<const uint4*>(p.q + (long long)(q0 + qr) * p.q_stride + h * HD + 8 * ch);
        }
        // buffer it&1 was last read by the QK products of unit it-2
        mbar_wait(bar_qf0 + 8 * (it & 1), (uint32_t)(((it >> 1) & 1) ^ 1));
        uint8_t* dst = sQ + (it & 1) * XT_Q_BYTES;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const int i = (lane - 1) + 31 * r;
          const int qr = i >> 3, ch = i & 7;
          if (i < 128) *reinterpret_cast<uint4*>(dst + qr * 128 + ((ch ^ (qr & 7)) << 4)) = u[r];
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_q0 + 8 * (it & 1));
      }
    }
  } else if (warp == 5) {
    {  // ---------------- MMA issuer: the whole warp runs the loop (warp-uniform state in uniform registers), an
       // elected lane issues -- a single divergent lane paid a chain of R2UR moves per MMA
      constexpr uint32_t idescS = make_idesc_bf16(128, 16, 0, 0);  // K (K-major) x Q (K-major)
      constexpr uint32_t idescO = make_idesc_bf16(128, 64, 0, 1);  // P (K-major) x V (MN-major)
      uint32_t slot = 0, phase = 0, g0 = 0;
      for (int it = 0; xt_unit(it, n_items, n_tiles, un, p.pieces); ++it) {
        const int nt = un.t1 - un.t0;
        const uint32_t par = it & 1, upar = (uint32_t)((it >> 1) & 1);
        mbar_wait(bar_q0 + 8 * par, upar);
        tc_fence_after();
        const uint64_t qd = make_sw128_desc(smem_u32(sQ + par * XT_Q_BYTES));
        const uint32_t d_o = tmem_O + 64 * par;
        for (int j = 0; j < nt + XS_LA; ++j) {
          if (j < nt) {  // S_j = K_j Q^T
            const uint32_t gs = g0 + j, ss = gs % XS_NS, us = gs / XS_NS;
            mbar_wait(full0 + 8 * slot, phase);
            mbar_wait(bar_sf0 + 8 * ss, (us & 1u) ^ 1u);
            tc_fence_after();
            const uint64_t kd = make_sw128_desc(smem_u32(ring + slot * TILE_BYTES));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tmem_S + 16 * ss, kd + 2 * k, qd + 2 * k, idescS, k != 0);
              tc_commit(empty0 + 8 * slot);
              tc_commit(bar_s0 + 8 * ss);
              if (j == nt - 1) tc_commit(bar_qf0 + 8 * par);
            }
            __syncwarp();
            if (++slot == XS_RING) { slot = 0; phase ^= 1; }
          }
          if (j >= XS_LA) {  // O += P_jj V_jj
            const int jj = j - XS_LA;
            const uint32_t gp = g0 + jj, ps = gp % XS_NP, up = gp / XS_NP;
            if (jj == 0) mbar_wait(bar_of0 + 8 * par, upar ^ 1u);  // the output of unit it-2 has left this O buffer
            mbar_wait(bar_p0 + 8 * ps, up & 1u);
            mbar_wait(full0 + 8 * slot, phase);
            tc_fence_after();
            const uint64_t vd = make_sw128_desc(smem_u32(ring + slot * TILE_BYTES));
            const uint64_t pd0 = make_sw128_desc(smem_u32(sP + ps * 2048)), pd1 = make_sw128_desc(smem_u32(sP + ps * 2048 + 1024));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                tc_mma_bf16(d_o, (k < 4 ? pd0 : pd1) + 2 * (k & 3), vd + 128 * k, idescO, (jj | k) != 0);
              tc_commit(empty0 + 8 * slot);
              tc_commit(bar_pf0 + 8 * ps);
            }
            __syncwarp();
            if (++slot == XS_RING) { slot = 0; phase ^= 1; }
          }
        }
        if (elect_one()) tc_commit(bar_o0 + 8 * par);
        __syncwarp();
        g0 += nt;
      }
    }
  } else {  // ---------------- warps 0..3: probabilities (thread = key row of every tile) and output
    pdl_wait();
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const int kc = tid & 63, half = tid >> 6;
    uint32_t g0 = 0;
    for (int it = 0; xt_unit(it, n_items, n_tiles, un, p.pieces); ++it) {
      const int g = un.item / p.n_heads, h = un.item - g * p.n_heads;
      const int q0 = p.g_start[g], nq = p.g_count[g];
      const int nt = un.t1 - un.t0;
      const uint32_t par = it & 1, upar = (uint32_t)((it >> 1) & 1);
      const int key0 = un.t0 * 128 + tid;
      float mref[8], sum[8], gmx[8];
      uint32_t sc[8];
      // reference value: the maximum of the unit's first tile
      {
        const uint32_t ss = g0 % XS_NS, us = g0 / XS_NS;
        mbar_wait(bar_s0 + 8 * ss, us & 1u);
        tc_fence_after();
        tmem_ld8(tmem_S + lane_sel + 16 * ss, sc);
        tmem_ld_wait();
        const bool valid = key0 < p.len;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float v = valid ? __uint_as_float(sc[q]) : -INFINITY;
          for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
          mref[q] = v;
        }
        if (lane < 8) {
          float v = mref[0];
#pragma unroll
          for (int q = 1; q < 8; ++q) v = lane == q ? mref[q] : v;
          s_red[warp * 8 + lane] = v;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          mref[q] = fmaxf(fmaxf(s_red[q], s_red[8 + q]), fmaxf(s_red[16 + q], s_red[24 + q])) * p.scale_log2e;
          sum[q] = 0.f;
          gmx[q] = -INFINITY;
        }
      }
#pragma unroll 1
      for (int j = 0; j < nt; ++j) {
        const uint32_t gs = g0 + j, ss = gs % XS_NS, us = gs / XS_NS, ps = gs % XS_NP, up = gs / XS_NP;
        if (j > 0) {
          mbar_wait(bar_s0 + 8 * ss, us & 1u);
          tc_fence_after();
          tmem_ld8(tmem_S + lane_sel + 16 * ss, sc);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(bar_sf0 + 8 * ss);  // the slot may take the scores of tile j + NS
        mbar_wait(bar_pf0 + 8 * ps, (up & 1u) ^ 1u);  // the PV product of tile j - NP is done with this slot
        const bool valid = key0 + j * 128 < p.len;
        uint8_t* pt = sP + ps * 2048 + half * 1024 + (kc & 7) * 2;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q < nq) {
            const float sv = __uint_as_float(sc[q]);
            const float e = valid ? ex2_approx(fmaf(sv, p.scale_log2e, -mref[q])) : 0.f;
            sum[q] += e;
            gmx[q] = valid ? fmaxf(gmx[q], sv) : gmx[q];
            *reinterpret_cast<bf16*>(pt + q * 128 + (((kc >> 3) ^ q) << 4)) = __float2bfloat16(e);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_p0 + 8 * ps);
      }
      g0 += nt;
      // unit end: row sums and the overflow guard
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        for (int o = 16; o > 0; o >>= 1) {
          sum[q] += __shfl_xor_sync(0xffffffffu, sum[q], o);
          gmx[q] = fmaxf(gmx[q], __shfl_xor_sync(0xffffffffu, gmx[q], o));
        }
      }
      if (lane < 8) {
        float v = sum[0], m = gmx[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) { v = lane == q ? sum[q] : v; m = lane == q ? gmx[q] : m; }
        s_red[32 + warp * 8 + lane] = v;
        s_red[64 + warp * 8 + lane] = m;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 0) {
        mbar_wait(bar_o0 + 8 * par, upar);
        tc_fence_after();
        float lsum = 1.f, mr = 0.f;
        if (lane < 8) {
          lsum = s_red[32 + lane] + s_red[40 + lane] + s_red[48 + lane] + s_red[56 + lane];
          const float gm = fmaxf(fmaxf(s_red[64 + lane], s_red[72 + lane]), fmaxf(s_red[80 + lane], s_red[88 + lane]));
          mr = fmaxf(fmaxf(s_red[lane], s_red[8 + lane]), fmaxf(s_red[16 + lane], s_red[24 + lane])) * p.scale_log2e;
          if (lane < nq && gm * p.scale_log2e - mr > 100.f) atomicExch(p.err, 1);  // never silent
        }
        float* rec = p.partial + ((long long)un.tail_idx * XT_MAX_PIECES + un.piece) * XT_PART_FLOATS;
        const float inv = 1.0f / lsum;
        uint32_t o_lo[32], o_hi[32];
        tmem_ld32(tmem_O + 64 * par, o_lo);
        tmem_ld32(tmem_O + 64 * par + 32, o_hi);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_of0 + 8 * par);  // the buffer may take unit it + 2
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          const uint32_t* o = hc == 0 ? o_lo : o_hi;
          if (un.n_pieces == 1) {
            if (lane < nq) {
              bf16* dst = p.out + (long long)(q0 + lane) * p.out_stride + h * HD + 32 * hc;
#pragma unroll
              for (int c = 0; c < 32; c += 8) {
                uint4 u;
                __nv_bfloat162* u2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  u2[i] = __floats2bfloat162_rn(__uint_as_float(o[c + 2 * i]) * inv, __uint_as_float(o[c + 2 * i + 1]) * inv);
                *reinterpret_cast<uint4*>(dst + c) = u;
              }
            }
          } else if (lane < 8) {
            float4* dst = reinterpret_cast<float4*>(rec + lane * 64 + 32 * hc);
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              dst[c >> 2] = make_float4(__uint_as_float(o[c]), __uint_as_float(o[c + 1]), __uint_as_float(o[c + 2]),
                                        __uint_as_float(o[c + 3]));
          }
        }
        if (un.n_pieces > 1) {
          if (lane < 8) {
            rec[512 + lane] = mr;
            rec[520 + lane] = lsum;
          }
          __threadfence();
          __syncwarp();
          int old = 0;
          if (lane == 0) old = atomicAdd(p.counters + un.tail_idx, 1);
          old = __shfl_sync(0xffffffffu, old, 0);
          if (old == un.n_pieces - 1) {
            __threadfence();
            const float* base = p.partial + (long long)un.tail_idx * XT_MAX_PIECES * XT_PART_FLOATS;
            const int q = lane >> 2, c0 = (lane & 3) * 16;  // lane -> (query, 16 dims)
            float M = -INFINITY;
            for (int s2 = 0; s2 < un.n_pieces; ++s2) M = fmaxf(M, __ldcg(base + s2 * XT_PART_FLOATS + 512 + q));
            float Lsum = 0.f, acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
            for (int s2 = 0; s2 < un.n_pieces; ++s2) {
              const float* r2 = base + s2 * XT_PART_FLOATS;
              const float w = ex2_approx(__ldcg(r2 + 512 + q) - M);
              Lsum = fmaf(__ldcg(r2 + 520 + q), w, Lsum);
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(r2 + q * 64 + c0 + i));
                acc[i] = fmaf(v.x, w, acc[i]); acc[i + 1] = fmaf(v.y, w, acc[i + 1]);
                acc[i + 2] = fmaf(v.z, w, acc[i + 2]); acc[i + 3] = fmaf(v.w, w, acc[i + 3]);
              }
            }
            if (q < nq) {
              const float invL = 1.0f / Lsum;
              bf16* dst = p.out + (long long)(q0 + q) * p.out_stride + h * HD + c0;
              uint4 u, w4;
              __nv_bfloat162* u2 = reinterpret_cast<__nv_bfloat162*>(&u);
              __nv_bfloat162* w2 = reinterpret_cast<__nv_bfloat162*>(&w4);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                u2[i] = __floats2bfloat162_rn(acc[2 * i] * invL, acc[2 * i + 1] * invL);
                w2[i] = __floats2bfloat162_rn(acc[8 + 2 * i] * invL, acc[8 + 2 * i + 1] * invL);
              }
              *reinterpret_cast<uint4*>(dst) = u;
              *reinterpret_cast<uint4*>(dst + 8) = w4;
            }
            if (lane == 0) p.counters[un.tail_idx] = 0;  // ready for the next launch
          }
        }
      }
      // s_red is rewritten by the next unit: everyone must be past their reads first
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tmem_dealloc(tmem_S, 256);
  }
}


static constexpr int XT_MAX_ITEMS = 64 * 32;  // (live group, head) items per launch: <= 64 groups, <= 32 heads
size_t cross_attention_tc_scratch_bytes() {
  return (size_t)XT_MAX_ITEMS * (XT_MAX_PIECES * XT_PART_FLOATS * sizeof(float) + sizeof(int)) + 1024;
}
size_t cross_attention_tc_error_offset() { return (size_t)XT_MAX_ITEMS * sizeof(int); }

void cross_attention_tc(const bf16* cross_kv, int n_streams, int n_layers, int layer, int n_heads, int len,
                        const bf16* q, int q_stride, const int* stream, const int* g_start, const int* g_count,
                        const int* n_live_groups, int n_groups, bf16* out, int out_stride, float scale, void* scratch,
                        cudaStream_t cs, int pieces) {
  if (n_groups <= 0) return;
  B200W_REQUIRE(len <= 128 * XT_MAX_TILES, "cross_attention_tc: at most 1536 keys");
  CUtensorMap tm;
  make_tmap_bf16_3d(&tm, cross_kv, HD, len, (uint64_t)n_streams * n_layers * 2 * n_heads, HD, (uint64_t)len * HD, HD,
                    128);
  CrossTcParams p;
  p.q = q; p.q_stride = q_stride; p.stream = stream; p.g_start = g_start; p.g_count = g_count;
  p.n_live_groups = n_live_groups; p.n_heads = n_heads; p.blocks_per_stream = n_layers * 2 * n_heads;
  p.layer_block0 = 2 * layer * n_heads; p.len = len; p.out = out; p.out_stride = out_stride;
  p.scale_log2e = scale * 1.4426950408889634f;
  const int n_tiles = (len + 127) >> 7;
  if (pieces < 1 || pieces > XT_MAX_PIECES || n_tiles % pieces != 0) pieces = 1;
  p.pieces = pieces;
  B200W_REQUIRE(n_groups * n_heads <= XT_MAX_ITEMS, "cross_attention_tc: too many (group, head) items");
  // scratch: zero-initialised once by the owner; [XT_MAX_ITEMS counters][XT_MAX_ITEMS * pieces * record]
  p.counters = static_cast<int*>(scratch);
  p.err = p.counters + XT_MAX_ITEMS;
  p.partial = reinterpret_cast<float*>(static_cast<char*>(scratch) + (((XT_MAX_ITEMS + 64) * sizeof(int) + 255) & ~(size_t)255));
  const int units = n_groups * n_heads * pieces;
  const int max_ctas = XT_CTAS_PER_SM * num_sms();
  ensure_dynamic_smem(cross_attn_sp_kernel, XS_SMEM);
  launch_pdl(cross_attn_sp_kernel, dim3(units < max_ctas ? units : max_ctas), dim3(192), (size_t)XS_SMEM, cs, tm, p);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

size_t cross_attention_scratch_bytes(int n_groups, int n_heads) {
  return (size_t)n_groups * n_heads * KSPLIT * 8 * 66 * sizeof(float);
}

template <int XQ>
static void launch_cross_attn(const CrossAttnParams& p, int n_groups, int n_heads, cudaStream_t stream) {
  const int len_pad = (p.len + 31) & ~31;
  const size_t smem = sizeof(float) * (XQ * HD + 8 + (size_t)XQ * len_pad + (size_t)XW * XQ * HD);
  ensure_dynamic_smem(cross_attn_kernel<XQ>, smem);
  launch_pdl(cross_attn_kernel<XQ>, dim3(n_groups, n_heads), dim3(XW * 32), smem, stream, p);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

void cross_attention(const CrossAttnParams& p, int n_groups, int n_heads, int max_group, cudaStream_t stream,
                     float* split_scratch) {
  if (n_groups <= 0) return;
  if (split_scratch && !p.head_export && max_group <= 8 && n_groups <= 64) {
    // few queries (decode steps): split the keys over 4 CTAs per (group, head) and merge
    if (max_group <= 1) launch_cross_split<1>(p, n_groups, n_heads, split_scratch, stream);
    else if (max_group <= 2) launch_cross_split<2>(p, n_groups, n_heads, split_scratch, stream);
    else if (max_group <= 4) launch_cross_split<4>(p, n_groups, n_heads, split_scratch, stream);
    else if (max_group <= 5) launch_cross_split<5>(p, n_groups, n_heads, split_scratch, stream);
    else launch_cross_split<8>(p, n_groups, n_heads, split_scratch, stream);
    return;
  }
  // specialised on the group size: a beam-5 decode step does 5/8 of the work of the generic kernel
  if (max_group <= 1) launch_cross_attn<1>(p, n_groups, n_heads, stream);
  else if (max_group <= 2) launch_cross_attn<2>(p, n_groups, n_heads, stream);
  else if (max_group <= 4) launch_cross_attn<4>(p, n_groups, n_heads, stream);
  else if (max_group <= 5) launch_cross_attn<5>(p, n_groups, n_heads, stream);
  else launch_cross_attn<8>(p, n_groups, n_heads, stream);
}

}  // namespace b200w

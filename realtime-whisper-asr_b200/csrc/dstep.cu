// Fused decode-step segments (persistent program kernel).
//
// A decode step of the engine is a chain of ~13 dependent small operations per decoder layer (LayerNorm, five
// weight-streaming GEMMs with 160 activation rows, their split-K reductions, two attentions).  As separate launches
// each link costs a launch + prologue + drain (measured floor 5.4 us for the swap-AB GEMM, profiles/
// r01_launch_floor.txt) and the weights of the next GEMM cannot be requested before its kernel starts.  Here the
// links between two attentions run as ONE kernel: a short program of phases executed by one persistent CTA per SM,
//     phase i  --grid barrier-->  phase i+1,
// with four roles per CTA:
//   warp 0   W producer: walks the GEMM phases of the whole program and streams this CTA's weight tiles through a
//            TMA ring as slots free up -- it never waits for a grid barrier (weights do not depend on activations),
//            so the weights of phase i+1 arrive while phase i computes;
//   warp 1   X producer: after the barrier that completes the activations of a GEMM phase, loads the activation
//            tiles of the same ring stages;
//   warp 2   MMA issuer: D[128 weight rows][rows] += W_tile . X_tile^T per 64-wide k-block (swap-AB, as skinny.cu),
//            fp32 accumulators in TMEM, double-buffered;
//   warps 3-10  epilogue (TMEM -> fp32 split-K slabs) and the row-wise phases (LayerNorm + slab reduction,
//            finalize, embedding); they run the grid barrier.
// The arithmetic of every phase is the one of the stand-alone kernels (skinny.cu, kernels.cu: same k ranges, same
// summation orders), so the fused step returns the same bits as the launch-per-op step.
//
// Grid barrier: sense-reversing counter in global memory (arrive: threadfence + atomicAdd; the last arriver
// resets the count and bumps the generation).  All CTAs are co-resident (grid = SM count, one CTA per SM by its
// shared-memory footprint).  Every spin has a watchdog that traps instead of hanging the GPU.
#include "dstep.h"

#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "gemm.h"
#include "ptx.cuh"

namespace b200w {

static constexpr int DS_THREADS = 352;  // 11 warps
static constexpr int DS_VEC_WARP0 = 3;
static constexpr int DS_VEC_THREADS = 256;
static constexpr int DS_BK = 64;
static constexpr int DS_MAX_STAGES = 8;
static constexpr int DS_W_BYTES = 128 * 128;  // one weight tile: 128 rows x 64 k bf16

struct __align__(16) DsPhase {
  int type, rows;
  int N, K, splits, n_tiles, total_kb, Mpad;  // GEMM
  int n_slabs, d, act, pad0;
  long long ldo, slab_stride;
  float* slabs;
  const float* bias;
  float* x;
  const float* gamma;
  const float* beta;
  bf16* y;
  bf16* out;
  const int* tok;
  const int* pos;
  const bf16* tok_emb;
  const float* pos_emb;
};

struct __align__(128) DsLaunch {
  CUtensorMap tm[2 * DS_MAX_PHASES];  // [2p] weights, [2p+1] activations of phase p (GEMM phases)
  DsPhase ph[DS_MAX_PHASES];
  int n_phases, stages, stage_bytes, reserved;
  unsigned* bar;  // [0] arrivals, [1] generation
  long long* trace;  // development aid: [cta][24] globaltimer stamps of the vector warps (phase work done / barrier passed), or null
};
__device__ __forceinline__ long long ds_gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define DS_TRACE(slot) do { if (trace && vt == 0) trace[blockIdx.x * 24 + (slot)] = ds_gtime(); } while (0)

__device__ __forceinline__ void ds_trap_if(bool bad) {
  if (bad) __trap();
}
// bounded mbarrier wait (a protocol bug must fault, never hang)
__device__ __noinline__ void ds_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfff) == 0) ds_trap_if(clock64() - t0 > 6000000000LL);
  }
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void vec_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void half_sync(int h) { asm volatile("bar.sync %0, 128;" ::"r"(2 + h) : "memory"); }

// Grid-wide barrier of the vector warps (all 256 vector threads of every CTA call it).
__device__ __forceinline__ void ds_grid_barrier(unsigned* bar, int vt) {
  vec_sync();  // every vector thread's global writes are ordered before thread 0's fence (cumulativity)
  if (vt == 0) {
    const unsigned gen = ld_acquire_u32(bar + 1);
    __threadfence();
    if (atomicAdd(bar, 1u) == gridDim.x - 1) {
      bar[0] = 0;
      __threadfence();
      st_release_u32(bar + 1, gen + 1);
    } else {
      const long long t0 = clock64();
      uint32_t spins = 0;
      while (ld_acquire_u32(bar + 1) == gen) {
        if ((++spins & 0x3ff) == 0) ds_trap_if(clock64() - t0 > 6000000000LL);
      }
    }
    __threadfence();
  }
  vec_sync();
}

// ---------------------------------------------------------------- row-wise phases (vector warps)
// LayerNorm with pending split-K slabs: arithmetic of layernorm_reduce_kernel (kernels.cu) -- "thread" t of that
// kernel owns float4 t of the row; here 128 threads own up to 3 float4 each (t, t+128, t+256), which keeps every
// float4 in the same lane of the same 32-wide group, so the shuffle trees and the group-order sums are identical.
__device__ __forceinline__ void ds_phase_ln(const DsPhase& p, int vt, float (*s_red)[2][12]) {
  const int half = vt >> 7, t = vt & 127, lane = t & 31, w = t >> 5;
  const int nf4 = p.d >> 2, n_groups = nf4 >> 5;  // float4 per row, 32-wide groups
  const long long st4 = p.slab_stride >> 2;
  for (int row = blockIdx.x * 2 + half; row < p.rows; row += gridDim.x * 2) {
    float4* xp = reinterpret_cast<float4*>(p.x + (long long)row * p.d);
    float4 v[3];
    float sum[3];
    if (p.n_slabs > 0) {
      // All loads of the row are issued before the first store: x and the slabs may alias as far as the compiler
      // knows, so a store between two groups would serialise their loads (measured: 7.4 -> ~2.5 us per phase).
      const float4* pp = reinterpret_cast<const float4*>(p.slabs + (long long)row * p.d);
      float4 acc[3][4];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[j][u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t + 128 * j < nf4) v[j] = xp[t + 128 * j];
      }
      int s = 0;
      for (; s + 4 <= p.n_slabs; s += 4) {
        float4 q[3][4];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (t + 128 * j < nf4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) q[j][u] = pp[(s + u) * st4 + t + 128 * j];
          }
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (t + 128 * j < nf4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              acc[j][u].x += q[j][u].x; acc[j][u].y += q[j][u].y; acc[j][u].z += q[j][u].z; acc[j][u].w += q[j][u].w;
            }
          }
      }
      for (; s < p.n_slabs; ++s) {
        float4 q[3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (t + 128 * j < nf4) q[j] = pp[s * st4 + t + 128 * j];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (t + 128 * j < nf4) { acc[j][0].x += q[j].x; acc[j][0].y += q[j].y; acc[j][0].z += q[j].z; acc[j][0].w += q[j].w; }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int idx = t + 128 * j;
        if (idx < nf4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias) + idx);
          v[j].x += bb.x + ((acc[j][0].x + acc[j][1].x) + (acc[j][2].x + acc[j][3].x));
          v[j].y += bb.y + ((acc[j][0].y + acc[j][1].y) + (acc[j][2].y + acc[j][3].y));
          v[j].z += bb.z + ((acc[j][0].z + acc[j][1].z) + (acc[j][2].z + acc[j][3].z));
          v[j].w += bb.w + ((acc[j][0].w + acc[j][1].w) + (acc[j][2].w + acc[j][3].w));
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (t + 128 * j < nf4) xp[t + 128 * j] = v[j];
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (t + 128 * j < nf4) v[j] = xp[t + 128 * j];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) sum[j] = t + 128 * j < nf4 ? v[j].x + v[j].y + v[j].z + v[j].w : 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      for (int o = 16; o > 0; o >>= 1) sum[j] += __shfl_xor_sync(0xffffffffu, sum[j], o);
      if (lane == 0 && w + 4 * j < n_groups) s_red[half][0][w + 4 * j] = sum[j];
    }
    half_sync(half);
    float tot = 0.f;
    for (int g = 0; g < n_groups; ++g) tot += s_red[half][0][g];
    const float mean = tot / (float)p.d;
    float sq[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      sq[j] = 0.f;
      if (t + 128 * j < nf4) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
        sq[j] = a * a + b * b + c * c + e * e;
      }
      for (int o = 16; o > 0; o >>= 1) sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], o);
      if (lane == 0 && w + 4 * j < n_groups) s_red[half][1][w + 4 * j] = sq[j];
    }
    half_sync(half);
    float tsq = 0.f;
    for (int g = 0; g < n_groups; ++g) tsq += s_red[half][1][g];
    const float rstd = rsqrtf(tsq / (float)p.d + 1e-5f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int idx = t + 128 * j;
      if (idx < nf4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + idx);
        const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta) + idx);
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
        __nv_bfloat162 lo = __floats2bfloat162_rn(a * rstd * g.x + be.x, b * rstd * g.y + be.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn(c * rstd * g.z + be.z, e * rstd * g.w + be.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&lo);
        u.y = *reinterpret_cast<uint32_t*>(&hi);
        reinterpret_cast<uint2*>(p.y + (long long)row * p.d)[idx] = u;
      }
    }
    half_sync(half);  // s_red is rewritten by the next row
  }
}

// out = bf16(act(bias + sum of slabs)): arithmetic of split_finalize_kernel (kernels.cu)
__device__ __forceinline__ void ds_phase_fin(const DsPhase& p, int vt) {
  const int nf4 = p.N >> 2;
  const long long total = (long long)p.rows * nf4;
  for (long long i = (long long)blockIdx.x * DS_VEC_THREADS + vt; i < total; i += (long long)gridDim.x * DS_VEC_THREADS) {
    const int row = (int)(i / nf4), c = (int)(i - (long long)row * nf4) * 4;
    float4 a = __ldg(reinterpret_cast<const float4*>(p.bias + c));
    const float* pp = p.slabs + (long long)row * p.N + c;
#pragma unroll 4
    for (int s = 0; s < p.n_slabs; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(pp + s * p.slab_stride);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    if (p.act == 1) { a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w); }
    __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p.out + (long long)row * p.N + c) = u;
  }
}

// x[r] = tok_emb[tok[r]] + pos_emb[pos[r]] (embed_kernel, kernels.cu)
__device__ __forceinline__ void ds_phase_embed(const DsPhase& p, int vt) {
  const int half = vt >> 7, t = vt & 127;
  for (int row = blockIdx.x * 2 + half; row < p.rows; row += gridDim.x * 2) {
    const bf16* te = p.tok_emb + (long long)p.tok[row] * p.d;
    const float* pe = p.pos_emb + (long long)p.pos[row] * p.d;
    for (int c = t * 2; c < p.d; c += 256) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(te + c));
      const float2 b = *reinterpret_cast<const float2*>(pe + c);
      *reinterpret_cast<float2*>(p.x + (long long)row * p.d + c) = make_float2(a.x + b.x, a.y + b.y);
    }
  }
}

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(DS_THREADS, 1) dstep_kernel(const DsLaunch* __restrict__ L) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ DsPhase ph[DS_MAX_PHASES];
  __shared__ __align__(8) uint64_t bars[3 * DS_MAX_STAGES + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_red[2][2][12];
  __shared__ int hdr[4];
  __shared__ unsigned go_count;  // grid barriers this CTA has passed (vector thread 0 -> X producer)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(L->ph);
    for (int i = tid; i < (int)(sizeof(DsPhase) * DS_MAX_PHASES / 4); i += DS_THREADS)
      reinterpret_cast<uint32_t*>(ph)[i] = __ldg(src + i);
    if (tid < 4) hdr[tid] = __ldg(&L->n_phases + tid);
  }
  const uint32_t fullW0 = smem_u32(bars), fullX0 = fullW0 + 8 * DS_MAX_STAGES, empty0 = fullX0 + 8 * DS_MAX_STAGES;
  const uint32_t accF0 = empty0 + 8 * DS_MAX_STAGES;  // [2] accumulator b is complete
  const uint32_t accE0 = accF0 + 16;                  // [2] accumulator b has been drained (8 warp arrivals)
  if (tid == 64) {
    for (int s = 0; s < 3 * DS_MAX_STAGES + 2; ++s) mbar_init(fullW0 + 8 * s, 1);
    mbar_init(accE0, 8);
    mbar_init(accE0 + 8, 8);
    go_count = 0;
    fence_barrier_init();
  }
  pdl_trigger();
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int n_phases = hdr[0], stages = hdr[1], stage_bytes = hdr[2];
  const uint32_t tmem_base = tmem_slot;
  const int G = gridDim.x, cta = blockIdx.x;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- W producer (runs ahead of the grid barriers)
      uint32_t stage = 0, phase = 0;
      for (int pi = 0; pi < n_phases; ++pi) {
        const DsPhase& p = ph[pi];
        if (p.type != DS_GEMM) continue;
        const CUtensorMap* tmW = &L->tm[2 * pi];
        tma_prefetch_desc(tmW);
        const int items = p.n_tiles * p.splits;
        for (int it = cta; it < items; it += G) {
          const int ks = it / p.n_tiles, nt = it - ks * p.n_tiles;
          const int kb_lo = (int)((long long)p.total_kb * ks / p.splits);
          const int n_kb = (int)((long long)p.total_kb * (ks + 1) / p.splits) - kb_lo;
          for (int kb = 0; kb < n_kb; ++kb) {
            ds_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t fb = fullW0 + 8 * stage;
            mbar_arrive_expect_tx(fb, DS_W_BYTES);
            tma_load_3d(smem_u32(smem + stage * stage_bytes), tmW, fb, (kb_lo + kb) * DS_BK, nt * 128, 0);
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- X producer (activations: after the barrier that completes them)
      uint32_t stage = 0, phase = 0;
      pdl_wait();
      for (int pi = 0; pi < n_phases; ++pi) {
        if (pi > 0) {  // a counter, not an mbarrier parity: the vector warps may be several barriers ahead
          const long long t0 = clock64();
          uint32_t spins = 0;
          while (*reinterpret_cast<volatile unsigned*>(&go_count) < (unsigned)pi) {
            if ((++spins & 0x3ff) == 0) ds_trap_if(clock64() - t0 > 6000000000LL);
          }
          __threadfence();
          fence_proxy_async_global();
        }
        const DsPhase& p = ph[pi];
        if (p.type != DS_GEMM) continue;
        const CUtensorMap* tmX = &L->tm[2 * pi + 1];
        const int x_bytes = p.Mpad * 128;
        const int items = p.n_tiles * p.splits;
        for (int it = cta; it < items; it += G) {
          const int ks = it / p.n_tiles;
          const int kb_lo = (int)((long long)p.total_kb * ks / p.splits);
          const int n_kb = (int)((long long)p.total_kb * (ks + 1) / p.splits) - kb_lo;
          for (int kb = 0; kb < n_kb; ++kb) {
            ds_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t fb = fullX0 + 8 * stage;
            mbar_arrive_expect_tx(fb, x_bytes);
            tma_load_3d(smem_u32(smem + stage * stage_bytes + DS_W_BYTES), tmX, fb, (kb_lo + kb) * DS_BK, 0, 0);
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {  // ---------------- MMA issuer
      uint32_t stage = 0, phase = 0, acc = 0, acc_par[2] = {0, 0};
      for (int pi = 0; pi < n_phases; ++pi) {
        const DsPhase& p = ph[pi];
        if (p.type != DS_GEMM) continue;
        const uint32_t idesc = make_idesc_bf16(128, p.Mpad, 0, 0);
        const int items = p.n_tiles * p.splits;
        for (int it = cta; it < items; it += G) {
          const int ks = it / p.n_tiles;
          const int kb_lo = (int)((long long)p.total_kb * ks / p.splits);
          const int n_kb = (int)((long long)p.total_kb * (ks + 1) / p.splits) - kb_lo;
          ds_wait(accE0 + 8 * acc, acc_par[acc] ^ 1);  // the epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 256;
          for (int kb = 0; kb < n_kb; ++kb) {
            ds_wait(fullW0 + 8 * stage, phase);
            ds_wait(fullX0 + 8 * stage, phase);
            tc_fence_after();
            const uint64_t wd = make_sw128_desc(smem_u32(smem + stage * stage_bytes));
            const uint64_t xd = make_sw128_desc(smem_u32(smem + stage * stage_bytes + DS_W_BYTES));
#pragma unroll
            for (int k = 0; k < DS_BK / 16; ++k) tc_mma_bf16(d_tmem, wd + 2 * k, xd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit(empty0 + 8 * stage);
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
          }
          tc_commit(accF0 + 8 * acc);
          acc_par[acc] ^= 1;
          acc ^= 1;
        }
      }
    }
  } else {  // ---------------- vector warps: epilogues, row-wise phases, grid barriers
    const int vt = tid - DS_VEC_WARP0 * 32;
    long long* trace = L->trace;
    DS_TRACE(0);
    pdl_wait();
    DS_TRACE(1);
    uint32_t acc = 0, acc_par[2] = {0, 0};
    for (int pi = 0; pi < n_phases; ++pi) {
      const DsPhase& p = ph[pi];
      if (p.type == DS_GEMM) {
        const int q = warp & 3, half = (warp - DS_VEC_WARP0) >> 2;
        const int items = p.n_tiles * p.splits;
        for (int it = cta; it < items; it += G) {
          const int ks = it / p.n_tiles, nt = it - ks * p.n_tiles;
          const int n = nt * 128 + q * 32 + lane;
          const bool n_ok = n < p.N;
          ds_wait(accF0 + 8 * acc, acc_par[acc]);
          tc_fence_after();
          const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
          const long long ldo = p.ldo;
          for (int c0 = half * 32; c0 < p.Mpad; c0 += 64) {
            uint32_t v[32];
            tmem_ld32(taddr + c0, v);
            tmem_ld_wait();
            if (!n_ok) continue;
            float* op = p.slabs + (long long)ks * p.slab_stride + n + (long long)c0 * ldo;
            if (c0 + 32 <= p.rows) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                *op = __uint_as_float(v[j]);
                op += ldo;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (c0 + j < p.rows) *op = __uint_as_float(v[j]);
                op += ldo;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(accE0 + 8 * acc);
          acc_par[acc] ^= 1;
          acc ^= 1;
        }
      } else if (p.type == DS_LN) {
        ds_phase_ln(p, vt, s_red);
      } else if (p.type == DS_FIN) {
        ds_phase_fin(p, vt);
      } else {
        ds_phase_embed(p, vt);
      }
      DS_TRACE(2 + 2 * pi);
      if (pi + 1 < n_phases) {
        ds_grid_barrier(L->bar, vt);
        DS_TRACE(3 + 2 * pi);
        if (vt == 0) {
          __threadfence();
          atomicAdd(&go_count, 1u);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- host side
struct DstepRecord {
  DsLaunch* dev = nullptr;
  std::vector<unsigned char> desc_bytes;  // the descriptors the record was built from
  size_t smem = 0;
};
static long long* g_ds_trace = nullptr;
static int g_ds_trace_key = -1;
void dstep_set_trace(void* buf, int key) { g_ds_trace = static_cast<long long*>(buf); g_ds_trace_key = key; }
struct DstepCache {
  std::mutex mu;
  std::map<int, DstepRecord> recs;
  unsigned* bar = nullptr;
};
DstepCache* dstep_cache_create() { return new DstepCache(); }
void dstep_cache_destroy(DstepCache* c) {
  if (!c) return;
  if (c->bar) cudaFree(c->bar);
  delete c;
}

bool dstep_supported(int rows, int d) { return rows >= 1 && rows <= 256 && d % 128 == 0 && d <= 1536; }

void dstep_run(DstepCache* cache, int key, const DsPhaseDesc* phases, int n_phases, cudaStream_t stream) {
  B200W_REQUIRE(cache && phases && n_phases >= 1 && n_phases <= DS_MAX_PHASES, "dstep_run: bad program");
  std::lock_guard<std::mutex> lock(cache->mu);
  if (!cache->bar) {
    B200W_CHECK_CUDA(cudaMalloc(&cache->bar, 256));
    B200W_CHECK_CUDA(cudaMemset(cache->bar, 0, 256));
  }
  const size_t nbytes = sizeof(DsPhaseDesc) * (size_t)n_phases;
  DstepRecord& rec = cache->recs[key];
  if (!rec.dev || rec.desc_bytes.size() != nbytes || memcmp(rec.desc_bytes.data(), phases, nbytes) != 0) {
    DsLaunch host;
    memset(&host, 0, sizeof(host));
    const int rows = phases[0].rows;
    B200W_REQUIRE(rows >= 1 && rows <= 256, "dstep_run: 1..256 activation rows");
    const int Mpad = (rows + 15) & ~15;
    for (int i = 0; i < n_phases; ++i) {
      const DsPhaseDesc& d = phases[i];
      B200W_REQUIRE(d.rows == rows, "dstep_run: every phase must have the same number of rows");
      DsPhase& p = host.ph[i];
      p.type = d.type;
      p.rows = rows;
      p.Mpad = Mpad;
      p.n_slabs = d.n_slabs;
      p.d = d.d;
      p.act = d.act;
      p.slabs = d.slabs;
      p.bias = d.bias;
      p.x = d.x; p.gamma = d.gamma; p.beta = d.beta; p.y = d.y; p.out = d.out;
      p.tok = d.tok; p.pos = d.pos; p.tok_emb = d.tok_emb; p.pos_emb = d.pos_emb;
      p.N = d.N;
      if (d.type == DS_GEMM) {
        B200W_REQUIRE(d.X && d.W && d.slabs && d.K % DS_BK == 0 && d.K > 0 && d.N > 0, "dstep_run: bad GEMM phase");
        p.K = d.K;
        p.splits = d.splits > 1 ? d.splits : 1;
        p.total_kb = d.K / DS_BK;
        B200W_REQUIRE(p.total_kb >= p.splits, "dstep_run: more splits than k-blocks");
        p.n_tiles = cdiv(d.N, 128);
        p.ldo = d.ldo > 0 ? d.ldo : d.N;
        p.slab_stride = (long long)rows * p.ldo;
        make_tmap_bf16_3d(&host.tm[2 * i], d.W, d.K, d.N, 1, d.K, (uint64_t)d.N * d.K, DS_BK, 128);
        make_tmap_bf16_3d(&host.tm[2 * i + 1], d.X, d.K, rows, 1, d.K, (uint64_t)rows * d.K, DS_BK, Mpad);
      } else if (d.type == DS_LN) {
        B200W_REQUIRE(d.x && d.gamma && d.beta && d.y && d.d % 128 == 0 && d.d <= 1536, "dstep_run: bad LayerNorm phase");
        B200W_REQUIRE(d.n_slabs == 0 || (d.slabs && d.bias), "dstep_run: LayerNorm slabs need a bias");
        p.slab_stride = (long long)rows * d.d;
      } else if (d.type == DS_FIN) {
        B200W_REQUIRE(d.slabs && d.bias && d.out && d.N % 4 == 0 && d.n_slabs >= 1, "dstep_run: bad finalize phase");
        p.slab_stride = (long long)rows * d.N;
      } else {
        B200W_REQUIRE(d.type == DS_EMBED && d.tok && d.pos && d.tok_emb && d.pos_emb && d.x && d.d % 2 == 0,
                      "dstep_run: bad embedding phase");
      }
    }
    host.n_phases = n_phases;
    host.stage_bytes = DS_W_BYTES + Mpad * 128;
    host.stages = (200 * 1024) / host.stage_bytes;
    if (host.stages > DS_MAX_STAGES) host.stages = DS_MAX_STAGES;
    host.bar = cache->bar;
    host.trace = (key % 128) == g_ds_trace_key ? g_ds_trace : nullptr;
    // (a record is only rebuilt when its call site changes shape, never inside a captured decode step; a rebuilt
    // record gets a NEW address: descriptors are cached by address)
    rec.dev = static_cast<DsLaunch*>(alloc_descriptor_record(sizeof(DsLaunch)));
    B200W_CHECK_CUDA(cudaMemcpyAsync(rec.dev, &host, sizeof(host), cudaMemcpyHostToDevice, stream));
    B200W_CHECK_CUDA(cudaStreamSynchronize(stream));
    rec.desc_bytes.assign(reinterpret_cast<const unsigned char*>(phases), reinterpret_cast<const unsigned char*>(phases) + nbytes);
    rec.smem = (size_t)host.stages * host.stage_bytes + 1024;
  }
  ensure_dynamic_smem(dstep_kernel, rec.smem);
  launch_pdl(dstep_kernel, dim3(num_sms()), dim3(DS_THREADS), rec.smem, stream, static_cast<const DsLaunch*>(rec.dev));
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

}  // namespace b200w

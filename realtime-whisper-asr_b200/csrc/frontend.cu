// Front end for the Whisper hot path (HBM-bound work, CUDA cores only):
//   logmel_kernel        reflect-pad framing + Hann window + 400-pt real FFT (as a 200-pt complex
//                        FFT, radix 8 x 5 x 5 through shared memory) + |.|^2 + sparse slaney mel
//                        filterbank + log10, and the per-call running max (for the max-8 clamp)
//   logmel_finalize      max(x, gmax-8), (x+4)/4 in place
//   pack_windows         fp32 [n_mels][frames] -> bf16 time-major [3000][128] encoder windows
//                        (zero-padded FEATURES, faster-whisper pad_or_trim semantics)
//   dedup_features       whole-chunk real DFT of arbitrary (smooth) length in shared memory +
//                        RMS / centroid / ZCR / roll-off / bandwidth, max-abs normalised
// Replaces faster-whisper FeatureExtractor.__call__ (NumPy) and
// /root/reference/audio_deduplicator.py:60-181.
#include <math.h>

#include <mutex>
#include <vector>

#include "common.h"
#include "frontend.h"

namespace b200w {

static constexpr int NFFT = 400, HOP = 160, NBIN = 201;
static constexpr int FT = 32;                       // frames per CTA tile
static constexpr int TILE_SAMPLES = FT * HOP + 240;  // 5360
static constexpr int PSTRIDE = 201;                 // odd -> conflict-free

struct MelTables {
  float window[NFFT];
  float2 tw200[8 * 25];  // [k1][n2] = exp(-2 pi i n2 k1 / 200)
  float2 tw25[5 * 5];    // [q1][m2] = exp(-2 pi i m2 q1 / 25)
  float2 tw400[101];     // exp(-2 pi i k / 400), k = 0..100
  int mel_lo[128];
  int mel_cnt[128];
  int mel_woff[128];
  float mel_w[512];
  int n_mels;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// forward 5-point DFT in place
__device__ __forceinline__ void dft5(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  const float2 t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
  const float2 m1 = make_float2(a0.x + c1 * t1.x + c2 * t2.x, a0.y + c1 * t1.y + c2 * t2.y);
  const float2 m2 = make_float2(a0.x + c2 * t1.x + c1 * t2.x, a0.y + c2 * t1.y + c1 * t2.y);
  const float2 u1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
  const float2 u2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
  a0 = make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
  a1 = make_float2(m1.x + u1.y, m1.y - u1.x);  // m1 - i u1
  a4 = make_float2(m1.x - u1.y, m1.y + u1.x);  // m1 + i u1
  a2 = make_float2(m2.x + u2.y, m2.y - u2.x);
  a3 = make_float2(m2.x - u2.y, m2.y + u2.x);
}

// forward 8-point DFT: y[k] = sum_n a[n] exp(-2 pi i n k / 8)
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
  const float h = 0.70710678118654752f;
  // radix-2 DIT: evens / odds 4-point DFTs
  float2 e0 = cadd(a[0], a[4]), e1 = csub(a[0], a[4]), e2 = cadd(a[2], a[6]), e3 = csub(a[2], a[6]);
  float2 E0 = cadd(e0, e2), E2 = csub(e0, e2);
  float2 E1 = make_float2(e1.x + e3.y, e1.y - e3.x);  // e1 - i e3
  float2 E3 = make_float2(e1.x - e3.y, e1.y + e3.x);  // e1 + i e3
  float2 o0 = cadd(a[1], a[5]), o1 = csub(a[1], a[5]), o2 = cadd(a[3], a[7]), o3 = csub(a[3], a[7]);
  float2 O0 = cadd(o0, o2), O2 = csub(o0, o2);
  float2 O1 = make_float2(o1.x + o3.y, o1.y - o3.x);
  float2 O3 = make_float2(o1.x - o3.y, o1.y + o3.x);
  // twiddles W8^k: 1, (1-i)h, -i, (-1-i)h
  float2 T1 = make_float2(h * (O1.x + O1.y), h * (O1.y - O1.x));
  float2 T2 = make_float2(O2.y, -O2.x);
  float2 T3 = make_float2(h * (O3.y - O3.x), -h * (O3.x + O3.y));
  a[0] = cadd(E0, O0); a[4] = csub(E0, O0);
  a[1] = cadd(E1, T1); a[5] = csub(E1, T1);
  a[2] = cadd(E2, T2); a[6] = csub(E2, T2);
  a[3] = cadd(E3, T3); a[7] = csub(E3, T3);
}

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF);
}

// index into the (audio zero-padded by `pad` samples) signal with numpy "reflect" boundary rule
__device__ __forceinline__ long long reflect_index(long long g, long long L) {
  if (L == 1) return 0;
  const long long period = 2 * (L - 1);
  g %= period;
  if (g < 0) g += period;
  return g < L ? g : period - g;
}

struct MelArgs {
  const float* audio;            // concatenated samples
  const long long* offsets;      // [n_streams + 1] sample offsets
  const long long* out_offsets;  // [n_streams] element offsets into out
  const int* tile_prefix;        // [n_streams + 1] cumulative tile counts
  float* out;                    // per stream [n_mels][n_frames]
  int* stream_max;               // [n_streams] ordered-int running max (init INT_MIN)
  int n_streams;
};

__global__ void __launch_bounds__(256) logmel_kernel(const MelTables* __restrict__ tab, MelArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* s_x = reinterpret_cast<float*>(smem_raw);                 // TILE_SAMPLES
  float2* s_z = reinterpret_cast<float2*>(s_x + TILE_SAMPLES);      // FT * 200
  float* s_p = reinterpret_cast<float*>(s_z + FT * 200);            // FT * PSTRIDE
  float* s_win = s_p + FT * PSTRIDE;                                // 400
  float2* s_tw200 = reinterpret_cast<float2*>(s_win + NFFT);        // 200
  float2* s_tw25 = s_tw200 + 200;                                   // 25
  float2* s_tw400 = s_tw25 + 25;                                    // 101 (+1 pad)
  float* s_melw = reinterpret_cast<float*>(s_tw400 + 102);          // 512
  int* s_mello = reinterpret_cast<int*>(s_melw + 512);              // 128
  int* s_melcnt = s_mello + 128;
  int* s_melwoff = s_melcnt + 128;
  __shared__ int s_blockmax;

  const int tid = threadIdx.x;
  // locate (stream, first frame) of this tile: binary search the cumulative tile table
  int lo = 0, hi = a.n_streams;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a.tile_prefix[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const int s = lo;
  const int f0 = ((int)blockIdx.x - a.tile_prefix[s]) * FT;
  const long long off = a.offsets[s];
  const long long N = a.offsets[s + 1] - off;
  const int n_frames = (int)(N / HOP) + 1;
  const long long Lz = N + HOP;  // audio zero-padded by one hop
  const int n_mels = tab->n_mels;

  // constant tables: global -> shared with cp.async as well (they share the wait below with the samples)
  auto cp4 = [](void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
  };
  auto cp8 = [](void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
  };
  for (int i = tid; i < NFFT; i += 256) cp4(s_win + i, tab->window + i);
  for (int i = tid; i < 200; i += 256) cp8(s_tw200 + i, tab->tw200 + i);
  if (tid < 25) cp8(s_tw25 + tid, tab->tw25 + tid);
  if (tid < 101) cp8(s_tw400 + tid, tab->tw400 + tid);
  for (int i = tid; i < 512; i += 256) cp4(s_melw + i, tab->mel_w + i);
  if (tid < 128) {
    cp4(s_mello + tid, tab->mel_lo + tid);
    cp4(s_melcnt + tid, tab->mel_cnt + tid);
    cp4(s_melwoff + tid, tab->mel_woff + tid);
  }
  if (tid == 0) s_blockmax = float_to_ordered(-INFINITY);

  // samples of the tile (coalesced), boundary rules resolved here
  const long long g0 = (long long)f0 * HOP - NFFT / 2;
  // In-range samples go global -> shared with cp.async (no register round trip, so all ~21 requests of a
  // thread are in flight together: the first version waited on DRAM once per 4 samples); only the
  // reflected / zero-padded borders of a clip take the scalar path.
  for (int i = tid; i < TILE_SAMPLES; i += 256) {
    const long long g = g0 + i;
    if (g >= 0 && g < N) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_x + i);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(a.audio + off + g) : "memory");
    } else {
      float v = 0.f;
      const long long r = reflect_index(g, Lz);
      if (r < N) v = __ldg(a.audio + off + r);
      s_x[i] = v;
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  const int nf = min(FT, n_frames - f0);  // valid frames in this tile

  // stage 1: window, pack even/odd samples as complex, 8-point DFTs over n1 (stride 25), twiddle
  for (int task = tid; task < nf * 25; task += 256) {
    const int f = task / 25, n2 = task - f * 25;
    float2 v[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const int n = 25 * n1 + n2;
      const float2 x = *reinterpret_cast<const float2*>(s_x + f * HOP + 2 * n);
      const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * n);
      v[n1] = make_float2(x.x * w.x, x.y * w.y);
    }
    dft8(v);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1)
      s_z[f * 200 + k1 * 25 + n2] = k1 == 0 ? v[0] : cmul(v[k1], s_tw200[k1 * 25 + n2]);
  }
  __syncthreads();

  // stage 2: 25-point DFT (5 x 5) over n2 for each k1, in place: row k1 now holds Z[k1 + 8 k2]
  for (int task = tid; task < nf * 8; task += 256) {
    float2* row = s_z + task * 25;  // task = f * 8 + k1
    float2 c[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) c[i] = row[i];  // c[5 m1 + m2]
#pragma unroll
    for (int m2 = 0; m2 < 5; ++m2) dft5(c[m2], c[5 + m2], c[10 + m2], c[15 + m2], c[20 + m2]);
    // c[5 q1 + m2] *= W25^(m2 q1)
#pragma unroll
    for (int q1 = 1; q1 < 5; ++q1)
#pragma unroll
      for (int m2 = 1; m2 < 5; ++m2) c[5 * q1 + m2] = cmul(c[5 * q1 + m2], s_tw25[q1 * 5 + m2]);
#pragma unroll
    for (int q1 = 0; q1 < 5; ++q1)
      dft5(c[5 * q1], c[5 * q1 + 1], c[5 * q1 + 2], c[5 * q1 + 3], c[5 * q1 + 4]);
    // c[5 q1 + q2] = B[k2 = q1 + 5 q2]
#pragma unroll
    for (int q1 = 0; q1 < 5; ++q1)
#pragma unroll
      for (int q2 = 0; q2 < 5; ++q2) row[q1 + 5 * q2] = c[5 * q1 + q2];
  }
  __syncthreads();

  // real-input split: X[k] = E[k] + W400^k O[k]; power spectrum for k and 200-k
  for (int task = tid; task < nf * 101; task += 256) {
    const int f = task / 101, k = task - f * 101;
    const int km = (200 - k) % 200;
    const float2 zk = s_z[f * 200 + (k & 7) * 25 + (k >> 3)];
    float2 zm = s_z[f * 200 + (km & 7) * 25 + (km >> 3)];
    zm.y = -zm.y;  // conj
    const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
    const float2 D = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y - zm.y));
    const float2 O = make_float2(D.y, -D.x);  // D / i
    const float2 WO = cmul(s_tw400[k], O);
    const float2 Xk = cadd(E, WO), Xm = csub(E, WO);
    s_p[f * PSTRIDE + k] = Xk.x * Xk.x + Xk.y * Xk.y;
    s_p[f * PSTRIDE + 200 - k] = Xm.x * Xm.x + Xm.y * Xm.y;
  }
  __syncthreads();

  // sparse mel filterbank + log10; frame index fastest so the stores coalesce along time
  float* out = a.out + a.out_offsets[s];
  float lmax = -INFINITY;
  for (int task = tid; task < n_mels * FT; task += 256) {
    const int m = task / FT, f = task - m * FT;
    if (f < nf) {
      const float* p = s_p + f * PSTRIDE + s_mello[m];
      const float* w = s_melw + s_melwoff[m];
      float acc = 0.f;
      for (int i = 0; i < s_melcnt[m]; ++i) acc = fmaf(w[i], p[i], acc);
      const float v = log10f(fmaxf(acc, 1e-10f));
      out[(long long)m * n_frames + f0 + f] = v;
      lmax = fmaxf(lmax, v);
    }
  }
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if ((tid & 31) == 0) atomicMax(&s_blockmax, float_to_ordered(lmax));
  __syncthreads();
  if (tid == 0) atomicMax(a.stream_max + s, s_blockmax);
}

__global__ void init_stream_max_kernel(int* m, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) m[i] = float_to_ordered(-INFINITY);
}

// in place: x = (max(x, gmax - 8) + 4) / 4.  grid.y = stream
__global__ void logmel_finalize_kernel(float* out, const long long* out_offsets,
                                       const long long* offsets, const int* stream_max, int n_mels) {
  const int s = blockIdx.y;
  const long long N = offsets[s + 1] - offsets[s];
  const long long total = (long long)n_mels * (N / HOP + 1);
  const float floor_v = ordered_to_float(stream_max[s]) - 8.0f;
  float* p = out + out_offsets[s];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = (fmaxf(p[i], floor_v) + 4.0f) / 4.0f;
}

// windows[w][t][c] (bf16, 3000 x 128) = features[c][seek + t] or 0.  grid = (94, 4, n_windows)
__global__ void pack_windows_kernel(const float* feats, const long long* feat_offsets,
                                    const int* n_frames, const int* seeks, const int* win_stream,
                                    int n_mels, bf16* windows) {
  __shared__ float tile[32][33];
  const int w = blockIdx.z;
  const int s = win_stream[w];
  const int nfr = n_frames[s], seek = seeks[w];
  const float* src = feats + feat_offsets[s];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, t = t0 + threadIdx.x;
    float v = 0.f;
    if (c < n_mels && t < 3000 && seek + t < nfr) v = src[(long long)c * nfr + seek + t];
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int t = t0 + j, c = c0 + threadIdx.x;
    if (t < 3000) windows[((long long)w * 3000 + t) * 128 + c] = __float2bfloat16(tile[threadIdx.x][j]);
  }
}

// ------------------------------------------------------------------------------ dedup features
struct FftPlan {
  int n_stages;
  int radix[24];
};

// generic in-place DIT butterfly of radix r over elements base + j * stride, twiddle W_L^(j k0)
template <int R>
__device__ __forceinline__ void butterfly_fixed(float2* z, int base, int stride, int k0, int L) {
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2 x = z[base + j * stride];
    if (j > 0 && k0 > 0) {
      const int idx = (int)(((long long)j * k0) % L);
      float sn, cs;
      sincospif(-2.0f * (float)idx / (float)L, &sn, &cs);
      x = cmul(x, make_float2(cs, sn));
    }
    v[j] = x;
  }
  if (R == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else if (R == 4) {
    const float2 e0 = cadd(v[0], v[2]), e1 = csub(v[0], v[2]);
    const float2 o0 = cadd(v[1], v[3]), o1 = csub(v[1], v[3]);
    v[0] = cadd(e0, o0);
    v[2] = csub(e0, o0);
    v[1] = make_float2(e1.x + o1.y, e1.y - o1.x);
    v[3] = make_float2(e1.x - o1.y, e1.y + o1.x);
  } else if (R == 5) {
    dft5(v[0], v[1], v[2], v[3], v[4]);
  } else {  // R == 3
    const float c = -0.5f, sn = 0.86602540378443865f;
    const float2 t1 = cadd(v[1], v[2]), t2 = csub(v[1], v[2]);
    const float2 m = make_float2(v[0].x + c * t1.x, v[0].y + c * t1.y);
    const float2 u = make_float2(sn * t2.x, sn * t2.y);
    v[0] = cadd(v[0], t1);
    v[1] = make_float2(m.x + u.y, m.y - u.x);
    v[2] = make_float2(m.x - u.y, m.y + u.x);
  }
#pragma unroll
  for (int j = 0; j < R; ++j) z[base + j * stride] = v[j];
}

static constexpr int RMAX = 64;
__device__ void butterfly_generic(float2* z, int base, int stride, int k0, int L, int r) {
  float2 v[RMAX], y[RMAX];
  for (int j = 0; j < r; ++j) {
    float2 x = z[base + j * stride];
    if (j > 0 && k0 > 0) {
      const int idx = (int)(((long long)j * k0) % L);
      float sn, cs;
      sincospif(-2.0f * (float)idx / (float)L, &sn, &cs);
      x = cmul(x, make_float2(cs, sn));
    }
    v[j] = x;
  }
  for (int q = 0; q < r; ++q) {
    float2 acc = v[0];
    for (int j = 1; j < r; ++j) {
      float sn, cs;
      sincospif(-2.0f * (float)((j * q) % r) / (float)r, &sn, &cs);
      acc = cadd(acc, cmul(v[j], make_float2(cs, sn)));
    }
    y[q] = acc;
  }
  for (int q = 0; q < r; ++q) z[base + q * stride] = y[q];
}

struct DedupArgs {
  const float* audio;
  const long long* offsets;  // [n_chunks + 1]
  float* out;                // [n_chunks][5]
  float* raw_out;            // optional [n_chunks][5] un-normalised (may be null)
  int sample_rate;
};

// one CTA per chunk; the chunk lives in shared memory as M complex points
__global__ void __launch_bounds__(512) dedup_features_kernel(DedupArgs a, int max_points) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float2* z = reinterpret_cast<float2*>(smem_raw);
  __shared__ double s_red[5][16];
  __shared__ double s_scan[16];
  __shared__ int s_roll;
  __shared__ FftPlan plan;

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int c = blockIdx.x;
  const long long off = a.offsets[c];
  const int N = (int)(a.offsets[c + 1] - off);
  float* out = a.out + (long long)c * 5;
  if (N == 0) {
    if (tid < 5) {
      out[tid] = 0.f;
      if (a.raw_out) a.raw_out[(long long)c * 5 + tid] = 0.f;
    }
    return;
  }
  const bool even = (N & 1) == 0;
  const int M = even ? N / 2 : N;
  const float* x = a.audio + off;

  if (tid == 0) {  // factorise M: radix 4 first, then 2, 3, 5, then remaining primes
    int m = M, ns = 0;
    bool ok = true;
    while (m % 4 == 0) { plan.radix[ns++] = 4; m /= 4; }
    while (m % 2 == 0) { plan.radix[ns++] = 2; m /= 2; }
    while (m % 3 == 0) { plan.radix[ns++] = 3; m /= 3; }
    while (m % 5 == 0) { plan.radix[ns++] = 5; m /= 5; }
    for (int p = 7; m > 1 && ok; p += 2) {
      while (m % p == 0) {
        if (p > RMAX || ns >= 24) { ok = false; break; }
        plan.radix[ns++] = p;
        m /= p;
      }
      if (p > RMAX && m > 1) ok = false;
    }
    plan.n_stages = ok ? ns : -1;
    s_roll = 0x7fffffff;
  }
  __syncthreads();
  const int nbins = N / 2 + 1;
  const bool direct = plan.n_stages < 0;  // a prime factor > RMAX: exact O(N^2) DFT instead
  const bool fits = direct ? ((long long)N + nbins) * 4 <= (long long)max_points * 8 : M <= max_points;
  if (!fits) {  // host sized the launch wrongly: NaN marks it (host raises)
    if (tid < 5) out[tid] = __int_as_float(0x7fc00000);
    return;
  }

  // --- time-domain sums (double)
  double sumsq = 0.0;
  long long zc = 0;
  for (int i = tid; i < N; i += nthr) {
    const float v = __ldg(x + i);
    sumsq += (double)v * (double)v;
    if (i + 1 < N) zc += (signbit(v) != signbit(__ldg(x + i + 1))) ? 1 : 0;
  }
  float* mag;
  if (direct) {
    float* xs = reinterpret_cast<float*>(smem_raw);
    mag = xs + N;
    for (int i = tid; i < N; i += nthr) xs[i] = __ldg(x + i);
    __syncthreads();
    for (int k = tid; k < nbins; k += nthr) {
      double re = 0.0, im = 0.0;
      int idx = 0;  // (n * k) mod N, kept exact
      for (int n = 0; n < N; ++n) {
        float sn, cs;
        sincospif(-2.0f * (float)idx / (float)N, &sn, &cs);
        re += (double)xs[n] * (double)cs;
        im += (double)xs[n] * (double)sn;
        idx += k;
        if (idx >= N) idx -= N;
      }
      mag[k] = (float)sqrt(re * re + im * im);
    }
    __syncthreads();
  } else {
    // digit-reversed load: position of input n under mixed-radix DIT (first radix = outermost)
    for (int n = tid; n < M; n += nthr) {
      int rem = n, pos = 0, blk = M;
      for (int st = 0; st < plan.n_stages; ++st) {
        const int r = plan.radix[st];
        blk /= r;
        pos += (rem % r) * blk;
        rem /= r;
      }
      z[pos] = even ? make_float2(__ldg(x + 2 * n), __ldg(x + 2 * n + 1)) : make_float2(__ldg(x + n), 0.f);
    }
    __syncthreads();

    // in-place decimation-in-time passes, innermost (last radix) first
    int Lprev = 1;
    for (int st = plan.n_stages - 1; st >= 0; --st) {
      const int r = plan.radix[st];
      const int L = Lprev * r;
      const int nb = M / r;  // butterflies this pass
      for (int t = tid; t < nb; t += nthr) {
        const int k0 = t % Lprev;
        const int base = (t / Lprev) * L + k0;
        if (r == 4) butterfly_fixed<4>(z, base, Lprev, k0, L);
        else if (r == 2) butterfly_fixed<2>(z, base, Lprev, k0, L);
        else if (r == 5) butterfly_fixed<5>(z, base, Lprev, k0, L);
        else if (r == 3) butterfly_fixed<3>(z, base, Lprev, k0, L);
        else butterfly_generic(z, base, Lprev, k0, L, r);
      }
      Lprev = L;
      __syncthreads();
    }

    // magnitudes of the N/2+1 real-FFT bins, stored over z: every thread first computes its bins
    // into registers (strided ownership), then all store after a barrier.
    mag = reinterpret_cast<float*>(smem_raw);
    float vals[64];
    int cnt = 0;
    for (int k = tid; k < nbins; k += nthr) {
      float2 X;
      if (even) {
        const float2 zk = z[k % M];
        float2 zm = z[(M - k) % M];
        zm.y = -zm.y;
        const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
        const float2 D = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y - zm.y));
        const float2 O = make_float2(D.y, -D.x);
        float sn, cs;
        sincospif(-2.0f * (float)k / (float)N, &sn, &cs);
        X = cadd(E, cmul(make_float2(cs, sn), O));
      } else {
        X = z[k];
      }
      vals[cnt++] = sqrtf(X.x * X.x + X.y * X.y);
    }
    __syncthreads();
    cnt = 0;
    for (int k = tid; k < nbins; k += nthr) mag[k] = vals[cnt++];
    __syncthreads();
  }

  // --- spectral sums in double: S0 = sum mag, S1 = sum f mag
  const double df = (double)a.sample_rate / (double)N;
  double s0 = 0.0, s1 = 0.0;
  for (int k = tid; k < nbins; k += nthr) {
    const double m = (double)mag[k];
    s0 += m;
    s1 += (double)k * df * m;
  }
  // block reduce helpers
  auto block_sum = [&](double v, int slot) -> double {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s_red[slot][tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (nthr >> 5); ++w) t += s_red[slot][w];
    __syncthreads();
    return t;
  };
  const double S0 = block_sum(s0, 0);
  const double S1 = block_sum(s1, 1);
  const double SSQ = block_sum(sumsq, 2);
  const double ZC = block_sum((double)zc, 3);

  const double nyq = (double)a.sample_rate / 2.0;
  const double centroid = S1 / (S0 + 1e-10) / nyq;

  // --- roll-off: first k with cumsum(mag)[k] >= 0.85 * total (contiguous ownership + block scan)
  const int per = (nbins + nthr - 1) / nthr;
  const int kb = tid * per, ke = min(nbins, kb + per);
  double local = 0.0;
  for (int k = kb; k < ke; ++k) local += (double)mag[k];
  // exclusive prefix over threads: warp scan then warp totals
  double incl = local;
  for (int o = 1; o < 32; o <<= 1) {
    const double n = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += n;
  }
  if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
  __syncthreads();
  double wbase = 0.0;
  for (int w = 0; w < (tid >> 5); ++w) wbase += s_scan[w];
  double run = wbase + incl - local;
  const double thr = 0.85 * S0;
  if (S0 > 1e-10) {
    for (int k = kb; k < ke; ++k) {
      run += (double)mag[k];
      if (run >= thr) {
        atomicMin(&s_roll, k);
        break;
      }
    }
  }
  // --- bandwidth
  double bw_acc = 0.0;
  const double cf = centroid * nyq;
  for (int k = tid; k < nbins; k += nthr) {
    const double d = (double)k * df - cf;
    bw_acc += d * d * (double)mag[k];
  }
  const double BW = block_sum(bw_acc, 4);  // also orders s_roll
  if (tid == 0) {
    const float rms = sqrtf((float)(SSQ / (double)N));
    const double zcr = N > 1 ? ZC / (double)N : 0.0;
    double rolloff;
    if (S0 > 1e-10) rolloff = s_roll == 0x7fffffff ? 1.0 : ((double)s_roll * df) / nyq;
    else rolloff = 0.0;
    const double bw = centroid > 0 ? sqrt(BW / (S0 + 1e-10)) / nyq : 0.0;
    float f[5] = {rms, (float)centroid, (float)zcr, (float)rolloff, (float)bw};
    if (a.raw_out)
      for (int i = 0; i < 5; ++i) a.raw_out[(long long)c * 5 + i] = f[i];
    float mx = 0.f;
    for (int i = 0; i < 5; ++i) mx = fmaxf(mx, fabsf(f[i]));
    for (int i = 0; i < 5; ++i) out[i] = mx > 1e-10f ? f[i] / mx : 0.f;
  }
}

// ------------------------------------------------------------------------------------- host side
static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static void build_tables(MelTables& t, int n_mels) {
  const double PI = 3.14159265358979323846;
  for (int i = 0; i < NFFT; ++i) t.window[i] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / NFFT));
  for (int k1 = 0; k1 < 8; ++k1)
    for (int n2 = 0; n2 < 25; ++n2) {
      const double ang = -2.0 * PI * (double)(n2 * k1) / 200.0;
      t.tw200[k1 * 25 + n2] = make_float2((float)cos(ang), (float)sin(ang));
    }
  for (int q1 = 0; q1 < 5; ++q1)
    for (int m2 = 0; m2 < 5; ++m2) {
      const double ang = -2.0 * PI * (double)(m2 * q1) / 25.0;
      t.tw25[q1 * 5 + m2] = make_float2((float)cos(ang), (float)sin(ang));
    }
  for (int k = 0; k <= 100; ++k) {
    const double ang = -2.0 * PI * (double)k / 400.0;
    t.tw400[k] = make_float2((float)cos(ang), (float)sin(ang));
  }
  // slaney mel filterbank (librosa.filters.mel, norm="slaney"), kept sparse
  std::vector<double> mel_f(n_mels + 2);
  const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(8000.0);
  for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
  int woff = 0;
  t.n_mels = n_mels;
  for (int m = 0; m < 128; ++m) t.mel_lo[m] = t.mel_cnt[m] = t.mel_woff[m] = 0;
  for (int m = 0; m < n_mels; ++m) {
    const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
    int lo = -1, cnt = 0;
    for (int k = 0; k < NBIN; ++k) {
      const double fk = 16000.0 / NFFT * k;
      const double lower = (fk - mel_f[m]) / (mel_f[m + 1] - mel_f[m]);
      const double upper = (mel_f[m + 2] - fk) / (mel_f[m + 2] - mel_f[m + 1]);
      const double w = fmax(0.0, fmin(lower, upper)) * enorm;
      const float wf = (float)w;
      if (wf != 0.f) {
        if (lo < 0) lo = k;
        // filters are contiguous triangles: k - lo == cnt
        if (woff + cnt >= 512) throw Error("mel filterbank has more non-zeros than expected");
        t.mel_w[woff + cnt] = wf;
        ++cnt;
      }
    }
    t.mel_lo[m] = lo < 0 ? 0 : lo;
    t.mel_cnt[m] = cnt;
    t.mel_woff[m] = woff;
    woff += cnt;
  }
}

static const MelTables* device_tables(int n_mels) {
  // one copy per (device, n_mels): a table uploaded to GPU 0 is not addressable from GPU 1's kernels
  static MelTables* tabs[64][2] = {};
  static std::mutex mu;
  B200W_REQUIRE(n_mels == 80 || n_mels == 128, "n_mels must be 80 or 128");
  int dev = 0;
  B200W_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  MelTables** slot = &tabs[dev & 63][n_mels == 80 ? 0 : 1];
  if (!*slot) {
    MelTables* h = new MelTables();
    build_tables(*h, n_mels);
    B200W_CHECK_CUDA(cudaMalloc(slot, sizeof(MelTables)));
    B200W_CHECK_CUDA(cudaMemcpy(*slot, h, sizeof(MelTables), cudaMemcpyHostToDevice));
    delete h;
  }
  return *slot;
}

static size_t logmel_smem_bytes() {
  return sizeof(float) * TILE_SAMPLES + sizeof(float2) * FT * 200 + sizeof(float) * FT * PSTRIDE +
         sizeof(float) * NFFT + sizeof(float2) * (200 + 25 + 102) + sizeof(float) * 512 +
         sizeof(int) * 3 * 128;
}

void logmel_launch(const float* audio, const long long* offsets, const long long* out_offsets,
                   const int* tile_prefix, int n_streams, int total_tiles, int n_mels, float* out,
                   int* stream_max, int max_frames, cudaStream_t stream) {
  if (n_streams <= 0) return;
  const MelTables* tab = device_tables(n_mels);
  const size_t smem = logmel_smem_bytes();
  ensure_dynamic_smem(logmel_kernel, smem);
  init_stream_max_kernel<<<cdiv(n_streams, 256), 256, 0, stream>>>(stream_max, n_streams);
  B200W_COUNT_LAUNCH();
  MelArgs a{audio, offsets, out_offsets, tile_prefix, out, stream_max, n_streams};
  logmel_kernel<<<total_tiles, 256, smem, stream>>>(tab, a);
  B200W_COUNT_LAUNCH();
  const long long per_stream = (long long)n_mels * max_frames;
  int gx = (int)((per_stream + 1023) / 1024);
  const int cap = cdiv(8 * num_sms(), n_streams);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  logmel_finalize_kernel<<<dim3(gx, n_streams), 256, 0, stream>>>(out, out_offsets, offsets,
                                                                 stream_max, n_mels);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

void pack_windows_launch(const float* feats, const long long* feat_offsets, const int* n_frames,
                         const int* seeks, const int* win_stream, int n_windows, int n_mels,
                         bf16* windows, cudaStream_t stream) {
  if (n_windows <= 0) return;
  pack_windows_kernel<<<dim3(94, 4, n_windows), dim3(32, 8), 0, stream>>>(
      feats, feat_offsets, n_frames, seeks, win_stream, n_mels, windows);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

void dedup_features_launch(const float* audio, const long long* offsets, int n_chunks,
                           int max_points, int sample_rate, float* out, float* raw_out,
                           cudaStream_t stream) {
  if (n_chunks <= 0) return;
  // max_points = max over chunks of (N even ? N/2 : N): the chunk lives in shared memory as that
  // many complex points (8 bytes each); 216 KB is what one CTA can take on sm_100a.
  B200W_REQUIRE(max_points <= 27648, "chunk too long for the shared-memory DFT (N <= 55296 even, 27648 odd)");
  size_t smem = (size_t)(max_points < 128 ? 128 : max_points) * 8;
  ensure_dynamic_smem(dedup_features_kernel, smem);
  DedupArgs a{audio, offsets, out, raw_out, sample_rate};
  dedup_features_kernel<<<n_chunks, 512, smem, stream>>>(a, (int)(smem / 8));
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------- band-pass pre-filter
// Zero-phase IIR band-pass of every chunk: scipy.signal.filtfilt(b, a, x) with its defaults (odd extension by
// padlen = 3 * max(len(a), len(b)) samples at both ends, steady-state initial conditions zi * first sample, forward
// pass, time-reversed pass), which the reference applies to each incoming chunk before the dedup check and the ASR
// buffer (SimpleFilterSeparator.separate, /root/reference/vocal_separation.py:335-365: 4th-order Butterworth,
// 85-3400 Hz, i.e. 9 + 9 coefficients; call site 一键实时识别麦克风.py:1474-1482).
// The recurrence is sequential in time, chunks are independent: one thread per chunk, 128 chunks per CTA; the samples
// go through shared memory in 32-sample pieces so that global traffic is coalesced (a warp moves one chunk's 128
// bytes per instruction) although a thread walks its chunk serially.  Arithmetic in fp64 registers (direct form II
// transposed, the same structure scipy's lfilter uses), storage fp32.  Forward pass writes y into `out`, the
// reversed pass rewrites it in place (position i only depends on y[i..]).  Chunks not longer than padlen are passed
// through unchanged (scipy raises there, the reference catches it and keeps the original audio).
static constexpr int BP_ORDER = 8;            // len(a) - 1 for the 4th-order band-pass
static constexpr int BP_PAD = 3 * (BP_ORDER + 1);
static constexpr int BP_CHUNKS = 128;         // chunks (= threads) per CTA
static constexpr int BP_PIECE = 32;           // samples per chunk per staged piece
struct BandpassCoeffs {
  double b[BP_ORDER + 1], a[BP_ORDER + 1], zi[BP_ORDER];
};
struct BandpassArgs {
  const float* x;
  const long long* offsets;  // [n_chunks + 1]
  float* out;
  float* rpad;               // [n_chunks][BP_PAD] forward-pass outputs over the right extension
  int n_chunks;
};

__device__ __forceinline__ double bp_step(double xin, double (&z)[BP_ORDER], const BandpassCoeffs& c) {
  const double y = fma(c.b[0], xin, z[0]);
#pragma unroll
  for (int i = 0; i < BP_ORDER - 1; ++i) z[i] = fma(c.b[i + 1], xin, fma(-c.a[i + 1], y, z[i + 1]));
  z[BP_ORDER - 1] = fma(c.b[BP_ORDER], xin, -c.a[BP_ORDER] * y);
  return y;
}

__global__ void __launch_bounds__(BP_CHUNKS) bandpass_filtfilt_kernel(const BandpassArgs g, const BandpassCoeffs c) {
  __shared__ float tile[BP_CHUNKS][BP_PIECE + 1];
  __shared__ float hist[BP_CHUNKS][BP_PAD + 2];  // last BP_PAD + 1 inputs of every chunk (circular)
  __shared__ long long s_off[BP_CHUNKS + 1];
  __shared__ int s_maxn;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int c0 = blockIdx.x * BP_CHUNKS;
  const int nc = min(BP_CHUNKS, g.n_chunks - c0);
  if (t == 0) s_maxn = 0;
  for (int i = t; i <= nc; i += BP_CHUNKS) s_off[i] = g.offsets[c0 + i];
  __syncthreads();
  const long long off = t < nc ? s_off[t] : 0;
  const int N = t < nc ? (int)(s_off[t + 1] - s_off[t]) : 0;
  atomicMax(&s_maxn, N);
  __syncthreads();
  const int maxn = s_maxn;
  const bool filt = N > BP_PAD;  // shorter chunks: passthrough
  double z[BP_ORDER];
  // ---------------- forward pass
  const int n_pieces = (maxn + BP_PIECE - 1) / BP_PIECE;
  for (int p = 0; p < n_pieces; ++p) {
    for (int r = warp; r < nc; r += BP_CHUNKS / 32) {  // coalesced: one warp, one chunk row
      const int n = (int)(s_off[r + 1] - s_off[r]), i = p * BP_PIECE + lane;
      tile[r][lane] = i < n ? g.x[s_off[r] + i] : 0.f;
    }
    __syncthreads();
    if (t < nc) {
      const int base = p * BP_PIECE;
      if (filt && p == 0) {
        // left extension 2 x[0] - x[BP_PAD - i], i = 0 .. BP_PAD-1 (N > BP_PAD >= piece 0 holds x[1..27])
        const double x0 = tile[t][0];
        const double e0 = 2.0 * x0 - (double)tile[t][BP_PAD];
#pragma unroll
        for (int i = 0; i < BP_ORDER; ++i) z[i] = c.zi[i] * e0;
        for (int i = 0; i < BP_PAD; ++i) bp_step(2.0 * x0 - (double)tile[t][BP_PAD - i], z, c);
      }
      const int cnt = min(BP_PIECE, N - base);
      for (int i = 0; i < cnt; ++i) {
        const float xv = tile[t][i];
        hist[t][(base + i) % (BP_PAD + 2)] = xv;
        tile[t][i] = filt ? (float)bp_step((double)xv, z, c) : xv;
      }
      if (filt && cnt > 0 && base + cnt == N) {
        // right extension 2 x[N-1] - x[N-2-j]: forward outputs kept for the start of the reversed pass
        const double xl = hist[t][(N - 1) % (BP_PAD + 2)];
        float* rp = g.rpad + (long long)(c0 + t) * BP_PAD;
        for (int j = 0; j < BP_PAD; ++j) rp[j] = (float)bp_step(2.0 * xl - (double)hist[t][(N - 2 - j) % (BP_PAD + 2)], z, c);
      }
    }
    __syncthreads();
    for (int r = warp; r < nc; r += BP_CHUNKS / 32) {
      const int n = (int)(s_off[r + 1] - s_off[r]), i = p * BP_PIECE + lane;
      if (i < n) g.out[s_off[r] + i] = tile[r][lane];
    }
    __syncthreads();
  }
  // ---------------- reversed pass over [right extension, y[N-1] .. y[0]] (the left extension's outputs are dropped)
  if (filt) {
    const float* rp = g.rpad + (long long)(c0 + t) * BP_PAD;
    const double e0 = rp[BP_PAD - 1];
#pragma unroll
    for (int i = 0; i < BP_ORDER; ++i) z[i] = c.zi[i] * e0;
    for (int j = BP_PAD - 1; j >= 0; --j) bp_step((double)rp[j], z, c);
  }
  for (int p = n_pieces - 1; p >= 0; --p) {
    for (int r = warp; r < nc; r += BP_CHUNKS / 32) {
      const int n = (int)(s_off[r + 1] - s_off[r]), i = p * BP_PIECE + lane;
      tile[r][lane] = i < n ? g.out[s_off[r] + i] : 0.f;
    }
    __syncthreads();
    if (t < nc && filt) {
      const int base = p * BP_PIECE;
      const int cnt = min(BP_PIECE, N - base);
      for (int i = cnt - 1; i >= 0; --i) tile[t][i] = (float)bp_step((double)tile[t][i], z, c);
    }
    __syncthreads();
    for (int r = warp; r < nc; r += BP_CHUNKS / 32) {
      const int n = (int)(s_off[r + 1] - s_off[r]), i = p * BP_PIECE + lane;
      if (i < n) g.out[s_off[r] + i] = tile[r][lane];
    }
    __syncthreads();
  }
  (void)off;
}

void bandpass_filtfilt_launch(const float* audio, const long long* offsets, int n_chunks, const double* coeffs_host,
                              float* out, float* rpad_scratch, cudaStream_t stream) {
  if (n_chunks <= 0) return;
  BandpassCoeffs c;
  memcpy(c.b, coeffs_host, sizeof(double) * (BP_ORDER + 1));
  memcpy(c.a, coeffs_host + BP_ORDER + 1, sizeof(double) * (BP_ORDER + 1));
  memcpy(c.zi, coeffs_host + 2 * (BP_ORDER + 1), sizeof(double) * BP_ORDER);
  B200W_REQUIRE(fabs(c.a[0] - 1.0) < 1e-12, "bandpass_filtfilt: a[0] must be 1 (normalised coefficients)");
  BandpassArgs g{audio, offsets, out, rpad_scratch, n_chunks};
  bandpass_filtfilt_kernel<<<cdiv(n_chunks, BP_CHUNKS), BP_CHUNKS, 0, stream>>>(g, c);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------- batched VAD gate
// The reference gates every stream through DynamicVADIterator (/root/reference/asr_components.py:12-78): incoming
// audio is re-chunked to 512-sample windows, a speech probability p is computed per window (Silero on the CPU), and
// upstream's VADIterator state machine turns the p sequence into {start} / {end} events (start when p >= threshold
// and not triggered; end when p < threshold - 0.15 for at least min_silence_samples).  For hundreds of
// streams: one launch scores all windows of all streams, a second one runs every stream's state machine (one thread
// per stream, windows in order) and merges the events of a call exactly as FixedVADIterator.__call__ does.
// vad_window_energy_kernel is the built-in scorer p = rms / (rms + k) (the energy stand-in the tests use for the
// un-vendored Silero network); any other scorer just writes the same probs buffer.
__global__ void vad_window_energy_kernel(const float* __restrict__ windows, int n_windows, float k,
                                         float* __restrict__ probs) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= n_windows) return;
  const float4* src = reinterpret_cast<const float4*>(windows + (long long)w * 512);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // 128 float4 per window, coalesced
    const float4 v = src[lane + 32 * i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float rms = sqrtf(acc / 512.0f);
    probs[w] = rms / (rms + k);
  }
}

// state per stream: [triggered, temp_end, current_sample] (int64), min_silence_samples (double: the reference stores
// sampling_rate * ms / 1000 as a float).  events per stream per call: [has_start, start, has_end, end].
__global__ void vad_gate_kernel(const float* __restrict__ probs, const int* __restrict__ win_offsets, int n_streams,
                                long long* __restrict__ state, const double* __restrict__ min_silence,
                                float threshold, long long speech_pad, long long* __restrict__ events) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  long long triggered = state[3 * s], temp_end = state[3 * s + 1], cur = state[3 * s + 2];
  const double min_sil = min_silence[s];
  bool has_ret = false, has_start = false, has_end = false;  // `ret` of FixedVADIterator.__call__
  long long start = 0, end = 0;
  for (int w = win_offsets[s]; w < win_offsets[s + 1]; ++w) {
    const float p = probs[w];
    cur += 512;
    bool r_start = false, r_end = false;
    long long rs = 0, re = 0;
    if (p >= threshold && temp_end) temp_end = 0;
    if (p >= threshold && !triggered) {
      triggered = 1;
      r_start = true;
      rs = cur - speech_pad - 512;
    } else if (p < threshold - 0.15f && triggered) {
      if (!temp_end) temp_end = cur;
      if (!((double)(cur - temp_end) < min_sil)) {
        r_end = true;
        re = temp_end + speech_pad - 512;
        temp_end = 0;
        triggered = 0;
      }
    }
    if (!has_ret) {  // ret is None: take r as it is (possibly None)
      if (r_start || r_end) {
        has_ret = true;
        has_start = r_start; start = rs;
        has_end = r_end; end = re;
      }
    } else if (r_start || r_end) {
      if (r_end) { has_end = true; end = re; }                      // the latter end
      if (r_start && has_end) has_end = false;                      // a new start merges the segments
    }
  }
  state[3 * s] = triggered; state[3 * s + 1] = temp_end; state[3 * s + 2] = cur;
  events[4 * s] = has_start; events[4 * s + 1] = start; events[4 * s + 2] = has_end; events[4 * s + 3] = end;
}

void vad_window_energy_launch(const float* windows, int n_windows, float k, float* probs, cudaStream_t stream) {
  if (n_windows <= 0) return;
  vad_window_energy_kernel<<<cdiv(n_windows, 8), 256, 0, stream>>>(windows, n_windows, k, probs);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}
void vad_gate_launch(const float* probs, const int* win_offsets, int n_streams, long long* state,
                     const double* min_silence, float threshold, long long speech_pad, long long* events,
                     cudaStream_t stream) {
  if (n_streams <= 0) return;
  vad_gate_kernel<<<cdiv(n_streams, 128), 128, 0, stream>>>(probs, win_offsets, n_streams, state, min_silence, threshold,
                                                            speech_pad, events);
  B200W_COUNT_LAUNCH();
  B200W_CHECK_CUDA(cudaGetLastError());
}

}  // namespace b200w

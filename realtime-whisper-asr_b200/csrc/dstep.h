// Fused decode-step segments: one persistent kernel runs a short PROGRAM of dependent phases (LayerNorm with
// split-K reduction, swap-AB tcgen05 GEMM, split-K finalize, token embedding) with grid-wide barriers between
// them instead of kernel launches, and streams the weights of later phases while earlier phases run.  See dstep.cu.
#pragma once
#include "common.h"

namespace b200w {

enum DsType : int { DS_LN = 0, DS_GEMM = 1, DS_FIN = 2, DS_EMBED = 3 };

// Host-side description of one phase.  All pointers are device pointers.
struct DsPhaseDesc {
  int type = DS_LN;
  int rows = 0;  // activation rows (<= 256), the same for every phase of a program
  // DS_GEMM: slabs[ks][r][n] = sum over k in range ks of X[r][k] W[n][k]   (ks < splits; fp32, row stride ldo)
  const bf16* X = nullptr;
  const bf16* W = nullptr;
  int N = 0, K = 0, splits = 1;
  float* slabs = nullptr;  // DS_GEMM: output; DS_LN / DS_FIN: input ([n_slabs][rows][width])
  long long ldo = 0;
  // DS_LN: x[r] += bias + sum_s slabs[s][r]  (n_slabs == 0: x unchanged);  y[r] = bf16(LayerNorm(x[r]) * gamma + beta)
  int n_slabs = 0;
  const float* bias = nullptr;
  float* x = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  bf16* y = nullptr;
  int d = 0;
  // DS_FIN: out[r][c] = bf16(act(bias[c] + sum_s slabs[s][r][c])), c < N; act 1 = exact-erf GELU
  bf16* out = nullptr;
  int act = 0;
  // DS_EMBED: x[r] = tok_emb[tok[r]] + pos_emb[pos[r]]   (width d)
  const int* tok = nullptr;
  const int* pos = nullptr;
  const bf16* tok_emb = nullptr;
  const float* pos_emb = nullptr;
};

static constexpr int DS_MAX_PHASES = 8;

struct DstepCache;  // launch records (tensor maps + programs in device memory) and the grid-barrier words of one engine
DstepCache* dstep_cache_create();
void dstep_cache_destroy(DstepCache* c);

// Run the program.  `key` names the call site (a program launched again with the same key and the same descriptors
// re-uses its device record, so decode steps -- which repeat the same programs on the same buffers -- build nothing).
void dstep_run(DstepCache* cache, int key, const DsPhaseDesc* phases, int n_phases, cudaStream_t stream);
bool dstep_supported(int rows, int d);
// development aid: programs whose key % 128 == key (built after this call) stamp [cta][24] globaltimer values into buf
void dstep_set_trace(void* dev_buf, int key);

}  // namespace b200w

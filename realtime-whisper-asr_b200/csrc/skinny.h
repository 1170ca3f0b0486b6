// Swap-AB tcgen05 GEMM for decode steps (at most 256 activation rows); see skinny.cu.
#pragma once
#include "common.h"

namespace b200w {

// Launch records (TMA descriptors + parameters, in device memory) of the distinct calls made through one owner.
// An engine owns one and frees it when it is destroyed; calls without an owner share a bounded process-wide one.
struct SkinnyCache;
SkinnyCache* skinny_cache_create();
void skinny_cache_destroy(SkinnyCache* c);  // caller guarantees no launch that uses the records is in flight

struct SkinnyDesc {
  const bf16* X = nullptr;  // activations [M][K], row stride ldx (elements)
  int64_t ldx = 0;
  int M = 0;
  const bf16* W = nullptr;  // weights [N][K] (torch Linear layout), row stride ldw
  int64_t ldw = 0;
  int N = 0, K = 0;
  const float* bias = nullptr;  // [N] or null (must be null with splits > 1)
  int act = 0;                  // 1 = exact-erf GELU after the bias
  void* out = nullptr;          // [M][ldo] bf16 | fp32, or fp32 slabs [splits][M][ldo]
  int out_fp32 = 0;
  int64_t ldo = 0;
  int splits = 1;
  int tile_rows = 128;  // weight rows per CTA: 128 or 64 (64 needs M >= 64)
  int64_t split_stride = 0;  // elements between slabs
  SkinnyCache* cache = nullptr;  // owner of the launch record (null: the shared default)
};
void skinny_gemm(const SkinnyDesc& d, cudaStream_t stream);
void skinny_set_trace(void* dev_buf);  // development aid: per-CTA timestamps of later launches
// K splits that give ceil(N/128) weight-row tiles about one wave of CTAs (1 = do not split)
int skinny_suggest_splits(int N, int K, int max_splits, int tile_rows = 128);

}  // namespace b200w

// bf16 tcgen05 GEMM with fused epilogues, also used as implicit-GEMM conv1d (time-major taps).
#pragma once
#include "common.h"

namespace b200w {

// One tap of an implicit-GEMM conv: the A tile for these k-blocks is read from
// rows (m + row_shift), columns [col_off, col_off + 64*kblocks).  Out-of-range rows read as zero.
struct GemmTap {
  int row_shift;
  int col_off;
  int kblocks;
};

struct GemmDesc {
  // A: bf16 [batches][a_rows][a_cols], K (columns) contiguous.
  const bf16* A = nullptr;
  int a_rows = 0, a_cols = 0, batches = 1;
  int64_t a_row_stride = 0, a_batch_stride = 0;  // elements
  // W: bf16 [N][K] row-major (torch Linear layout); out = A * W^T
  const bf16* W = nullptr;
  int N = 0, K = 0;
  int64_t w_row_stride = 0;
  int n_taps = 0;  // 0: plain GEMM (one tap covering K)
  GemmTap taps[3];
  int out_rows = 0;  // produced rows per batch
  const float* bias = nullptr;  // [N] fp32 or null
  int act = 0;                  // 0 none, 1 exact-erf GELU (applied after bias, before residual)
  const void* residual = nullptr;  // added last; bf16 or fp32
  int res_fp32 = 0;
  int64_t res_row_stride = 0, res_batch_stride = 0;
  void* out = nullptr;  // bf16 or fp32
  int out_fp32 = 0;
  int64_t out_row_stride = 0, out_batch_stride = 0;  // elements; row stride % 8 == 0
  int force_bn = 0;  // 0 = heuristic; else 64 / 128 / 256
  // Head-major scatter of the output (used for the decoder's cross K/V): rows are (stream, t) with
  // head_major_T positions per stream, columns are (slab, head, 64); element goes to
  // out[stream][slab][head][t][64].  0 = plain row-major output.
  int head_major_T = 0, head_major_d = 0;
  // split-K (n > 1): split s stores its fp32 partial sums at out + s * split_stride (plain
  // stores, no bias / activation / residual); the consumer adds the slabs in order (deterministic).
  int split_k = 1;
  int64_t split_stride = 0;
};

void gemm_bf16(const GemmDesc& d, cudaStream_t stream);

// conv1d, kernel 3, padding 1, stride 1|2 over time-major x bf16 [B][T][C]; w bf16 [N][3C] tap-major;
// out = GELU(conv + bias) + addend (fp32 [T_out][N], shared by all batches; may be null).
void conv1d_k3(const bf16* x, const bf16* w, const float* bias, const float* addend, void* out, int B,
               int T, int C, int N, int stride, int out_fp32, cudaStream_t stream);

// Per-launch CUDA-event timing of one tile-width class (64 / 128 / 256); see bench.py.
void gemm_profile_begin(int bn);
void gemm_profile_end(double* total_ms, double* total_flops, long long* launches);

// Development aid: every CTA of later GEMM launches writes 16 int64 timestamps to dev_buf (null = off).
void gemm_set_trace(void* dev_buf);

// Build a 128B-swizzled TMA descriptor over a bf16 tensor viewed as [d2][d1][d0] (d0 contiguous).
void make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1);

}  // namespace b200w

// What does one dependent kernel launch cost inside a CUDA graph on B200, as a function of what the
// kernel's prologue does?  (development aid; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_empty(float* out) { if (threadIdx.x == 0 && out[0] == 123.f) out[1] = 1.f; }

__global__ void k_smem(float* out) {
  extern __shared__ uint8_t sm[];
  if (threadIdx.x == 0) sm[0] = 1;
  __syncthreads();
  if (threadIdx.x == 0 && out[0] == 123.f) out[1] = sm[0];
}

__global__ void k_bar(float* out) {
  extern __shared__ uint8_t sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 11; ++i) {
      uint32_t a = (uint32_t)__cvta_generic_to_shared(bars + i);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && out[0] == 123.f) out[1] = 1.f;
}

__global__ void k_tmem(float* out) {
  extern __shared__ uint8_t sm[];
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm);
  if (threadIdx.x < 32) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(a) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t base = *slot;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base) : "memory");
  if (threadIdx.x == 0 && out[0] == 123.f) out[1] = 1.f;
}

__global__ void k_pdl(float* out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0 && out[0] == 123.f) out[1] = 1.f;
}
__global__ void k_ldg(float* out, const uint32_t* rec) {
  __shared__ uint32_t sh[32];
  if (threadIdx.x < 32) sh[threadIdx.x] = __ldg(rec + threadIdx.x);
  __syncthreads();
  if (threadIdx.x == 0 && sh[3] == 123u) out[1] = 1.f;
}
__global__ void k_tmap_param(float* out, const __grid_constant__ CUtensorMap tm) {
  if (threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
  if (threadIdx.x == 0 && out[0] == 123.f) out[1] = 1.f;
}
__global__ void k_tmap_global(float* out, const CUtensorMap* tm) {
  if (threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
  if (threadIdx.x == 0 && out[0] == 123.f) out[1] = 1.f;
}
// one 16 KB TMA tile load into smem, wait for it
__global__ void k_tma(float* out, const __grid_constant__ CUtensorMap tm) {
  extern __shared__ uint8_t sm[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 16384);
  uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(base);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(16384) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(d), "l"(&tm), "r"(b), "r"(0), "r"((int)blockIdx.x * 128), "r"(0) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(0) : "memory");
    if (out[0] == 123.f) out[1] = base[5];
  }
}

// straight-line code of N dependent FMAs per thread (code size = 16 B x N), executed once
template <int N>
__global__ void k_code(float* out, float a) {
  float x[8] = {a, a + 1, a + 2, a + 3, a + 4, a + 5, a + 6, a + 7};  // 8 independent chains: issue-bound, not latency-bound
#pragma unroll
  for (int i = 0; i < N; ++i) x[i & 7] = fmaf(x[i & 7], a, 1.0f + i);
  if (x[0] + x[1] + x[2] + x[3] + x[4] + x[5] + x[6] + x[7] == 123.f) out[1] = x[0];
}

// every thread stores n floats, coalesced
__global__ void k_store(float* out, int n) {
  for (int j = 0; j < n; ++j) out[1024 + (size_t)j * 4096 + blockIdx.x * blockDim.x + threadIdx.x] = (float)j;
}

template <typename F>
static float time_chain(F launch, cudaStream_t s, int L = 32, int reps = 20) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < L; ++i) launch();
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
  for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  return ms * 1000.f / (reps * L);
}

int main() {
  float* out;
  CK(cudaMalloc(&out, 64 << 20));
  CK(cudaMemset(out, 0, 64 << 20));
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  const int big = 200 * 1024;
  CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(k_bar, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  uint32_t* rec;
  CK(cudaMalloc(&rec, 32 * 4096 * 4));
  CK(cudaMemset(rec, 0, 32 * 4096 * 4));
  // a 2-D bf16 tensor [148*128 rows][64] for the TMA tests
  void* tdata;
  CK(cudaMalloc(&tdata, 148 * 128 * 128));
  CK(cudaMemset(tdata, 0, 148 * 128 * 128));
  typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
  CUtensorMap tm;
  {
    cuuint64_t dims[3] = {64, 148 * 128, 1};
    cuuint64_t strides[2] = {128, 148 * 128 * 128};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((PFN)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, tdata, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  CUtensorMap* tm_dev;
  CK(cudaMalloc(&tm_dev, sizeof(tm)));
  CK(cudaMemcpy(tm_dev, &tm, sizeof(tm), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024));
  for (int grid : {1, 148}) {
    printf("grid %d\n", grid);
    printf("  empty, 32 threads            %6.2f us\n", time_chain([&] { k_empty<<<grid, 32, 0, s>>>(out); }, s));
    printf("  empty, 320 threads           %6.2f us\n", time_chain([&] { k_empty<<<grid, 320, 0, s>>>(out); }, s));
    printf("  smem 1 KB                    %6.2f us\n", time_chain([&] { k_smem<<<grid, 320, 1024, s>>>(out); }, s));
    printf("  smem 72 KB                   %6.2f us\n", time_chain([&] { k_smem<<<grid, 320, 72 * 1024, s>>>(out); }, s));
    printf("  smem 200 KB                  %6.2f us\n", time_chain([&] { k_smem<<<grid, 320, big, s>>>(out); }, s));
    printf("  200 KB + 11 mbarriers        %6.2f us\n", time_chain([&] { k_bar<<<grid, 320, big, s>>>(out); }, s));
    printf("  200 KB + TMEM alloc/dealloc  %6.2f us\n", time_chain([&] { k_tmem<<<grid, 320, big, s>>>(out); }, s));
    printf("  1 coalesced store / thread   %6.2f us\n", time_chain([&] { k_store<<<grid, 256, 0, s>>>(out, 1); }, s));
    printf("  32 coalesced stores / thread %6.2f us\n", time_chain([&] { k_store<<<grid, 256, 0, s>>>(out, 32); }, s));
    printf("  96 coalesced stores / thread %6.2f us\n", time_chain([&] { k_store<<<grid, 256, 0, s>>>(out, 96); }, s));
    printf("  griddepcontrol trigger+wait  %6.2f us\n", time_chain([&] { k_pdl<<<grid, 320, 0, s>>>(out); }, s));
    printf("  ldg 128 B record (same)      %6.2f us\n", time_chain([&] { k_ldg<<<grid, 320, 0, s>>>(out, rec); }, s));
    { int i = 0; printf("  ldg 128 B record (distinct)  %6.2f us\n", time_chain([&] { k_ldg<<<grid, 320, 0, s>>>(out, rec + 4096 * (i++ % 32)); }, s)); }
    printf("  prefetch.tensormap (param)   %6.2f us\n", time_chain([&] { k_tmap_param<<<grid, 320, 0, s>>>(out, tm); }, s));
    printf("  prefetch.tensormap (global)  %6.2f us\n", time_chain([&] { k_tmap_global<<<grid, 320, 0, s>>>(out, tm_dev); }, s));
    printf("  one 16 KB TMA load + wait    %6.2f us\n", time_chain([&] { k_tma<<<grid, 320, 20 * 1024, s>>>(out, tm); }, s));
    printf("  straight-line  256 instr     %6.2f us\n", time_chain([&] { k_code<256><<<grid, 32, 0, s>>>(out, 0.5f); }, s));
    printf("  straight-line 1024 instr     %6.2f us\n", time_chain([&] { k_code<1024><<<grid, 32, 0, s>>>(out, 0.5f); }, s));
    printf("  straight-line 4096 instr     %6.2f us\n", time_chain([&] { k_code<4096><<<grid, 32, 0, s>>>(out, 0.5f); }, s));
    printf("  alternate 0 KB / 200 KB smem %6.2f us\n", time_chain([&] { k_smem<<<grid, 320, big, s>>>(out); k_empty<<<grid, 320, 0, s>>>(out); }, s) );
  }
  return 0;
}
